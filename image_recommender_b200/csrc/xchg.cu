// K-exchange: the cross-GPU step of the row-sharded search over NVLink peer memory.
//
// Every rank owns a receive buffer (one slot per source rank, double buffered by epoch parity) that
// all ranks of the box map (CUDA IPC).  After the local top-k is final,
//   push  : one kernel stores the rank's [nq, k] (ip, dist, label) records straight into its slot of
//           EVERY rank's receive buffer (P2P stores through NVSwitch), fences, and the last CTA
//           publishes the epoch in every rank's flag word (release, system scope);
//   merge : one kernel waits (acquire) until all flags show the epoch, then merges the G lists with
//           the global order (higher ip, then lower offset) — the same K-merge as the NCCL path.
// B·k·16 bytes per rank pair (160 B at batch 1): latency-bound, so the win over an NCCL all-gather
// is the removed launch/proxy/copy overhead, not bandwidth.  Double buffering suffices: a rank can
// only be two epochs ahead of a peer after that peer has pushed the epoch in between, which it does
// after finishing its own previous merge (stream order).
#include <string.h>
#include <new>

#include "common.cuh"
#include "kernels.h"

namespace b2k {

struct XchgLayout {
  int64_t cap;          // entries per slot
  int32_t world;
  __host__ __device__ size_t ip_off(int parity, int src) const { return ((size_t)parity * world + src) * cap * 4; }
  __host__ __device__ size_t dist_off(int parity, int src) const { return (size_t)2 * world * cap * 4 + ip_off(parity, src); }
  __host__ __device__ size_t lab_off(int parity, int src) const { return (size_t)4 * world * cap * 4 + ((size_t)parity * world + src) * cap * 8; }
  __host__ __device__ size_t flag_off(int parity, int src) const { return (size_t)8 * world * cap * 4 + ((size_t)parity * world + src) * 8; }
  __host__ __device__ size_t bytes() const { return flag_off(1, world) + 64; }
};

struct XchgPeers { unsigned char* base[16]; };

__global__ void __launch_bounds__(256)
xchg_push_kernel(XchgPeers peers, XchgLayout lay, int rank, int parity, unsigned long long epoch,
                 const float* __restrict__ ip, const float* __restrict__ dist, const int64_t* __restrict__ lab,
                 int64_t n, unsigned int* __restrict__ done_ctas) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const float a = ip[e], d = dist[e];
    const int64_t l = lab[e];
    for (int g = 0; g < lay.world; ++g) {
      unsigned char* b = peers.base[g];
      if (!b) continue;                 // not a target (single-process groups merge on the root only)
      reinterpret_cast<float*>(b + lay.ip_off(parity, rank))[e] = a;
      reinterpret_cast<float*>(b + lay.dist_off(parity, rank))[e] = d;
      reinterpret_cast<int64_t*>(b + lay.lab_off(parity, rank))[e] = l;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_ctas, 1u);
    if (prev == gridDim.x - 1) {          // every CTA's records are visible system-wide: publish
      *done_ctas = 0u;
      __threadfence_system();
      for (int g = 0; g < lay.world; ++g) {
        if (!peers.base[g]) continue;
        unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.base[g] + lay.flag_off(parity, rank));
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
      }
    }
  }
}

// Wait for all ranks' records of this epoch, then merge.  One warp per query.  `mine` is written by the peers
// over NVLink while this kernel may already be spinning: no const/__restrict__ on it, and the records are read
// with L2-coherent loads after the acquire.  A peer that does not show up within kXchgTimeoutNs (a dead or
// stalled rank) must neither hang this GPU nor kill its context (the resident index lives in it): the kernel
// reports the timeout in *status, pads its outputs faiss-style (label -1) and returns; the host raises.
constexpr unsigned long long kXchgTimeoutNs = 10000000000ull;      // 10 s

__global__ void __launch_bounds__(128)
xchg_merge_kernel(unsigned char* mine, XchgLayout lay, int parity, unsigned long long epoch,
                  int nq, int k, float* __restrict__ out_ip, float* __restrict__ out_dist,
                  int64_t* __restrict__ out_labels, unsigned int* status) {
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  if (threadIdx.x < lay.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + lay.flag_off(parity, threadIdx.x));
    unsigned long long v = 0, t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (unsigned int spins = 0;; ++spins) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      if ((spins & 1023u) == 1023u) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > kXchgTimeoutNs) { s_timeout = 1; atomicOr(status, 1u << threadIdx.x); break; }
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  if (s_timeout) {
    for (int j = lane; j < k; j += 32) {
      if (out_ip) out_ip[(int64_t)q * k + j] = -3.402823466e38f;
      out_dist[(int64_t)q * k + j] = 3.402823466e38f;
      out_labels[(int64_t)q * k + j] = -1;
    }
    return;
  }
  const float* ip = reinterpret_cast<const float*>(mine + lay.ip_off(parity, 0));
  const float* dist = reinterpret_cast<const float*>(mine + lay.dist_off(parity, 0));
  const int64_t* lab = reinterpret_cast<const int64_t*>(mine + lay.lab_off(parity, 0));
  const int64_t cap = lay.cap;
  warp_merge_sorted<true>(ip, dist, lab, [=](int g, int j) { return (int64_t)g * cap + (int64_t)q * k + j; }, lay.world, k, lane,
                    out_ip ? out_ip + (int64_t)q * k : nullptr, out_dist + (int64_t)q * k, out_labels + (int64_t)q * k);
}

}  // namespace b2k

using namespace b2k;

struct b2k_xchg {
  int device = 0, rank = 0, world = 1;
  XchgLayout lay;
  unsigned char* mine = nullptr;
  XchgPeers peers;
  bool opened[16] = {false};
  unsigned int* done_ctas = nullptr;      // [0] push: finished CTAs; [1] merge: bit g set = rank g timed out
  unsigned long long epoch = 0;
  bool connected = false;
};

extern "C" {

int b2k_xchg_create(int32_t device, int32_t rank, int32_t world, int64_t max_entries, b2k_xchg** out) {
  if (!out || world < 1 || world > 16 || rank < 0 || rank >= world || max_entries < 1) {
    set_error("xchg_create: bad argument");
    return B2K_E_INVALID;
  }
  *out = nullptr;
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(device));
  b2k_xchg* x = new (std::nothrow) b2k_xchg();
  if (!x) { set_error("xchg_create: out of host memory"); return B2K_E_NOMEM; }
  x->device = device; x->rank = rank; x->world = world;
  x->lay.cap = max_entries; x->lay.world = world;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->mine), x->lay.bytes());
  if (e == cudaSuccess) e = cudaMemset(x->mine, 0, x->lay.bytes());
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->done_ctas), 2 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(x->done_ctas, 0, 2 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("xchg_create: %s", cudaGetErrorString(e));
    if (x->mine) cudaFree(x->mine);
    if (x->done_ctas) cudaFree(x->done_ctas);
    delete x;
    if (prev >= 0) cudaSetDevice(prev);
    return (int)e;
  }
  for (int g = 0; g < 16; ++g) x->peers.base[g] = nullptr;
  x->peers.base[rank] = x->mine;
  x->connected = world == 1;
  if (prev >= 0) cudaSetDevice(prev);
  *out = x;
  return 0;
}

void b2k_xchg_destroy(b2k_xchg* x) {
  if (!x) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int g = 0; g < x->world; ++g)
    if (x->opened[g]) cudaIpcCloseMemHandle(x->peers.base[g]);
  if (x->mine) cudaFree(x->mine);
  if (x->done_ctas) cudaFree(x->done_ctas);
  delete x;
  if (prev >= 0) cudaSetDevice(prev);
}

int b2k_xchg_handle(b2k_xchg* x, void* handle64, void** raw_ptr) {
  if (!x || !handle64) { set_error("xchg_handle: bad argument"); return B2K_E_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  B2K_CUDA(cudaIpcGetMemHandle(&h, x->mine));
  memcpy(handle64, &h, 64);
  if (raw_ptr) *raw_ptr = x->mine;
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_xchg_connect(b2k_xchg* x, const void* handles, const void* const* raw_ptrs) {
  if (!x || (!handles && !raw_ptrs)) { set_error("xchg_connect: bad argument"); return B2K_E_INVALID; }
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  for (int g = 0; g < x->world; ++g) {
    if (g == x->rank) continue;
    if (raw_ptrs) {                       // same-process peers (tests): plain device pointers
      x->peers.base[g] = reinterpret_cast<unsigned char*>(const_cast<void*>(raw_ptrs[g]));
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, reinterpret_cast<const unsigned char*>(handles) + (size_t)g * 64, 64);
      void* p = nullptr;
      B2K_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      x->peers.base[g] = reinterpret_cast<unsigned char*>(p);
      x->opened[g] = true;
    }
  }
  x->connected = true;
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_xchg_push(b2k_xchg* x, const float* ip, const float* dist, const int64_t* labels, int32_t nq,
                  int32_t k, void* stream) {
  if (!x || !x->connected || !ip || !dist || !labels || nq < 1 || k < 1 || (int64_t)nq * k > x->lay.cap) {
    set_error("xchg_push: bad argument or not connected (nq*k <= %lld)", x ? (long long)x->lay.cap : 0ll);
    return B2K_E_INVALID;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  x->epoch += 1;
  const int parity = (int)(x->epoch & 1ull);
  const int64_t n = (int64_t)nq * k;
  int grid = (int)((n + 255) / 256);
  if (grid > 64) grid = 64;
  xchg_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x->peers, x->lay, x->rank, parity, x->epoch, ip, dist,
                                                         labels, n, x->done_ctas);
  B2K_CHECK_LAUNCH();
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_xchg_skip(b2k_xchg* x, void* stream) {
  if (!x || !x->connected) { set_error("xchg_skip: bad argument or not connected"); return B2K_E_INVALID; }
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  x->epoch += 1;
  const int parity = (int)(x->epoch & 1ull);
  // n = 0: no records are stored; the single CTA publishes the epoch in every target's flag word
  xchg_push_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x->peers, x->lay, x->rank, parity, x->epoch, nullptr, nullptr,
                                                       nullptr, 0, x->done_ctas);
  B2K_CHECK_LAUNCH();
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_xchg_merge(b2k_xchg* x, int32_t nq, int32_t k, float* out_ip, float* out_dist, int64_t* out_labels,
                   void* stream) {
  if (!x || !out_dist || !out_labels || nq < 1 || k < 1 || k > B2K_MAX_K || x->epoch == 0) {
    set_error("xchg_merge: bad argument (push first)");
    return B2K_E_INVALID;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  const int parity = (int)(x->epoch & 1ull);
  xchg_merge_kernel<<<(nq + 3) / 4, 128, 0, (cudaStream_t)stream>>>(x->mine, x->lay, parity, x->epoch, nq, k, out_ip,
                                                                  out_dist, out_labels, x->done_ctas + 1);
  B2K_CHECK_LAUNCH();
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_xchg_status(b2k_xchg* x, uint32_t* timed_out_ranks) {
  if (!x || !timed_out_ranks) { set_error("xchg_status: bad argument"); return B2K_E_INVALID; }
  int prev = -1;
  cudaGetDevice(&prev);
  B2K_CUDA(cudaSetDevice(x->device));
  unsigned int v = 0;
  B2K_CUDA(cudaMemcpy(&v, x->done_ctas + 1, sizeof(v), cudaMemcpyDeviceToHost));   // synchronises the device
  if (v) B2K_CUDA(cudaMemset(x->done_ctas + 1, 0, sizeof(v)));
  *timed_out_ranks = v;
  if (prev >= 0) cudaSetDevice(prev);
  if (v) { set_error("xchg: ranks 0x%x did not publish their top-k within 10 s (results were padded with -1)", v); return B2K_E_PEER; }
  return 0;
}

}  // extern "C"
