// b2k_group: ONE process (one caller thread) driving the row shards on several GPUs of a box.
//
// The reference's CLI and ImageRecommender are a single process (main/search_from_image.py:430-441), and SURVEY
// §8(b) sketches "one context drives all G devices".  The torchrun deployment (sharded.py) needs a process group,
// CUDA-IPC handle exchange and NCCL for the plumbing; this context needs none of it: one worker thread per
// device (its CUDA calls block only itself), the devices' receive buffers connected by plain peer pointers
// (cudaDeviceEnablePeerAccess, same address space), every device pushing its [nq, k] records into the ROOT
// device's buffer over NVLink and the root merging them (K-exchange, xchg.cu).  Built purely on the public C
// ABI (b2k_search_device, b2k_xchg_*, b2k_load): nothing here touches a shard's internals.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <string.h>

#include "common.cuh"

using namespace b2k;

namespace {

struct DevState {
  int device = 0;
  b2k_index* shard = nullptr;
  bool owned = false;
  b2k_xchg* xchg = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float* q = nullptr;            // [nq_cap, D]
  float* parts = nullptr;        // [parts_cap, D] image vectors of search_groups
  int32_t* goffs = nullptr;      // [nq_cap + 1]
  int64_t parts_cap = 0;
  float *dist = nullptr, *ip = nullptr;
  int64_t* lab = nullptr;        // local [nq_cap, k_cap]
  float *m_dist = nullptr, *m_ip = nullptr;
  int64_t* m_lab = nullptr;      // merged (root only)
  float run_ms = 0.f;
};

void free_bufs(DevState& d) {
  cudaFree(d.q); cudaFree(d.goffs); cudaFree(d.dist); cudaFree(d.ip); cudaFree(d.lab);
  cudaFree(d.m_dist); cudaFree(d.m_ip); cudaFree(d.m_lab);
  d.q = nullptr; d.goffs = nullptr; d.dist = d.ip = d.m_dist = d.m_ip = nullptr; d.lab = d.m_lab = nullptr;
}

}  // namespace

struct b2k_group {
  int n = 0;
  std::vector<DevState> dev;
  int32_t D = 0;
  int64_t nq_cap = 0, k_cap = 0;
  int32_t last_nq = 0, last_k = 0;
  // worker pool: job(rank) runs on the thread bound to that device
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t gen = 0;
  int pending = 0;
  bool stop = false;
  // the same three, readable without the mutex: workers and the caller poll them for a short while before they
  // sleep on the condition variables (a futex wake-up costs tens of microseconds per hand-off, twice per search:
  // batch 1 on 8 GPUs 0.88 ms against 0.83 ms under torchrun)
  std::atomic<uint64_t> gen_a{0};
  std::atomic<int> pending_a{0};
  std::atomic<bool> stop_a{false};
  std::function<int(int)> job;
  std::vector<int> status;
  std::vector<std::string> errs;
};

namespace {

// Polls `done` for at most `us` microseconds (then the caller sleeps on its condition variable as before).
constexpr int kSpinIdleUs = 200;       // a worker between two searches of a query stream
constexpr int kSpinWaitUs = 3000;      // the caller while the devices work (a batch-1 search takes 0.8 .. 6 ms)
template <typename F>
void spin_until(F done, int us) {
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; !done(); ++i) {
    if ((i & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(us)) return;
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

void worker(b2k_group* g, int rank) {
  cudaSetDevice(g->dev[rank].device);
  uint64_t seen = 0;
  for (;;) {
    std::function<int(int)> job;
    spin_until([&] { return g->stop_a.load(std::memory_order_acquire) || g->gen_a.load(std::memory_order_acquire) != seen; }, kSpinIdleUs);
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_go.wait(lk, [&] { return g->stop || g->gen != seen; });
      if (g->stop) return;
      seen = g->gen;
      job = g->job;
    }
    const int rc = job(rank);
    std::string err = rc ? b2k_last_error() : "";
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->status[rank] = rc;
      g->errs[rank] = err;
      g->pending_a.store(g->pending - 1, std::memory_order_release);
      if (--g->pending == 0) g->cv_done.notify_all();
    }
  }
}

// Runs job(rank) on every device's thread; returns the first failure (its message becomes the caller's).
int run_all(b2k_group* g, std::function<int(int)> job) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->job = std::move(job);
    g->pending = g->n;
    g->pending_a.store(g->n, std::memory_order_release);
    g->gen += 1;
    g->gen_a.store(g->gen, std::memory_order_release);
  }
  g->cv_go.notify_all();
  spin_until([&] { return g->pending_a.load(std::memory_order_acquire) == 0; }, kSpinWaitUs);
  std::unique_lock<std::mutex> lk(g->mu);
  g->cv_done.wait(lk, [&] { return g->pending == 0; });
  // every failing rank is named: the root's "peer did not publish" is usually the CONSEQUENCE of a peer's own error
  std::string all;
  int first = 0;
  for (int r = 0; r < g->n; ++r)
    if (g->status[r]) {
      if (!first) first = g->status[r];
      if (!all.empty()) all += "; ";
      all += "device " + std::to_string(g->dev[r].device) + " (rank " + std::to_string(r) + "): " + g->errs[r];
    }
  if (first) set_error("%s", all.c_str());
  return first;
}

#define G_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error("%s -> %s", #expr, cudaGetErrorString(e_)); return (int)e_; } } while (0)

bool ready(const b2k_group* g) {
  if (!g) return false;
  for (const DevState& d : g->dev) if (!d.shard) return false;
  return true;
}

// (re)allocate the per-device search buffers and the exchange for nq x k results
int ensure_buffers(b2k_group* g, int64_t nq, int64_t k) {
  if (nq <= g->nq_cap && k <= g->k_cap) return 0;
  const int64_t nq_cap = std::max<int64_t>(nq, g->nq_cap), k_cap = std::max<int64_t>(k, g->k_cap);
  int rc = run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    G_CUDA(cudaStreamSynchronize(d.stream));
    free_bufs(d);
    if (d.xchg) { b2k_xchg_destroy(d.xchg); d.xchg = nullptr; }
    G_CUDA(cudaMalloc(&d.q, (size_t)nq_cap * g->D * sizeof(float)));
    G_CUDA(cudaMalloc(&d.goffs, (size_t)(nq_cap + 1) * sizeof(int32_t)));
    G_CUDA(cudaMalloc(&d.dist, (size_t)nq_cap * k_cap * sizeof(float)));
    G_CUDA(cudaMalloc(&d.ip, (size_t)nq_cap * k_cap * sizeof(float)));
    G_CUDA(cudaMalloc(&d.lab, (size_t)nq_cap * k_cap * sizeof(int64_t)));
    if (r == 0) {
      G_CUDA(cudaMalloc(&d.m_dist, (size_t)nq_cap * k_cap * sizeof(float)));
      G_CUDA(cudaMalloc(&d.m_ip, (size_t)nq_cap * k_cap * sizeof(float)));
      G_CUDA(cudaMalloc(&d.m_lab, (size_t)nq_cap * k_cap * sizeof(int64_t)));
    }
    return g->n > 1 ? b2k_xchg_create(d.device, r, g->n, nq_cap * k_cap, &d.xchg) : 0;
  });
  if (rc) return rc;
  g->nq_cap = nq_cap; g->k_cap = k_cap;
  if (g->n == 1) return 0;
  // same address space: the root's receive buffer is a plain peer pointer for everyone; only the root merges,
  // so the other ranks' buffers are not targets (null entries are skipped by the push kernel)
  unsigned char h[64];
  void* root = nullptr;
  rc = b2k_xchg_handle(g->dev[0].xchg, h, &root);
  if (rc) return rc;
  return run_all(g, [=](int r) -> int {
    std::vector<const void*> ptrs((size_t)g->n, nullptr);
    ptrs[0] = root;
    return b2k_xchg_connect(g->dev[r].xchg, nullptr, ptrs.data());
  });
}

// local search + push (+ merge on the root) of nq queries already in d.q on every device
int run_search(b2k_group* g, int32_t nq, int32_t k) {
  g->last_nq = nq; g->last_k = k;
  return run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    G_CUDA(cudaEventRecord(d.ev[0], d.stream));
    const bool single = g->n == 1;
    int rc = b2k_search_device(d.shard, d.q, nq, k, single ? d.m_dist : d.dist, single ? d.m_lab : d.lab,
                               single ? d.m_ip : d.ip, d.stream);
    if (rc) {
      // the root must not wait 10 s for records that will never come: publish the epoch, keep this rank's error
      if (!single && r != 0) {
        const std::string keep = b2k_last_error();
        b2k_xchg_skip(d.xchg, d.stream);
        cudaStreamSynchronize(d.stream);
        set_error("%s", keep.c_str());
      }
      return rc;
    }
    if (!single) {
      rc = b2k_xchg_push(d.xchg, d.ip, d.dist, d.lab, nq, k, d.stream);
      if (rc) return rc;
      if (r == 0) {
        rc = b2k_xchg_merge(d.xchg, nq, k, d.m_ip, d.m_dist, d.m_lab, d.stream);
        if (rc) return rc;
      }
    }
    G_CUDA(cudaEventRecord(d.ev[1], d.stream));
    G_CUDA(cudaStreamSynchronize(d.stream));
    G_CUDA(cudaEventElapsedTime(&d.run_ms, d.ev[0], d.ev[1]));
    if (!single && r == 0) {
      uint32_t missing = 0;
      rc = b2k_xchg_status(d.xchg, &missing);
      if (rc) return rc;
    }
    return 0;
  });
}

}  // namespace

extern "C" {

int b2k_group_create(const int32_t* devices, int32_t n_devices, b2k_group** out) {
  if (!out) { set_error("group_create: out is null"); return B2K_E_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("group_create: no CUDA device (this engine has no CPU path)");
    return B2K_E_NODEVICE;
  }
  // (a device may appear twice: two ranks on one GPU, as the single-GPU tests do; devices == NULL means 0 .. n-1)
  if (n_devices < 1 || n_devices > 16 || (!devices && n_devices > ndev)) { set_error("group_create: %d devices requested, %d visible", n_devices, ndev); return B2K_E_INVALID; }
  b2k_group* g = new (std::nothrow) b2k_group();
  if (!g) { set_error("group_create: out of host memory"); return B2K_E_NOMEM; }
  g->n = n_devices;
  g->dev.resize((size_t)n_devices);
  g->status.assign((size_t)n_devices, 0);
  g->errs.assign((size_t)n_devices, "");
  for (int r = 0; r < n_devices; ++r) {
    g->dev[r].device = devices ? devices[r] : r;
    if (g->dev[r].device < 0 || g->dev[r].device >= ndev) { set_error("group_create: device %d of %d", g->dev[r].device, ndev); delete g; return B2K_E_INVALID; }
  }
  for (int r = 0; r < n_devices; ++r) g->threads.emplace_back(worker, g, r);
  int rc = run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    G_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    G_CUDA(cudaEventCreate(&d.ev[0]));
    G_CUDA(cudaEventCreate(&d.ev[1]));
    // every device stores into the root's receive buffer: peer access to the root (a no-op for the root itself)
    if (r != 0 && d.device != g->dev[0].device) {      // (two ranks on one device: tests on a single-GPU box)
      int can = 0;
      G_CUDA(cudaDeviceCanAccessPeer(&can, d.device, g->dev[0].device));
      if (!can) { set_error("device %d cannot access device %d (no peer path)", d.device, g->dev[0].device); return B2K_E_NODEVICE; }
      cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[0].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess -> %s", cudaGetErrorString(e)); return (int)e; }
      cudaGetLastError();
    }
    return 0;
  });
  if (rc) { b2k_group_destroy(g); return rc; }
  *out = g;
  return 0;
}

void b2k_group_destroy(b2k_group* g) {
  if (!g) return;
  if (!g->threads.empty()) {
    run_all(g, [=](int r) -> int {
      DevState& d = g->dev[r];
      if (d.stream) cudaStreamSynchronize(d.stream);
      free_bufs(d);
      cudaFree(d.parts); d.parts = nullptr;
      if (d.xchg) b2k_xchg_destroy(d.xchg);
      if (d.owned && d.shard) b2k_destroy(d.shard);
      for (int i = 0; i < 2; ++i) if (d.ev[i]) cudaEventDestroy(d.ev[i]);
      if (d.stream) cudaStreamDestroy(d.stream);
      return 0;
    });
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->stop = true;
      g->stop_a.store(true, std::memory_order_release);
    }
    g->cv_go.notify_all();
    for (std::thread& t : g->threads) t.join();
  }
  delete g;
}

int32_t b2k_group_size(const b2k_group* g) { return g ? g->n : 0; }

int b2k_group_set_shard(b2k_group* g, int32_t rank, b2k_index* shard) {
  if (!g || rank < 0 || rank >= g->n || !shard) { set_error("group_set_shard: bad argument"); return B2K_E_INVALID; }
  if (g->D != 0 && b2k_dim(shard) != g->D) { set_error("group_set_shard: shard dimension %d, group %d", b2k_dim(shard), g->D); return B2K_E_INVALID; }
  DevState& d = g->dev[rank];
  if (d.owned && d.shard) b2k_destroy(d.shard);
  d.shard = shard; d.owned = false;
  g->D = b2k_dim(shard);
  return 0;
}

b2k_index* b2k_group_shard(b2k_group* g, int32_t rank) { return g && rank >= 0 && rank < g->n ? g->dev[rank].shard : nullptr; }

int b2k_group_load(b2k_group* g, const char* path) {
  if (!g || !path) { set_error("group_load: bad argument"); return B2K_E_INVALID; }
  int64_t n_rows = 0;
  int rc = b2k_file_info(path, &n_rows, nullptr, nullptr, nullptr);
  if (rc) return rc;
  const int64_t per = (n_rows + g->n - 1) / g->n;          // shard_range() of sharded.py
  const std::string p(path);
  rc = run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    if (d.owned && d.shard) { b2k_destroy(d.shard); d.shard = nullptr; }
    const int64_t r0 = std::min<int64_t>(n_rows, r * per), r1 = std::min<int64_t>(n_rows, r0 + per);
    int rc2 = b2k_load(p.c_str(), d.device, r0, r1, 0, &d.shard);
    d.owned = rc2 == 0;
    return rc2;
  });
  if (rc) return rc;
  g->D = b2k_dim(g->dev[0].shard);
  return 0;
}

int64_t b2k_group_ntotal(const b2k_group* g) {
  int64_t n = 0;
  if (g) for (const DevState& d : g->dev) n += d.shard ? b2k_ntotal(d.shard) : 0;
  return n;
}
int32_t b2k_group_dim(const b2k_group* g) { return g ? g->D : 0; }

int b2k_group_put_queries(b2k_group* g, const float* q_host, int32_t nq, int32_t k) {
  if (!ready(g) || !q_host || nq < 1 || k < 1 || k > B2K_MAX_K) { set_error("group_put_queries: bad argument (every shard set?)"); return B2K_E_INVALID; }
  int rc = ensure_buffers(g, nq, k);
  if (rc) return rc;
  return run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    G_CUDA(cudaMemcpyAsync(d.q, q_host, (size_t)nq * g->D * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    G_CUDA(cudaStreamSynchronize(d.stream));
    return 0;
  });
}

int b2k_group_run(b2k_group* g, int32_t nq, int32_t k) {
  if (!ready(g) || nq < 1 || nq > g->nq_cap || k < 1 || k > g->k_cap) { set_error("group_run: put_queries first"); return B2K_E_INVALID; }
  return run_search(g, nq, k);
}

int b2k_group_get_results(b2k_group* g, float* dist_host, int64_t* labels_host, float* ip_host) {
  if (!ready(g) || !dist_host || !labels_host || g->last_nq < 1) { set_error("group_get_results: run first"); return B2K_E_INVALID; }
  const size_t m = (size_t)g->last_nq * g->last_k;
  DevState& d = g->dev[0];
  int prev = -1;
  cudaGetDevice(&prev);
  G_CUDA(cudaSetDevice(d.device));
  G_CUDA(cudaMemcpy(dist_host, d.m_dist, m * sizeof(float), cudaMemcpyDeviceToHost));
  G_CUDA(cudaMemcpy(labels_host, d.m_lab, m * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (ip_host) G_CUDA(cudaMemcpy(ip_host, d.m_ip, m * sizeof(float), cudaMemcpyDeviceToHost));
  if (prev >= 0) cudaSetDevice(prev);
  return 0;
}

int b2k_group_search(b2k_group* g, const float* q_host, int32_t nq, int32_t k, float* dist_host, int64_t* labels_host,
                     float* ip_host) {
  int rc = b2k_group_put_queries(g, q_host, nq, k);
  if (!rc) rc = b2k_group_run(g, nq, k);
  if (!rc) rc = b2k_group_get_results(g, dist_host, labels_host, ip_host);
  return rc;
}

int b2k_group_search_groups(b2k_group* g, const float* parts_host, int64_t n_images, const int32_t* group_offsets,
                            int32_t n_groups, int32_t k, float* dist_host, int64_t* labels_host, float* ip_host) {
  if (!ready(g) || !parts_host || !group_offsets || n_groups < 1 || k < 1 || k > B2K_MAX_K || n_images < n_groups ||
      group_offsets[0] != 0 || group_offsets[n_groups] != n_images) {
    set_error("group_search_groups: bad argument");
    return B2K_E_INVALID;
  }
  for (int i = 0; i < n_groups; ++i)
    if (group_offsets[i + 1] <= group_offsets[i]) { set_error("group_search_groups: group %d is empty", i); return B2K_E_INVALID; }
  int rc = ensure_buffers(g, n_groups, k);
  if (rc) return rc;
  rc = run_all(g, [=](int r) -> int {
    DevState& d = g->dev[r];
    if (n_images > d.parts_cap) {
      G_CUDA(cudaStreamSynchronize(d.stream));
      cudaFree(d.parts); d.parts = nullptr; d.parts_cap = 0;
      G_CUDA(cudaMalloc(&d.parts, (size_t)n_images * g->D * sizeof(float)));
      d.parts_cap = n_images;
    }
    G_CUDA(cudaMemcpyAsync(d.parts, parts_host, (size_t)n_images * g->D * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    G_CUDA(cudaMemcpyAsync(d.goffs, group_offsets, (size_t)(n_groups + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
    return b2k_prep_groups_device(d.parts, d.goffs, n_groups, g->D, d.q, d.device, d.stream);
  });
  if (!rc) rc = run_search(g, n_groups, k);
  if (!rc) rc = b2k_group_get_results(g, dist_host, labels_host, ip_host);
  return rc;
}

int b2k_group_last_run_ms(const b2k_group* g, float* max_ms) {
  if (!g || !max_ms) { set_error("group_last_run_ms: bad argument"); return B2K_E_INVALID; }
  float m = 0.f;
  for (const DevState& d : g->dev) m = std::max(m, d.run_ms);
  *max_ms = m;
  return 0;
}

}  // extern "C"
