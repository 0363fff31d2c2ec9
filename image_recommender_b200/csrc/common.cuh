// Shared device/host helpers for the b2k kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/b2k.h"

namespace b2k {

constexpr int kList = B2K_LIST;  // entries per (query, split) partial list

// ---------------------------------------------------------------------------------------
// error plumbing (api.cu owns the thread-local message)
void set_error(const char* fmt, ...);

#define B2K_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      b2k::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return (int)e_;                                                                    \
    }                                                                                    \
  } while (0)

#define B2K_CHECK_LAUNCH() B2K_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------
// (score, row) records.  Partial lists and candidate buffers are arrays of these.
struct __align__(8) Cand {
  float score;
  int32_t row;  // local row in the shard, -1 = empty
};

// Total order used everywhere: higher score first, then lower row/offset.  Mapped to one
// uint64 so that "better" == numerically larger.  NaN scores sort last.
__host__ __device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } v; v.f = f; b = v.u;
#endif
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0u;       // NaN -> worst
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ uint64_t cand_key(float score, int32_t row) {
  // empty slots (row < 0) get key 0: below every real entry
  if (row < 0) return 0ull;
  return ((uint64_t)float_key(score) << 32) | (uint32_t)(0x7fffffff - row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t k = (uint32_t)(key >> 32);
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } v; v.u = b; return v.f;
#endif
}
__host__ __device__ __forceinline__ int32_t key_row(uint64_t key) {
  return key == 0ull ? -1 : (int32_t)(0x7fffffff - (uint32_t)(key & 0xffffffffu));
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// Arithmetic specs shared bit-for-bit with oracle/b2k_oracle.c.
//
// sumsq32 / dot64: element i of a vector belongs to lane (i/4) % 32; each lane folds its
// elements in increasing i with one FMA each; lanes are combined by the xor butterfly
// 16,8,4,2,1 (a+b is commutative, so every lane ends with identical bits).

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Lane-partial Σ x_i² over x[0..d) following the spec (vector loads when 16B-aligned).
__device__ __forceinline__ float lane_sumsq(const float* __restrict__ x, int d, int lane) {
  float p = 0.f;
  const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int c = lane; c < (d >> 2); c += 32) {
      float4 v = x4[c];
      p = __fmaf_rn(v.x, v.x, p); p = __fmaf_rn(v.y, v.y, p);
      p = __fmaf_rn(v.z, v.z, p); p = __fmaf_rn(v.w, v.w, p);
    }
  } else {
    for (int c = lane; c * 4 < d; c += 32) {
      for (int j = 0; j < 4; ++j) {
        int i = c * 4 + j;
        if (i < d) { float v = x[i]; p = __fmaf_rn(v, v, p); }
      }
    }
  }
  return p;
}

// score of an order-preserving key - 2*eps, rounded down (conservative candidate threshold)
__device__ __forceinline__ float key_minus_2eps(uint32_t key, float eps) {
  const uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  return __fsub_rd(__uint_as_float(b), __fmul_ru(2.0f, eps));
}

// Lane-partial of Spec R's exact score: fp32 products accumulated in fp64, element i owned by lane
// (i/4) % 32 and folded in increasing i (combine the lanes with warp_sum_f64, round once to fp32).
// The row is a latency-bound gather, so eight 16-byte loads per lane are issued before the dependent
// fp64 chain consumes them; the FMA order is unchanged.
__device__ __forceinline__ double lane_dot64(const float* __restrict__ q, const float* __restrict__ x, int d, int lane) {
  double p = 0.0;
  const bool vec = ((d & 3) == 0) && (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(x)) & 15) == 0);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    const int n4 = d >> 2;
    int i = lane;
    for (; i + 7 * 32 < n4; i += 8 * 32) {
      float4 v[8], w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { v[u] = __ldg(x4 + i + u * 32); w[u] = __ldg(q4 + i + u * 32); }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p = __fma_rn((double)w[u].x, (double)v[u].x, p); p = __fma_rn((double)w[u].y, (double)v[u].y, p);
        p = __fma_rn((double)w[u].z, (double)v[u].z, p); p = __fma_rn((double)w[u].w, (double)v[u].w, p);
      }
    }
    for (; i < n4; i += 32) {
      const float4 v = __ldg(x4 + i), w = __ldg(q4 + i);
      p = __fma_rn((double)w.x, (double)v.x, p); p = __fma_rn((double)w.y, (double)v.y, p);
      p = __fma_rn((double)w.z, (double)v.z, p); p = __fma_rn((double)w.w, (double)v.w, p);
    }
  } else {
    for (int i4 = lane; i4 * 4 < d; i4 += 32)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i4 * 4 + c;
        if (i < d) p = __fma_rn((double)q[i], (double)x[i], p);
      }
  }
  return p;
}

// Two candidate rows at once against one query: the query chunk is loaded and widened once, sixteen 16-byte row
// loads per lane are in flight instead of eight (the re-rank of a small batch is a chain of latency-bound row
// gathers: half the rounds).  Per row the FMA order is exactly lane_dot64's.
__device__ __forceinline__ void lane_dot64_x2(const float* __restrict__ q, const float* __restrict__ x0,
                                              const float* __restrict__ x1, int d, int lane, double& p0, double& p1) {
  p0 = 0.0; p1 = 0.0;
  const bool vec = ((d & 3) == 0) && (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(x0) |
                                        reinterpret_cast<uintptr_t>(x1)) & 15) == 0);
  if (!vec) { p0 = lane_dot64(q, x0, d, lane); p1 = lane_dot64(q, x1, d, lane); return; }
  const float4* a4 = reinterpret_cast<const float4*>(x0);
  const float4* b4 = reinterpret_cast<const float4*>(x1);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const int n4 = d >> 2;
  int i = lane;
  for (; i + 7 * 32 < n4; i += 8 * 32) {
    float4 va[8], vb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { va[u] = __ldg(a4 + i + u * 32); vb[u] = __ldg(b4 + i + u * 32); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 w = __ldg(q4 + i + u * 32);
      const double wx = (double)w.x, wy = (double)w.y, wz = (double)w.z, ww = (double)w.w;
      p0 = __fma_rn(wx, (double)va[u].x, p0); p0 = __fma_rn(wy, (double)va[u].y, p0);
      p0 = __fma_rn(wz, (double)va[u].z, p0); p0 = __fma_rn(ww, (double)va[u].w, p0);
      p1 = __fma_rn(wx, (double)vb[u].x, p1); p1 = __fma_rn(wy, (double)vb[u].y, p1);
      p1 = __fma_rn(wz, (double)vb[u].z, p1); p1 = __fma_rn(ww, (double)vb[u].w, p1);
    }
  }
  for (; i < n4; i += 32) {
    const float4 va = __ldg(a4 + i), vb = __ldg(b4 + i), w = __ldg(q4 + i);
    const double wx = (double)w.x, wy = (double)w.y, wz = (double)w.z, ww = (double)w.w;
    p0 = __fma_rn(wx, (double)va.x, p0); p0 = __fma_rn(wy, (double)va.y, p0);
    p0 = __fma_rn(wz, (double)va.z, p0); p0 = __fma_rn(ww, (double)va.w, p0);
    p1 = __fma_rn(wx, (double)vb.x, p1); p1 = __fma_rn(wy, (double)vb.y, p1);
    p1 = __fma_rn(wz, (double)vb.z, p1); p1 = __fma_rn(ww, (double)vb.w, p1);
  }
}

// Same arithmetic with the query already widened to fp64 (shared memory): one fp32->fp64 conversion per
// element instead of two — the conversion rate (16/clk/SM), not HBM, bounds a throughput-bound re-rank.
__device__ __forceinline__ double lane_dot64_qd(const double* __restrict__ qd, const float* __restrict__ x, int d, int lane) {
  double p = 0.0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int n4 = d >> 2;
  int i = lane;
  for (; i + 7 * 32 < n4; i += 8 * 32) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(x4 + i + u * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double2 w01 = *reinterpret_cast<const double2*>(qd + 4 * (i + u * 32));
      const double2 w23 = *reinterpret_cast<const double2*>(qd + 4 * (i + u * 32) + 2);
      p = __fma_rn(w01.x, (double)v[u].x, p); p = __fma_rn(w01.y, (double)v[u].y, p);
      p = __fma_rn(w23.x, (double)v[u].z, p); p = __fma_rn(w23.y, (double)v[u].w, p);
    }
  }
  for (; i < n4; i += 32) {
    const float4 v = __ldg(x4 + i);
    const double2 w01 = *reinterpret_cast<const double2*>(qd + 4 * i);
    const double2 w23 = *reinterpret_cast<const double2*>(qd + 4 * i + 2);
    p = __fma_rn(w01.x, (double)v.x, p); p = __fma_rn(w01.y, (double)v.y, p);
    p = __fma_rn(w23.x, (double)v.z, p); p = __fma_rn(w23.y, (double)v.w, p);
  }
  return p;
}

// G-way merge (G <= 32) of per-shard result lists, each already in the global order (higher ip, then
// lower offset) with label < 0 padding at its end: one warp per query, lane g walks list g, every
// step is one warp arg-best.  O(k log G) for any k.  addr(g, j) = flat index of entry j of list g.
// kCoherent: the lists were written by OTHER GPUs during this kernel (K-exchange): every read goes to L2
// (ld.global.cg) — never through the non-coherent path a const __restrict__ pointer permits.
template <bool kCoherent, typename T>
__device__ __forceinline__ T merge_ld(const T* p) { return kCoherent ? __ldcg(p) : *p; }

template <bool kCoherent = false, typename Addr>
__device__ __forceinline__ void warp_merge_sorted(const float* ip, const float* dist,
                                                  const int64_t* lab, Addr addr, int G, int k, int lane,
                                                  float* out_ip, float* out_dist, int64_t* out_lab) {
  int h = 0;
  uint32_t key = 0u;
  int64_t off = INT64_MAX, src = -1;
  bool valid = false;
  auto load = [&]() {
    valid = lane < G && h < k;
    if (valid) {
      src = addr(lane, h);
      off = merge_ld<kCoherent>(lab + src);
      valid = off >= 0;
      if (valid) key = float_key(merge_ld<kCoherent>(ip + src));
    }
  };
  load();
  for (int j = 0; j < k; ++j) {
    uint32_t bk = valid ? key : 0u;
    int64_t bo = valid ? off : INT64_MAX;
    int bl = valid ? lane : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int64_t oo = __shfl_xor_sync(0xffffffffu, bo, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      const bool take = ol >= 0 && (bl < 0 || ok > bk || (ok == bk && (oo < bo || (oo == bo && ol < bl))));
      if (take) { bk = ok; bo = oo; bl = ol; }
    }
    if (bl < 0) {                      // every list is exhausted: faiss-style padding
      if (lane == 0) {
        if (out_ip) out_ip[j] = -3.402823466e38f;
        out_dist[j] = 3.402823466e38f;
        out_lab[j] = -1;
      }
      continue;
    }
    if (lane == bl) {
      if (out_ip) out_ip[j] = merge_ld<kCoherent>(ip + src);
      out_dist[j] = merge_ld<kCoherent>(dist + src);
      out_lab[j] = off;
      ++h;
      load();
    }
  }
}

// Block-wide bitonic sort of P (a power of two) u64 keys in shared memory, descending; 0 = empty
// slot, sorts last.  All threads of the block must call.
__device__ __forceinline__ void block_bitonic_desc(uint64_t* keys, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool desc = (i & k) == 0;
          const uint64_t x = keys[i], y = keys[ixj];
          if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t h) {
  return __uint_as_float(((uint32_t)h) << 16);
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// Synthetic generator (Spec G) — integer hashing + exact int->float steps only, so that
// oracle/b2k_oracle.c (orc_synth_rows) reproduces every bit on the CPU.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// Irwin–Hall(4) of 16-bit uniforms, unit variance.
__host__ __device__ __forceinline__ float gauss4(uint64_t h) {
  int32_t a = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) +
              (int32_t)((h >> 32) & 0xffff) + (int32_t)(h >> 48);
  return (float)(a - 131070) * (1.0f / 37837.8f);
}

}  // namespace b2k
