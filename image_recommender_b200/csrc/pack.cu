// K-pack / K-prep: per-table L2-normalise + concat + bf16-pack (build half of the hot path)
// and the query-side normalise / bf16 prep.  Replaces the arithmetic of
//   main/create_index.py:176-188 (_process_batch: asarray(float32).ravel + np.concatenate)
//   main/create_index.py:310-311 (np.stack(...).astype("float32"); index.add)
//   main/search_from_image.py:322 (faiss.normalize_L2)
// HBM-bound, one warp per row, 16-byte accesses, fixed summation order (oracle/b2k_oracle.c
// follows the same order bit for bit).
#include "common.cuh"
#include "kernels.h"

namespace b2k {

// One warp per row.  For each table: s = Σx² (spec order), inv = 1/sqrtf(s) if s > 0,
// y = x*inv -> fp32 row, bf16(y) -> bf16 row; also ||y||² and ||bf16(y)-y||² per row.
__global__ void __launch_bounds__(256)
pack_rows_kernel(PackArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= a.n) return;
  const int64_t r_out = a.row0 + warp;
  float* __restrict__ of = a.out_f32 ? a.out_f32 + r_out * (int64_t)a.D : nullptr;
  uint16_t* __restrict__ ob = a.out_bf16 ? a.out_bf16 + r_out * (int64_t)a.Dp : nullptr;

  float n2 = 0.f, e2 = 0.f;
  for (int t = 0; t < a.n_tables; ++t) {
    const int d = a.dims[t];
    const int off = a.col_off[t];
    const float* __restrict__ x = a.tables[t] + warp * (int64_t)a.strides[t];
    float inv = 1.0f;
    if (a.normalize) {
      const float s = warp_sum_f32(lane_sumsq(x, d, lane));
      if (s > 0.f) inv = __fdiv_rn(1.0f, __fsqrt_rn(s));
    }
    float p2 = 0.f, pe = 0.f;
    const bool vec = ((d & 3) == 0) && ((off & 3) == 0) && ((a.D & 3) == 0) && ((a.Dp & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      // the second read of x hits L1 (the warp just streamed the row for the norm)
      for (int c = lane; c < (d >> 2); c += 32) {
        float4 v = x4[c];
        v.x = __fmul_rn(v.x, inv); v.y = __fmul_rn(v.y, inv);
        v.z = __fmul_rn(v.z, inv); v.w = __fmul_rn(v.w, inv);
        const uint16_t b0 = f32_to_bf16_bits(v.x), b1 = f32_to_bf16_bits(v.y);
        const uint16_t b2 = f32_to_bf16_bits(v.z), b3 = f32_to_bf16_bits(v.w);
        p2 = __fmaf_rn(v.x, v.x, p2); p2 = __fmaf_rn(v.y, v.y, p2);
        p2 = __fmaf_rn(v.z, v.z, p2); p2 = __fmaf_rn(v.w, v.w, p2);
        float d0 = __fsub_rn(bf16_bits_to_f32(b0), v.x), d1 = __fsub_rn(bf16_bits_to_f32(b1), v.y);
        float d2 = __fsub_rn(bf16_bits_to_f32(b2), v.z), d3 = __fsub_rn(bf16_bits_to_f32(b3), v.w);
        pe = __fmaf_rn(d0, d0, pe); pe = __fmaf_rn(d1, d1, pe);
        pe = __fmaf_rn(d2, d2, pe); pe = __fmaf_rn(d3, d3, pe);
        if (of) __stcs(reinterpret_cast<float4*>(of + off + 4 * c), v);      // streaming stores: the
        if (ob) {                                                             // packed rows are not re-read
          uint2 pk;
          pk.x = (uint32_t)b0 | ((uint32_t)b1 << 16);
          pk.y = (uint32_t)b2 | ((uint32_t)b3 << 16);
          __stcs(reinterpret_cast<uint2*>(ob + off + 4 * c), pk);
        }
      }
    } else {
      for (int c = lane; c * 4 < d; c += 32) {
        for (int j = 0; j < 4; ++j) {
          const int i = c * 4 + j;
          if (i >= d) break;
          const float v = __fmul_rn(x[i], inv);
          const uint16_t b = f32_to_bf16_bits(v);
          p2 = __fmaf_rn(v, v, p2);
          const float df = __fsub_rn(bf16_bits_to_f32(b), v);
          pe = __fmaf_rn(df, df, pe);
          if (of) of[off + i] = v;
          if (ob) ob[off + i] = b;
        }
      }
    }
    n2 = __fadd_rn(n2, warp_sum_f32(p2));
    e2 = __fadd_rn(e2, warp_sum_f32(pe));
  }
  if (ob) for (int i = a.D + lane; i < a.Dp; i += 32) ob[i] = 0;
  if (lane == 0) {
    if (a.out_norm2) a.out_norm2[r_out] = n2;
    // non-negative floats order like their bit patterns; NaN (bad input rows) poisons the
    // bound on purpose: its pattern is above +inf, so certificates fail -> exact scan.
    // The maxima only grow: skip the atomic unless this row raises them (one hot L2 address
    // shared by every warp would otherwise serialise the whole kernel's tail).
    if (a.stat_bits) {
      const unsigned int eb = __float_as_uint(e2), nb = __float_as_uint(n2);
      if (eb > __ldcg(a.stat_bits + 0)) atomicMax(a.stat_bits + 0, eb);
      if (nb > __ldcg(a.stat_bits + 1)) atomicMax(a.stat_bits + 1, nb);
    }
  }
}

// faiss.normalize_L2 semantics: x *= 1/sqrtf(Σx²) iff Σx² > 0 (spec order), in place.
__global__ void __launch_bounds__(256)
normalize_rows_kernel(float* __restrict__ x, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= n) return;
  float* row = x + warp * (int64_t)d;
  const float s = warp_sum_f32(lane_sumsq(row, d, lane));
  if (!(s > 0.f)) return;
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(s));
  for (int i = lane; i < d; i += 32) row[i] = __fmul_rn(row[i], inv);
}

// Query acquisition for groups of images (main/search_from_image.py:305-322: per image the concatenated
// parts [1, D]; np.mean over the images of the group; faiss.normalize_L2 of the mean).  One warp per group:
//   mean : fp32, images added in order, then one division by the count — np.mean's arithmetic for a float32
//          stack reduced along its first axis (sequential row adds, true_divide by n);
//   norm : Spec S sum of squares of the mean, x *= 1/sqrtf(s) iff s > 0 (normalize_rows_kernel's arithmetic).
// parts: [n_images, D] row-major, group g = rows [offs[g], offs[g+1]); q_out: [n_groups, D].
__global__ void __launch_bounds__(256)
group_mean_normalize_kernel(const float* __restrict__ parts, const int32_t* __restrict__ offs, int n_groups, int d,
                            float* q_out) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_groups) return;
  const int r0 = offs[g], r1 = offs[g + 1];
  const float cnt = (float)(r1 - r0);
  float* row = q_out + (int64_t)g * d;
  // element i belongs to lane (i / 4) % 32 and is folded in increasing i, exactly as lane_sumsq does: the sum
  // of squares is accumulated while the mean is produced, and every lane later re-reads only what it wrote
  float p = 0.f;
  for (int c = lane; c * 4 < d; c += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c * 4 + j;
      if (i >= d) break;
      float acc = parts[(int64_t)r0 * d + i];
      for (int r = r0 + 1; r < r1; ++r) acc = __fadd_rn(acc, parts[(int64_t)r * d + i]);
      const float m = __fdiv_rn(acc, cnt);
      row[i] = m;
      p = __fmaf_rn(m, m, p);
    }
  }
  const float s = warp_sum_f32(p);
  if (!(s > 0.f)) return;
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(s));
  for (int c = lane; c * 4 < d; c += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c * 4 + j;
      if (i < d) row[i] = __fmul_rn(row[i], inv);
    }
  }
}

// Query prep: fp32 [nq, D] -> bf16 [nq_pad, Dp] (zero padded), ||q||², and the per-query bound
// eps >= |approximate score - exact score| over every row of the shard (DESIGN.md "certificate"):
//   tensor path : q_b·x_b vs q·x  ->  ||q_b||·E + ||q - q_b||·(X + E) + D·2^-22·||q_b||·(X + E)
//   scan path   : q·x_b   vs q·x  ->  ||q||·E                      + D·2^-23·||q||·(X + E)
// with E = max_r ||bf16(x_r) - x_r||, X = max_r ||x_r|| from the shard's pack statistics; the
// last terms bound the fp32 accumulation (tensor core: 4 ulp per accumulated product assumed).
// Every norm is inflated by 1.001 against its own rounding; + 2^-23·||q||·X covers Spec R's
// single rounding.
__global__ void __launch_bounds__(256)
query_prep_kernel(QueryPrepArgs a) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= a.nq_pad) return;
  // q_bf16 == nullptr: the K-scan path multiplies the fp32 query directly (nothing to pack)
  uint16_t* ob = a.q_bf16 ? a.q_bf16 + (int64_t)w * a.Dp : nullptr;
  if (w >= a.nq) {
    if (ob) {   // zero rows pad the query tile (Dp is a multiple of 64: 16-byte stores)
      uint4* o4 = reinterpret_cast<uint4*>(ob);
      for (int i = lane; i < (a.Dp >> 3); i += 32) o4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const float* q = a.q + (int64_t)w * a.D;
  const float qn2 = warp_sum_f32(lane_sumsq(q, a.D, lane));
  float pb = 0.f, pd = 0.f;
  const bool vec = ((a.D & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  if (vec) {
    // 4 elements per lane and step (16-byte loads, 8-byte bf16 stores); Dp is a multiple of 64
    const float4* q4 = reinterpret_cast<const float4*>(q);
    for (int c = lane; c < (a.Dp >> 2); c += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < (a.D >> 2)) v = q4[c];
      const uint16_t b0 = f32_to_bf16_bits(v.x), b1 = f32_to_bf16_bits(v.y);
      const uint16_t b2 = f32_to_bf16_bits(v.z), b3 = f32_to_bf16_bits(v.w);
      const float f0 = bf16_bits_to_f32(b0), f1 = bf16_bits_to_f32(b1), f2 = bf16_bits_to_f32(b2), f3 = bf16_bits_to_f32(b3);
      pb = __fmaf_rn(f0, f0, pb); pb = __fmaf_rn(f1, f1, pb); pb = __fmaf_rn(f2, f2, pb); pb = __fmaf_rn(f3, f3, pb);
      const float d0 = __fsub_rn(f0, v.x), d1 = __fsub_rn(f1, v.y), d2 = __fsub_rn(f2, v.z), d3 = __fsub_rn(f3, v.w);
      pd = __fmaf_rn(d0, d0, pd); pd = __fmaf_rn(d1, d1, pd); pd = __fmaf_rn(d2, d2, pd); pd = __fmaf_rn(d3, d3, pd);
      if (ob) {
        uint2 pk;
        pk.x = (uint32_t)b0 | ((uint32_t)b1 << 16);
        pk.y = (uint32_t)b2 | ((uint32_t)b3 << 16);
        *reinterpret_cast<uint2*>(ob + 4 * c) = pk;
      }
    }
  } else {
    for (int i = lane; i < a.Dp; i += 32) {
      uint16_t b = 0;
      if (i < a.D) {
        const float v = q[i];
        b = f32_to_bf16_bits(v);
        const float vb = bf16_bits_to_f32(b);
        pb = __fmaf_rn(vb, vb, pb);
        const float df = __fsub_rn(vb, v);
        pd = __fmaf_rn(df, df, pd);
      }
      if (ob) ob[i] = b;
    }
  }
  const float qb2 = warp_sum_f32(pb), qd2 = warp_sum_f32(pd);
  if (lane == 0) {
    if (a.floor_init) a.floor_init[w] = -INFINITY;
    a.qn2[w] = qn2;
    const float E = sqrtf(__uint_as_float(a.stat_bits[0])) * 1.001f;
    const float X = sqrtf(__uint_as_float(a.stat_bits[1])) * 1.001f;
    const float qb = sqrtf(qb2) * 1.001f, qd = sqrtf(qd2) * 1.001f, qn = sqrtf(qn2) * 1.001f;
    const float XE = X + E;
    const float r1 = 1.1920929e-07f * qn * X;                       // 2^-23
    a.eps_scan[w] = qn * E * 1.001f + (float)a.D * 1.1920929e-07f * qn * XE + r1;
    a.eps_tc[w] = (qb * E + qd * XE) * 1.001f + (float)a.D * 2.3841858e-07f * qb * XE + r1;
  }
}

// ---------------------------------------------------------------------------------------
// Synthetic rows (Spec G, mirrored by oracle/b2k_oracle.c: orc_synth_rows): raw (un-normalised) per-table fp32.
__device__ __forceinline__ uint64_t h2(uint64_t a, uint64_t b) { return mix64(mix64(a) ^ b); }

__global__ void __launch_bounds__(256)
synth_rows_kernel(SynthArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= a.n) return;
  uint64_t r;          // global DB row this output row is drawn from
  uint64_t hq = 0;     // per-query noise stream
  if (a.query_mode) {
    const uint64_t i = (uint64_t)(a.first + warp);
    r = h2(a.qseed ^ 0x71726f77ull, i) % (uint64_t)a.total_rows;
    hq = mix64(h2(a.qseed ^ 0x716e6f69ull, i));
  } else {
    r = (uint64_t)(a.first + warp);
  }
  const uint64_t c = h2(a.seed ^ 0x636c7573ull, r) % (uint64_t)a.n_clusters;
  const uint64_t hc = mix64(h2(a.seed ^ 0x63656e74ull, c));
  const uint64_t hn = mix64(h2(a.seed ^ 0x6e6f6973ull, r));
  for (int t = 0; t < a.n_tables; ++t) {
    const int d = a.dims[t];
    float* out = a.tables[t] + warp * (int64_t)d;
    const bool ab = (a.abs_mask >> t) & 1u;
    for (int j = lane; j < d; j += 32) {
      const uint64_t key = ((uint64_t)t << 32) | (uint32_t)j;
      float x = __fmaf_rn(a.sigma, gauss4(mix64(hn ^ key)), gauss4(mix64(hc ^ key)));
      if (a.query_mode) x = __fmaf_rn(a.sigma_q, gauss4(mix64(hq ^ key)), x);
      out[j] = ab ? fabsf(x) : x;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Two alternatives were built and measured in round 2 and lost to this kernel (profiles/experiments/r02_exp_pack.log,
// 131072 combo rows per launch; cool chip at 1965 MHz / after 3 s of GEMM at ~1300 MHz under the power cap):
//   this two-pass kernel (second pass hits L1)                      0.435 ms / 0.53 ms
//   register-resident rows, every load issued up front (112 regs)   0.545 ms / 0.70 ms   (16 warps per SM)
//   cp.async.bulk rows through shared memory, 16 warps x 1 slot     0.467 ms / 0.61 ms   (row arithmetic issue-bound)
// A device-to-device copy of the same size runs at 6513 GB/s cool and 5995 GB/s in the capped phase: the memory
// system itself slows under the cap, and this kernel sits at 0.91 / 0.81 of the copy measured in the same phase.
int launch_pack(const PackArgs& a, cudaStream_t st) {
  if (a.n <= 0) return 0;
  const int wpb = 8;
  const int64_t blocks = (a.n + wpb - 1) / wpb;
  pack_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, st>>>(a);
  B2K_CHECK_LAUNCH();
  return 0;
}
int launch_normalize(float* x, int64_t n, int d, cudaStream_t st) {
  if (n <= 0) return 0;
  const int wpb = 8;
  normalize_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(x, n, d);
  B2K_CHECK_LAUNCH();
  return 0;
}
int launch_group_prep(const float* parts, const int32_t* offs, int n_groups, int d, float* q_out, cudaStream_t st) {
  if (n_groups <= 0) return 0;
  const int wpb = 8;
  group_mean_normalize_kernel<<<(n_groups + wpb - 1) / wpb, wpb * 32, 0, st>>>(parts, offs, n_groups, d, q_out);
  B2K_CHECK_LAUNCH();
  return 0;
}
int launch_query_prep(const QueryPrepArgs& a, cudaStream_t st) {
  const int wpb = 8;
  query_prep_kernel<<<(a.nq_pad + wpb - 1) / wpb, wpb * 32, 0, st>>>(a);
  B2K_CHECK_LAUNCH();
  return 0;
}
int launch_synth(const SynthArgs& a, cudaStream_t st) {
  if (a.n <= 0) return 0;
  const int wpb = 8;
  synth_rows_kernel<<<(unsigned)((a.n + wpb - 1) / wpb), wpb * 32, 0, st>>>(a);
  B2K_CHECK_LAUNCH();
  return 0;
}

}  // namespace b2k
