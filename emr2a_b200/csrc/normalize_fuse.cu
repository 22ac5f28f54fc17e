// K1: fused L2-normalise + weight + concatenate (+ bf16 hi/lo split) -- one pass over HBM.
//
// Roofline: HBM-bound.  Algorithmic bytes per row =
//   (d0 + d1) * sizeof(in)  read  +  (d0 + d1) * 4 (fp32 out, if requested)
//   + 2 * ld_bf16 * 2 (hi/lo planes, if requested) + 4 (inverse norm, if requested).
// One warp owns one row; the row lives in registers between the norm reductions and
// the store (read once, written once), loads are 128-bit and coalesced, streaming
// (L1::no_allocate).  The arithmetic mirrors the reference op for op -- true IEEE
// divisions by (sqrt(sum x^2) + 1e-8f) in fp32 -- so the fp32 output differs from
// numpy only through the summation order of the norm (<= ~1e-7 relative).
#include <type_traits>
#include "row_math.cuh"

namespace emr2a {

struct NfParams {
  const void* seg0;
  const void* seg1;
  int64_t n;
  int d0, d1;
  int64_t ld0, ld1;
  float w0, w1;
  int flags;
  float* out_f32;
  int64_t ld_f32;
  uint16_t* out_hi;
  uint16_t* out_lo;
  int64_t ld_bf16;
  float* inv_norm;
  float* stats;     // [0] = max row norm, [1] = max ||row - bf16(row)||  (atomic max on the float bits)
  const float* col_std;   // EMR2A_NF_STANDARDIZE: [3][d0 + d1] = per-column mean | scale | RN(1 / scale)
  float* row_div;         // optional [n][4]: the divisors this pass used for the row (row_math.cuh: LazyRows)
};

__device__ __forceinline__ void stats_flush(float* stats, float max_norm2, float max_res2) {
  // one atomic pair per warp at the end of the kernel; non-negative floats order like their bit patterns
  unsigned int* u = reinterpret_cast<unsigned int*>(stats);
  atomicMax(u, __float_as_uint(__fsqrt_ru(max_norm2)));
  atomicMax(u + 1, __float_as_uint(__fsqrt_ru(max_res2)));
}

// Register-cached path: every lane keeps MAXC chunks of 4 elements.
template <typename InT, int MAXC>
__device__ __forceinline__ void load_row(const NfParams& p, int64_t row, int lane, int c0, int ctot, float4 (&v)[MAXC]) {
#pragma unroll
  for (int j = 0; j < MAXC; ++j) {
    const int c = lane + 32 * j;
    if (c < c0) v[j] = Loader<InT>::load4(p.seg0, row * p.ld0 + 4 * c);
    else if (c < ctot) v[j] = Loader<InT>::load4(p.seg1, row * p.ld1 + 4 * (c - c0));
    else v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// StandardScaler.transform on the loaded chunk: (x - mean) / scale per column, both IEEE fp32 -- the division as the
// correctly rounded Markstein sequence with the per-column reciprocal the host precomputed (see RowDiv).
__device__ __forceinline__ float std1(float x, float m, float s, float r) {
  const float t = __fsub_rn(x, m);
  const float q0 = t * r;
  const float rem = fmaf(-q0, s, t);
  return fmaf(rem, r, q0);
}
__device__ __forceinline__ void standardize4(float4& v, const float* __restrict__ col_std, int dtot, int col) {
  const float4 m = __ldg(reinterpret_cast<const float4*>(col_std + col));
  const float4 s = __ldg(reinterpret_cast<const float4*>(col_std + dtot + col));
  const float4 r = __ldg(reinterpret_cast<const float4*>(col_std + 2 * dtot + col));
  v.x = std1(v.x, m.x, s.x, r.x); v.y = std1(v.y, m.y, s.y, r.y);
  v.z = std1(v.z, m.z, s.z, r.z); v.w = std1(v.w, m.w, s.w, r.w);
}

template <typename InT, int MAXC, bool WANT_LO, bool STATS, bool STD = false>
__global__ void __launch_bounds__(256, (MAXC <= 8 ? 2 : 1)) normalize_fuse_vec_kernel(const NfParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int c0 = p.d0 >> 2;
  const int ctot = (p.d0 + p.d1) >> 2;
  const int cpad = p.out_hi ? static_cast<int>(p.ld_bf16 >> 2) : ctot;
  constexpr bool PREFETCH = MAXC <= 8;      // next row's loads are in flight while this row is processed
  float max_n2 = 0.f, max_r2 = 0.f;

  float4 nxt[PREFETCH ? MAXC : 1];
  if (PREFETCH && warp0 < p.n) load_row<InT, MAXC>(p, warp0, lane, c0, ctot, reinterpret_cast<float4 (&)[MAXC]>(nxt));

  for (int64_t row = warp0; row < p.n; row += nwarps) {
    float4 v[MAXC];
    if (PREFETCH) {
#pragma unroll
      for (int j = 0; j < MAXC; ++j) v[j] = nxt[PREFETCH ? j : 0];
      if (row + nwarps < p.n) load_row<InT, MAXC>(p, row + nwarps, lane, c0, ctot, reinterpret_cast<float4 (&)[MAXC]>(nxt));
    } else {
      load_row<InT, MAXC>(p, row, lane, c0, ctot, v);
    }
    if (STD) {                              // per-fold StandardScaler fused into this pass (utils/cv_evaluator.py:78-80)
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < ctot) standardize4(v[j], p.col_std, p.d0 + p.d1, 4 * c);
      }
    }
    float ss0 = 0.f, ss1 = 0.f;
    float4 divs = make_float4(1.f, 1.f, 1.f, 0.f);      // what row_div_out records (row_math.cuh: LazyRows)
    if (p.flags & EMR2A_NF_SEGNORM) {
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < c0) ss0 += sq4(v[j]); else ss1 += sq4(v[j]);
      }
    }
    if (p.flags & EMR2A_NF_SEGNORM) {
      const RowDiv n0 = row_div(__fsqrt_rn(warp_sum(ss0)) + EMR2A_EPS);
      const RowDiv n1 = row_div(__fsqrt_rn(warp_sum(ss1)) + EMR2A_EPS);
      divs.x = n0.d; divs.y = n1.d;
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < c0) div4(v[j], n0); else if (c < ctot) div4(v[j], n1);
      }
    }
    if (p.w0 != 1.0f || p.w1 != 1.0f) {
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < c0) mul4(v[j], p.w0); else if (c < ctot) mul4(v[j], p.w1);
      }
    }
    float inv = 1.0f;
    if (p.flags & (EMR2A_NF_ROWNORM | EMR2A_NF_ZERO_GUARD)) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < MAXC; ++j) ss += sq4(v[j]);
      float nrm = __fsqrt_rn(warp_sum(ss));
      const bool guard = (p.flags & EMR2A_NF_ZERO_GUARD) != 0;
      if (!guard) nrm += EMR2A_EPS;
      if (!(guard && nrm == 0.f)) {
        const RowDiv dv = row_div(nrm);
#pragma unroll
        for (int j = 0; j < MAXC; ++j) div4(v[j], dv);
        inv = dv.r;
        divs.z = nrm; divs.w = 1.f;
      }
    }
    if (p.inv_norm && lane == 0) p.inv_norm[row] = inv;
    if (p.row_div && lane == 0) reinterpret_cast<float4*>(p.row_div)[row] = divs;
    if (p.out_f32) {
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < ctot) *reinterpret_cast<float4*>(p.out_f32 + row * p.ld_f32 + 4 * c) = v[j];
      }
    }
    if (p.out_hi) {
      float n2 = 0.f, r2 = 0.f;
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        if (c < cpad) {
          // packed conversion: one cvt.rn.bf16x2.f32 per pair
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[j].x, v[j].y);
          const __nv_bfloat162 h23 = __floats2bfloat162_rn(v[j].z, v[j].w);
          const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01);
          const uint32_t u23 = *reinterpret_cast<const uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(p.out_hi + row * p.ld_bf16 + 4 * c) = make_uint2(u01, u23);
          if (WANT_LO || STATS) {
            const float e0 = v[j].x - __uint_as_float(u01 << 16);
            const float e1 = v[j].y - __uint_as_float(u01 & 0xFFFF0000u);
            const float e2 = v[j].z - __uint_as_float(u23 << 16);
            const float e3 = v[j].w - __uint_as_float(u23 & 0xFFFF0000u);
            if (STATS) {
              n2 += sq4(v[j]);
              r2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            }
            if (WANT_LO) {
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(e0, e1);
              const __nv_bfloat162 l23 = __floats2bfloat162_rn(e2, e3);
              *reinterpret_cast<uint2*>(p.out_lo + row * p.ld_bf16 + 4 * c) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
          }
        }
      }
      if (STATS) {
        max_n2 = fmaxf(max_n2, warp_sum(n2));
        max_r2 = fmaxf(max_r2, warp_sum(r2));
      }
    }
  }
  if (STATS && lane == 0 && warp0 < p.n) stats_flush(p.stats, max_n2, max_r2);
}

// Wide rows (D > 2048): one 256-thread block per row, CPT chunks of 4 elements per thread, block-wide
// reductions through shared memory; the next row's loads are in flight while this one is processed.
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(a, b);
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < 8; ++w) { t.x += red[w].x; t.y += red[w].y; }
  __syncthreads();
  return t;
}

template <typename InT, int CPT>
__device__ __forceinline__ void load_row_block(const NfParams& p, int64_t row, int c0, int ctot,
                                               typename Loader<InT>::Raw (&v)[CPT]) {
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int c = threadIdx.x + 256 * j;
    if (c < c0) v[j] = Loader<InT>::load_raw(p.seg0, row * p.ld0 + 4 * c);
    else if (c < ctot) v[j] = Loader<InT>::load_raw(p.seg1, row * p.ld1 + 4 * (c - c0));
  }
}

template <typename InT, int CPT>
__global__ void __launch_bounds__(256, 2) normalize_fuse_block_kernel(const NfParams p) {
  typedef typename Loader<InT>::Raw Raw;
  // rows are prefetched UNCONVERTED; bf16 rows are half as many bytes, so two of them are kept in flight
  constexpr int PF = sizeof(Raw) == 16 ? 1 : 2;
  __shared__ float2 red[8];
  const int c0 = p.d0 >> 2;
  const int ctot = (p.d0 + p.d1) >> 2;
  const int cpad = p.out_hi ? static_cast<int>(p.ld_bf16 >> 2) : ctot;
  float max_n2 = 0.f, max_r2 = 0.f;
  Raw ring[PF][CPT];
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    const int64_t r = blockIdx.x + static_cast<int64_t>(s) * gridDim.x;
    if (r < p.n) load_row_block<InT, CPT>(p, r, c0, ctot, ring[s]);
  }
  for (int64_t base = blockIdx.x; base < p.n; base += static_cast<int64_t>(PF) * gridDim.x) {
#pragma unroll
   for (int s = 0; s < PF; ++s) {
    const int64_t row = base + static_cast<int64_t>(s) * gridDim.x;
    if (row >= p.n) break;
    float4 v[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = threadIdx.x + 256 * j;
      v[j] = c < ctot ? Loader<InT>::widen(ring[s][j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {
      const int64_t r2 = row + static_cast<int64_t>(PF) * gridDim.x;
      if (r2 < p.n) load_row_block<InT, CPT>(p, r2, c0, ctot, ring[s]);
    }
    float4 divs = make_float4(1.f, 1.f, 1.f, 0.f);
    if (p.flags & EMR2A_NF_SEGNORM) {
      float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < c0) ss0 += sq4(v[j]); else ss1 += sq4(v[j]);
      }
      const float2 t = block_sum2(ss0, ss1, red);
      const RowDiv n0 = row_div(__fsqrt_rn(t.x) + EMR2A_EPS);
      const RowDiv n1 = row_div(__fsqrt_rn(t.y) + EMR2A_EPS);
      divs.x = n0.d; divs.y = n1.d;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < c0) div4(v[j], n0); else if (c < ctot) div4(v[j], n1);
      }
    }
    if (p.w0 != 1.0f || p.w1 != 1.0f) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < c0) mul4(v[j], p.w0); else if (c < ctot) mul4(v[j], p.w1);
      }
    }
    float inv = 1.0f;
    if (p.flags & (EMR2A_NF_ROWNORM | EMR2A_NF_ZERO_GUARD)) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) ss += sq4(v[j]);
      float nrm = __fsqrt_rn(block_sum2(ss, 0.f, red).x);
      const bool guard = (p.flags & EMR2A_NF_ZERO_GUARD) != 0;
      if (!guard) nrm += EMR2A_EPS;
      if (!(guard && nrm == 0.f)) {
        const RowDiv dv = row_div(nrm);
#pragma unroll
        for (int j = 0; j < CPT; ++j) div4(v[j], dv);
        inv = dv.r;
        divs.z = nrm; divs.w = 1.f;
      }
    }
    if (p.inv_norm && threadIdx.x == 0) p.inv_norm[row] = inv;
    if (p.row_div && threadIdx.x == 0) reinterpret_cast<float4*>(p.row_div)[row] = divs;
    if (p.out_f32) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < ctot) *reinterpret_cast<float4*>(p.out_f32 + row * p.ld_f32 + 4 * c) = v[j];
      }
    }
    if (p.out_hi) {
      float n2 = 0.f, r2 = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = threadIdx.x + 256 * j;
        if (c < cpad) {
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[j].x, v[j].y);
          const __nv_bfloat162 h23 = __floats2bfloat162_rn(v[j].z, v[j].w);
          const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01);
          const uint32_t u23 = *reinterpret_cast<const uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(p.out_hi + row * p.ld_bf16 + 4 * c) = make_uint2(u01, u23);
          if (p.out_lo || p.stats) {
            const float e0 = v[j].x - __uint_as_float(u01 << 16);
            const float e1 = v[j].y - __uint_as_float(u01 & 0xFFFF0000u);
            const float e2 = v[j].z - __uint_as_float(u23 << 16);
            const float e3 = v[j].w - __uint_as_float(u23 & 0xFFFF0000u);
            n2 += sq4(v[j]);
            r2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            if (p.out_lo) {
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(e0, e1);
              const __nv_bfloat162 l23 = __floats2bfloat162_rn(e2, e3);
              *reinterpret_cast<uint2*>(p.out_lo + row * p.ld_bf16 + 4 * c) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
          }
        }
      }
      if (p.stats) {
        const float2 t = block_sum2(n2, r2, red);
        max_n2 = fmaxf(max_n2, t.x);
        max_r2 = fmaxf(max_r2, t.y);
      }
    }
   }
  }
  if (p.stats && threadIdx.x == 0 && blockIdx.x < p.n) stats_flush(p.stats, max_n2, max_r2);
}

// Generic path (any d0/d1/alignment): scalar loads, the row is re-read through L1/L2.
template <typename InT>
__global__ void __launch_bounds__(256) normalize_fuse_scalar_kernel(const NfParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int dtot = p.d0 + p.d1;
  const bool segnorm = (p.flags & EMR2A_NF_SEGNORM) != 0;
  float max_n2 = 0.f, max_r2 = 0.f;
  for (int64_t row = warp0; row < p.n; row += nwarps) {
    float n0 = 1.f, n1 = 1.f;
    if (segnorm) {
      float ss0 = 0.f, ss1 = 0.f;
      for (int e = lane; e < p.d0; e += 32) { float x = Loader<InT>::load1(p.seg0, row * p.ld0 + e); ss0 += x * x; }
      for (int e = lane; e < p.d1; e += 32) { float x = Loader<InT>::load1(p.seg1, row * p.ld1 + e); ss1 += x * x; }
      n0 = __fsqrt_rn(warp_sum(ss0)) + EMR2A_EPS;
      n1 = __fsqrt_rn(warp_sum(ss1)) + EMR2A_EPS;
    }
    auto value = [&](int e) -> float {
      float x;
      if (e < p.d0) {
        x = Loader<InT>::load1(p.seg0, row * p.ld0 + e);
        if (segnorm) x = __fdiv_rn(x, n0);
        x *= p.w0;
      } else {
        x = Loader<InT>::load1(p.seg1, row * p.ld1 + (e - p.d0));
        if (segnorm) x = __fdiv_rn(x, n1);
        x *= p.w1;
      }
      return x;
    };
    float nrm = 1.f, inv = 1.f;
    bool divide = false;
    if (p.flags & (EMR2A_NF_ROWNORM | EMR2A_NF_ZERO_GUARD)) {
      float ss = 0.f;
      for (int e = lane; e < dtot; e += 32) { float x = value(e); ss += x * x; }
      nrm = __fsqrt_rn(warp_sum(ss));
      const bool guard = (p.flags & EMR2A_NF_ZERO_GUARD) != 0;
      if (!guard) nrm += EMR2A_EPS;
      divide = !(guard && nrm == 0.f);
      if (divide) inv = __fdiv_rn(1.0f, nrm);
    }
    if (p.inv_norm && lane == 0) p.inv_norm[row] = inv;
    if (p.row_div && lane == 0) reinterpret_cast<float4*>(p.row_div)[row] = make_float4(n0, n1, divide ? nrm : 1.f, divide ? 1.f : 0.f);
    const int epad = p.out_hi ? static_cast<int>(p.ld_bf16) : dtot;
    float n2 = 0.f, r2 = 0.f;
    for (int e = lane; e < epad; e += 32) {
      float x = 0.f;
      if (e < dtot) {
        x = value(e);
        if (divide) x = __fdiv_rn(x, nrm);
        if (p.out_f32) p.out_f32[row * p.ld_f32 + e] = x;
      }
      if (p.out_hi) {
        uint16_t h, l;
        split_bf16(x, h, l);
        p.out_hi[row * p.ld_bf16 + e] = h;
        if (p.out_lo) p.out_lo[row * p.ld_bf16 + e] = l;
        const float er = x - __uint_as_float(static_cast<uint32_t>(h) << 16);
        n2 += x * x; r2 += er * er;
      }
    }
    if (p.stats && p.out_hi) {
      max_n2 = fmaxf(max_n2, warp_sum(n2));
      max_r2 = fmaxf(max_r2, warp_sum(r2));
    }
  }
  if (p.stats && p.out_hi && lane == 0 && warp0 < p.n) stats_flush(p.stats, max_n2, max_r2);
}

template <typename InT, int MAXC>
static void launch_vec(const NfParams& p, unsigned blocks, int threads, cudaStream_t st) {
  const bool lo = p.out_lo != nullptr, stats = p.stats != nullptr;
  if (p.flags & EMR2A_NF_STANDARDIZE) {       // fp32 rows in, fp32 rows out (checked by the entry point)
    if constexpr (std::is_same<InT, float>::value) normalize_fuse_vec_kernel<InT, MAXC, false, false, true><<<blocks, threads, 0, st>>>(p);
    return;
  }
  if (lo && stats) normalize_fuse_vec_kernel<InT, MAXC, true, true><<<blocks, threads, 0, st>>>(p);
  else if (lo) normalize_fuse_vec_kernel<InT, MAXC, true, false><<<blocks, threads, 0, st>>>(p);
  else if (stats) normalize_fuse_vec_kernel<InT, MAXC, false, true><<<blocks, threads, 0, st>>>(p);
  else normalize_fuse_vec_kernel<InT, MAXC, false, false><<<blocks, threads, 0, st>>>(p);
}

template <typename InT>
static int launch_nf(const NfParams& p, bool vec_ok, cudaStream_t st) {
  const int threads = 256;
  const int64_t warps_needed = p.n;
  const int sms = sm_count();
  const int ctot = (p.d0 + p.d1) >> 2;
  const int cpad = p.out_hi ? static_cast<int>(p.ld_bf16 >> 2) : ctot;
  const int cmax = cpad > ctot ? cpad : ctot;
  int64_t blocks = (warps_needed + 7) / 8;
  if (vec_ok && cmax > 32 * 16 && cmax <= 256 * 8) {
    // wide rows: a block per row
    int64_t g = p.n < static_cast<int64_t>(sms) * 2 ? p.n : static_cast<int64_t>(sms) * 2;
    const unsigned gg = static_cast<unsigned>(g);
    if (cmax <= 256 * 3) normalize_fuse_block_kernel<InT, 3><<<gg, threads, 0, st>>>(p);
    else if (cmax <= 256 * 5) normalize_fuse_block_kernel<InT, 5><<<gg, threads, 0, st>>>(p);
    else normalize_fuse_block_kernel<InT, 8><<<gg, threads, 0, st>>>(p);
  } else if (vec_ok && cmax <= 32 * 48) {
    // persistent grid-stride grid: a multiple of the SM count (resident CTAs per SM follow the register footprint)
    int per_sm = cmax <= 32 * 8 ? 2 : 1;
    const int64_t cap = static_cast<int64_t>(sms) * per_sm;
    if (blocks > cap) blocks = cap;
    const unsigned g = static_cast<unsigned>(blocks);
    if (cmax <= 32 * 4) launch_vec<InT, 4>(p, g, threads, st);
    else if (cmax <= 32 * 8) launch_vec<InT, 8>(p, g, threads, st);
    else if (cmax <= 32 * 16) launch_vec<InT, 16>(p, g, threads, st);
    else launch_vec<InT, 48>(p, g, threads, st);
  } else {
    int64_t cap = static_cast<int64_t>(sms) * 8; if (blocks > cap) blocks = cap;
    normalize_fuse_scalar_kernel<InT><<<static_cast<unsigned>(blocks), threads, 0, st>>>(p);
  }
  EMR2A_LAUNCH_CHECK("normalize_fuse kernel");
  return EMR2A_OK;
}

}  // namespace emr2a

using namespace emr2a;

extern "C" int emr2a_normalize_fuse(const void* seg0, const void* seg1, int64_t n, int d0, int d1,
                                    int64_t ld0, int64_t ld1, float w0, float w1, int flags, int in_dtype,
                                    float* out_f32, int64_t ld_f32, uint16_t* out_hi, uint16_t* out_lo,
                                    int64_t ld_bf16, float* inv_norm_out, float* stats_out, const float* col_std,
                                    float* row_div_out, void* stream) {
  if (n < 0 || d0 <= 0 || d1 < 0) return fail(EMR2A_ERR_INVALID, "normalize_fuse: bad shape n=%lld d0=%d d1=%d", (long long)n, d0, d1);
  if (!seg0 || (d1 > 0 && !seg1)) return fail(EMR2A_ERR_INVALID, "normalize_fuse: null segment pointer");
  if (ld0 < d0 || (d1 > 0 && ld1 < d1)) return fail(EMR2A_ERR_INVALID, "normalize_fuse: leading dimension smaller than row");
  if (in_dtype != EMR2A_F32 && in_dtype != EMR2A_BF16) return fail(EMR2A_ERR_INVALID, "normalize_fuse: unknown dtype %d", in_dtype);
  if (out_lo && !out_hi) return fail(EMR2A_ERR_INVALID, "normalize_fuse: out_lo without out_hi");
  if (stats_out && !out_hi) return fail(EMR2A_ERR_INVALID, "normalize_fuse: stats_out needs the bf16 planes");
  if (out_f32 && ld_f32 < d0 + d1) return fail(EMR2A_ERR_INVALID, "normalize_fuse: ld_f32 too small");
  if (out_hi && ld_bf16 < d0 + d1) return fail(EMR2A_ERR_INVALID, "normalize_fuse: ld_bf16 too small");
  if (n == 0) return EMR2A_OK;
  NfParams p{seg0, seg1, n, d0, d1, ld0, d1 > 0 ? ld1 : 0, w0, w1, flags,
             out_f32, ld_f32, out_hi, out_lo, ld_bf16, inv_norm_out, stats_out, col_std, row_div_out};
  if (row_div_out && ((flags & EMR2A_NF_STANDARDIZE) || (reinterpret_cast<uintptr_t>(row_div_out) & 15)))
    return fail(EMR2A_ERR_UNSUPPORTED, "normalize_fuse: row_div_out needs 16-byte alignment and no fused standardisation");
  const size_t in_align = in_dtype == EMR2A_F32 ? 16 : 8;
  auto aligned = [](const void* q, size_t a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  bool vec_ok = (d0 % 4 == 0) && (d1 % 4 == 0) && (ld0 % 4 == 0) && (d1 == 0 || ld1 % 4 == 0) &&
                aligned(seg0, in_align) && (d1 == 0 || aligned(seg1, in_align));
  if (out_f32) vec_ok = vec_ok && (ld_f32 % 4 == 0) && aligned(out_f32, 16);
  if (out_hi) vec_ok = vec_ok && (ld_bf16 % 4 == 0) && aligned(out_hi, 8) && (!out_lo || aligned(out_lo, 8));
  if (flags & EMR2A_NF_STANDARDIZE) {
    if (!col_std) return fail(EMR2A_ERR_INVALID, "normalize_fuse: EMR2A_NF_STANDARDIZE needs col_std");
    const int cmax = (d0 + d1) >> 2;
    if (in_dtype != EMR2A_F32 || out_hi || !out_f32 || !vec_ok || !aligned(col_std, 16) || cmax > 32 * 16)
      return fail(EMR2A_ERR_UNSUPPORTED, "normalize_fuse: the fused standardisation takes aligned fp32 rows of up to 2048 "
                                         "columns (multiple of 4) and writes fp32 rows only");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == EMR2A_F32) return launch_nf<float>(p, vec_ok, st);
  return launch_nf<__nv_bfloat16>(p, vec_ok, st);
}
