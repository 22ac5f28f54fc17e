// Late fusion with per-query score normalisation WITHOUT the [Q, N] score matrix.
//
// retrieval/fusion.py:4-14,31-42 combines, for one query,
//     fused = w * (ts - a_t) / b_t  +  (1 - w) * (is - a_i) / b_i
// with (a, b) = (mean, std + 1e-8) [z-score] or (min, max - min + 1e-8) [min-max] of that query's N scores.
// Per query this is an affine map of the two similarity vectors:
//     fused[d] = < [g_t * Tq ; g_i * Iq], [Td ; Id] >  -  c,      g_t = w / b_t,  g_i = (1 - w) / b_i,
//                                                                 c   = g_t * a_t + g_i * a_i
// so the Top-K comes out of the ordinary fused search (K2) with the query segments scaled per row
// (emr2a_scale_segments) and the constant applied to the K winning scores afterwards (emr2a_keys_add_offset).
// The statistics (a, b) come from database moments (z-score) or two K=1 searches (min-max); see
// Engine.late_fusion_search.
#include "common.cuh"

namespace emr2a {

__global__ void __launch_bounds__(256) scale_segments_kernel(float* __restrict__ x, int64_t n, int d0, int d1, int64_t ld,
                                                             const float* __restrict__ g0, const float* __restrict__ g1) {
  const int64_t row = blockIdx.x;
  const float a = g0[row], b = d1 > 0 ? g1[row] : 0.f;
  float* p = x + row * ld;
  for (int c = threadIdx.x; c < d0 + d1; c += blockDim.x) p[c] = __fmul_rn(p[c], c < d0 ? a : b);
}

__global__ void __launch_bounds__(256) keys_add_offset_kernel(uint64_t* __restrict__ keys, int64_t total, int K,
                                                              const float* __restrict__ offset) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint64_t k = keys[i];
  if (k == 0ull) return;
  keys[i] = pack_key(__fadd_rn(key_score(k), offset[i / K]), key_index(k));
}

}  // namespace emr2a

using namespace emr2a;

extern "C" int emr2a_scale_segments(float* x, int64_t n, int d0, int d1, int64_t ld, const float* g0, const float* g1,
                                    void* stream) {
  if (!x || !g0 || n < 0 || d0 <= 0 || d1 < 0 || ld < d0 + d1 || (d1 > 0 && !g1))
    return fail(EMR2A_ERR_INVALID, "scale_segments: bad arguments");
  if (n == 0) return EMR2A_OK;
  if (n > 0x7fffffffLL) return fail(EMR2A_ERR_UNSUPPORTED, "scale_segments: too many rows for one call");
  scale_segments_kernel<<<static_cast<unsigned>(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, d0, d1, ld, g0, g1);
  EMR2A_LAUNCH_CHECK("scale_segments_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_keys_add_offset(uint64_t* keys, int64_t Q, int K, const float* offset, void* stream) {
  if (!keys || !offset || Q < 0 || K <= 0) return fail(EMR2A_ERR_INVALID, "keys_add_offset: bad arguments");
  if (Q == 0) return EMR2A_OK;
  const int64_t total = Q * K;
  keys_add_offset_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, total, K,
                                                                                                                   offset);
  EMR2A_LAUNCH_CHECK("keys_add_offset_kernel");
  return EMR2A_OK;
}
