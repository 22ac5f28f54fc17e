// Per-fold preprocessing kernels: the HBM-bound halves of
// StandardScaler -> PCA as CVRetrievalEvaluator.process_embeddings applies them to the train fold
// (utils/cv_evaluator.py:73-93; retrieval/evaluator.py:44-73).
//
//   emr2a_column_moments   per-column sum / sum of squares of (x - shift) over a block of rows, float64
//                          accumulators as in sklearn's StandardScaler (it reduces float32 input in float64),
//                          deterministic (fixed partition + ordered second stage, no atomics)
//   emr2a_standardize      out = (x - f32(mean)) / f32(scale), both operations IEEE fp32: StandardScaler.transform
//                          casts mean_ / scale_ to the dtype of X first (`X -= astype(mean_, X.dtype)`, sklearn 1.9)
//
//   emr2a_gram_f64         G = Z^T Z and column sums of Z in float64, Z = the standardised rows computed ON THE FLY
//                          from the raw fp32 rows (sklearn's fp32 `(x - mean) / scale`): the PCA fit reads the raw
//                          train rows once per tile column and never materialises Z.  Hand-written float64 FMA
//                          contraction (CUDA cores; B200 has a full-rate FP64 pipe), 64 x 64 tiles, upper
//                          triangle only, fixed row partition + ordered second stage: deterministic.
//
// The first two read every element once with 128-bit loads (a warp covers 512 contiguous bytes of a row) and keep
// the per-column constants in registers.  The projection of the PCA is emr2a_project (simt_paths.cu): the fp32 FMA
// GEMM with the standardisation fused into its operand load.  Only the D x D symmetric eigen-decomposition is a
// library call (cuSOLVER through torch), issued by the host layer (emr2a_b200/preprocess.py).
#include "common.cuh"

namespace emr2a {

constexpr int PP_THREADS = 256;
constexpr int PP_ROWS = PP_THREADS / 32;     // row lanes per block
constexpr int PP_COLS = 128;                 // columns per block (32 threads x 4)

struct Quad { float v[4]; };

// 4 consecutive columns of one row; `vec` promises 16-byte alignment and 4 valid columns
__device__ __forceinline__ Quad load_quad(const float* __restrict__ row, int c0, int valid, bool vec) {
  Quad q;
  if (vec) {
    const float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(row + c0));
    q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) q.v[j] = j < valid ? __ldg(row + c0 + j) : 0.f;
  }
  return q;
}

// partial[(by * 2 + {0,1}) * D + c] = sum / sum of squares over the block's rows
__global__ void __launch_bounds__(PP_THREADS) column_moments_kernel(const float* __restrict__ x, int64_t ld, int64_t n,
                                                                    int D, const float* __restrict__ shift,
                                                                    int64_t rows_per_block, bool vec,
                                                                    double* __restrict__ partial) {
  __shared__ double red[PP_ROWS][PP_COLS + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * PP_COLS + tx * 4;
  const int valid = max(0, min(4, D - c0));
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  double sh[4], s[4], ss[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sh[j] = (shift != nullptr && j < valid) ? static_cast<double>(__ldg(shift + c0 + j)) : 0.0;
    s[j] = 0.0;
    ss[j] = 0.0;
  }
  if (valid > 0) {
    int64_t r = r_begin + ty;
    // four rows in flight per thread
    for (; r + 3 * PP_ROWS < r_end; r += 4 * PP_ROWS) {
      Quad q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = load_quad(x + (r + u * PP_ROWS) * ld, c0, valid, vec);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double d = static_cast<double>(q[u].v[j]) - sh[j];
          s[j] += d;
          ss[j] = fma(d, d, ss[j]);
        }
    }
    for (; r < r_end; r += PP_ROWS) {
      const Quad q = load_quad(x + r * ld, c0, valid, vec);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double d = static_cast<double>(q.v[j]) - sh[j];
        s[j] += d;
        ss[j] = fma(d, d, ss[j]);
      }
    }
  }
  // ordered reduction over the 8 row lanes, sums first, then squares
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = pass == 0 ? s[j] : ss[j];
    __syncthreads();
    if (threadIdx.x < PP_COLS) {
      const int c = blockIdx.x * PP_COLS + threadIdx.x;
      if (c < D) {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < PP_ROWS; ++l) a += red[l][threadIdx.x];
        partial[(static_cast<int64_t>(blockIdx.y) * 2 + pass) * D + c] = a;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) column_moments_finish_kernel(const double* __restrict__ partial, int blocks_y, int D,
                                                                    double* __restrict__ sum, double* __restrict__ sumsq) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  double a = 0.0, b = 0.0;
  for (int by = 0; by < blocks_y; ++by) {
    a += partial[(static_cast<int64_t>(by) * 2 + 0) * D + c];
    b += partial[(static_cast<int64_t>(by) * 2 + 1) * D + c];
  }
  sum[c] = a;
  sumsq[c] = b;
}

__global__ void __launch_bounds__(PP_THREADS) standardize_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int D,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ scale,
                                                                 int64_t rows_per_block, bool vec,
                                                                 float* __restrict__ out, int64_t ld_out) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * PP_COLS + tx * 4;
  const int valid = max(0, min(4, D - c0));
  if (valid == 0) return;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  float mu[4], sc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mu[j] = j < valid ? __ldg(mean + c0 + j) : 0.f;
    sc[j] = j < valid ? __ldg(scale + c0 + j) : 1.f;
  }
  auto emit = [&](const Quad& q, int64_t r) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = __fdiv_rn(__fsub_rn(q.v[j], mu[j]), sc[j]);      // X -= mean_; X /= scale_
    }
    float* dst = out + r * ld_out + c0;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < valid) dst[j] = o[j];
    }
  };
  int64_t r = r_begin + ty;
  for (; r + 3 * PP_ROWS < r_end; r += 4 * PP_ROWS) {
    Quad q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = load_quad(x + (r + u * PP_ROWS) * ld, c0, valid, vec);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(q[u], r + u * PP_ROWS);
  }
  for (; r < r_end; r += PP_ROWS) emit(load_quad(x + r * ld, c0, valid, vec), r);
}

// ---- float64 Gram matrix of the standardised rows -------------------------------------------------------------
constexpr int G_T = 64;            // tile edge (columns of Z)
constexpr int G_R = 32;            // rows per shared-memory stage

// partial[split][a][b] for the upper-triangle tiles (bi <= bj); zsum_partial[split][a] from the diagonal tiles
__global__ void __launch_bounds__(256) gram_f64_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int D,
                                                       const float* __restrict__ mean, const float* __restrict__ scale,
                                                       int64_t rows_per_split, int tiles_1d,
                                                       double* __restrict__ partial, double* __restrict__ zsum_partial) {
  __shared__ __align__(16) double As[G_R][G_T + 2];
  __shared__ __align__(16) double Bs[G_R][G_T + 2];
  // linear upper-triangle tile index -> (bi, bj), bi <= bj
  int t = blockIdx.x, bi = 0;
  while (t >= tiles_1d - bi) { t -= tiles_1d - bi; ++bi; }
  const int bj = bi + t;
  const bool diag = bi == bj;
  const int split = blockIdx.y;
  const int64_t r_begin = static_cast<int64_t>(split) * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double colsum[4] = {0.0, 0.0, 0.0, 0.0};
  // this thread's load slots: 8 elements of each operand stage (32 rows x 64 columns / 256 threads)
  const int lc = threadIdx.x & 63, lr0 = threadIdx.x >> 6;          // column, first row (rows lr0, lr0+4, ...)
  const int ca = bi * G_T + lc, cb = bj * G_T + lc;
  const float ma = (mean != nullptr && ca < D) ? __ldg(mean + ca) : 0.f;
  const float sa = (scale != nullptr && ca < D) ? __ldg(scale + ca) : 1.f;
  const float mb = (mean != nullptr && cb < D) ? __ldg(mean + cb) : 0.f;
  const float sb = (scale != nullptr && cb < D) ? __ldg(scale + cb) : 1.f;
  const bool std_on = mean != nullptr;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += G_R) {
#pragma unroll
    for (int u = 0; u < G_R / 4; ++u) {
      const int rr = lr0 + 4 * u;
      const int64_t r = r0 + rr;
      float va = 0.f, vb = 0.f;
      if (r < r_end) {
        if (ca < D) { va = __ldg(x + r * ld + ca); if (std_on) va = __fdiv_rn(__fsub_rn(va, ma), sa); }
        if (!diag && cb < D) { vb = __ldg(x + r * ld + cb); if (std_on) vb = __fdiv_rn(__fsub_rn(vb, mb), sb); }
      }
      As[rr][lc] = static_cast<double>(va);
      if (!diag) Bs[rr][lc] = static_cast<double>(vb);
    }
    __syncthreads();
    const double (*Bp)[G_T + 2] = diag ? As : Bs;
#pragma unroll 8
    for (int rr = 0; rr < G_R; ++rr) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[rr][ty * 4 + i]; b[i] = Bp[rr][tx + 16 * i]; }     // b: conflict-free
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      if (diag && tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) colsum[i] += a[i];
      }
    }
    __syncthreads();
  }
  double* out = partial + static_cast<int64_t>(split) * D * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = bi * G_T + ty * 4 + i;
    if (a >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = bj * G_T + tx + 16 * j;
      if (b < D) out[static_cast<int64_t>(a) * D + b] = acc[i][j];
    }
    if (diag && tx == 0) zsum_partial[static_cast<int64_t>(split) * D + a] = colsum[i];
  }
}

// ordered sum over the splits; the lower triangle mirrors the upper one
__global__ void __launch_bounds__(256) gram_finish_kernel(const double* __restrict__ partial,
                                                          const double* __restrict__ zsum_partial, int splits, int D,
                                                          double* __restrict__ gram, double* __restrict__ zsum) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < static_cast<int64_t>(D) * D) {
    const int a = static_cast<int>(e / D), b = static_cast<int>(e % D);
    const int ta = a / G_T, tb = b / G_T;
    const int64_t src = ta <= tb ? static_cast<int64_t>(a) * D + b : static_cast<int64_t>(b) * D + a;
    double v = 0.0;
    for (int s = 0; s < splits; ++s) v += partial[static_cast<int64_t>(s) * D * D + src];
    gram[e] = v;
  }
  if (e < D) {
    double v = 0.0;
    for (int s = 0; s < splits; ++s) v += zsum_partial[static_cast<int64_t>(s) * D + e];
    zsum[e] = v;
  }
}

static void gram_partition(int64_t n, int D, int* tiles_1d, int* splits, int64_t* rows_per_split) {
  const int t1 = (D + G_T - 1) / G_T;
  const int64_t tiles = static_cast<int64_t>(t1) * (t1 + 1) / 2;
  int64_t s = (3 * 148 + tiles - 1) / tiles;                  // about three waves of blocks
  const int64_t max_s = (n + 4 * G_R - 1) / (4 * G_R);        // at least 128 rows per split
  s = max(static_cast<int64_t>(1), min(min(s, max_s), static_cast<int64_t>(64)));
  int64_t rps = (n + s - 1) / s;
  rps = (rps + G_R - 1) / G_R * G_R;
  s = max(static_cast<int64_t>(1), (n + rps - 1) / rps);
  *tiles_1d = t1; *splits = static_cast<int>(s); *rows_per_split = rps;
}

// fixed partition of the rows: about two waves of blocks, at least 64 rows per block
static void pp_partition(int64_t n, int D, int* blocks_x, int* blocks_y, int64_t* rows_per_block) {
  const int bx = (D + PP_COLS - 1) / PP_COLS;
  const int64_t target = max(static_cast<int64_t>(1), static_cast<int64_t>(2 * 8 * 148) / bx);
  int64_t by = min(target, (n + 63) / 64);
  by = max(static_cast<int64_t>(1), min(by, static_cast<int64_t>(65535)));
  int64_t rpb = (n + by - 1) / by;
  rpb = (rpb + PP_ROWS - 1) / PP_ROWS * PP_ROWS;
  by = (n + rpb - 1) / rpb;
  *blocks_x = bx;
  *blocks_y = static_cast<int>(max(static_cast<int64_t>(1), by));
  *rows_per_block = rpb;
}

static bool pp_vec_ok(const void* p, int64_t ld, int D) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0 && (D & 3) == 0;
}

}  // namespace emr2a

using namespace emr2a;

extern "C" size_t emr2a_column_moments_workspace_bytes(int64_t n, int D) {
  if (n <= 0 || D <= 0) return 16;
  int bx, by;
  int64_t rpb;
  pp_partition(n, D, &bx, &by, &rpb);
  return static_cast<size_t>(by) * 2 * D * sizeof(double);
}

extern "C" int emr2a_column_moments(const float* x, int64_t ld, int64_t n, int D, const float* shift,
                                    double* sum, double* sumsq, void* workspace, size_t ws_bytes, void* stream) {
  if (!x || !sum || !sumsq || n < 0 || D <= 0 || ld < D) return fail(EMR2A_ERR_INVALID, "column_moments: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    EMR2A_CUDA_TRY(cudaMemsetAsync(sum, 0, sizeof(double) * D, st));
    EMR2A_CUDA_TRY(cudaMemsetAsync(sumsq, 0, sizeof(double) * D, st));
    return EMR2A_OK;
  }
  if (!workspace || ws_bytes < emr2a_column_moments_workspace_bytes(n, D) ||
      (reinterpret_cast<uintptr_t>(workspace) & 7) != 0)
    return fail(EMR2A_ERR_WORKSPACE, "column_moments: workspace too small or misaligned");
  int bx, by;
  int64_t rpb;
  pp_partition(n, D, &bx, &by, &rpb);
  double* partial = static_cast<double*>(workspace);
  column_moments_kernel<<<dim3(bx, by), PP_THREADS, 0, st>>>(x, ld, n, D, shift, rpb, pp_vec_ok(x, ld, D), partial);
  EMR2A_LAUNCH_CHECK("column_moments_kernel");
  column_moments_finish_kernel<<<(D + 127) / 128, 128, 0, st>>>(partial, by, D, sum, sumsq);
  EMR2A_LAUNCH_CHECK("column_moments_finish_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_standardize(const float* x, int64_t ld, int64_t n, int D, const float* mean,
                                 const float* scale, float* out, int64_t ld_out, void* stream) {
  if (!x || !mean || !scale || !out || n < 0 || D <= 0 || ld < D || ld_out < D)
    return fail(EMR2A_ERR_INVALID, "standardize: bad arguments");
  if (n == 0) return EMR2A_OK;
  int bx, by;
  int64_t rpb;
  pp_partition(n, D, &bx, &by, &rpb);
  const bool vec = pp_vec_ok(x, ld, D) && pp_vec_ok(out, ld_out, D);
  standardize_kernel<<<dim3(bx, by), PP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, n, D, mean, scale, rpb,
                                                                                      vec, out, ld_out);
  EMR2A_LAUNCH_CHECK("standardize_kernel");
  return EMR2A_OK;
}

extern "C" size_t emr2a_gram_f64_workspace_bytes(int64_t n, int D) {
  if (n <= 0 || D <= 0) return 16;
  int t1, splits;
  int64_t rps;
  gram_partition(n, D, &t1, &splits, &rps);
  return sizeof(double) * static_cast<size_t>(splits) * (static_cast<size_t>(D) * D + D);
}

extern "C" int emr2a_gram_f64(const float* x, int64_t ld, int64_t n, int D, const float* mean, const float* scale,
                              double* gram, double* zsum, void* workspace, size_t ws_bytes, void* stream) {
  if (!x || !gram || !zsum || n < 0 || D <= 0 || ld < D || ((mean == nullptr) != (scale == nullptr)))
    return fail(EMR2A_ERR_INVALID, "gram_f64: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    EMR2A_CUDA_TRY(cudaMemsetAsync(gram, 0, sizeof(double) * D * D, st));
    EMR2A_CUDA_TRY(cudaMemsetAsync(zsum, 0, sizeof(double) * D, st));
    return EMR2A_OK;
  }
  if (!workspace || ws_bytes < emr2a_gram_f64_workspace_bytes(n, D) || (reinterpret_cast<uintptr_t>(workspace) & 7) != 0)
    return fail(EMR2A_ERR_WORKSPACE, "gram_f64: workspace too small or misaligned");
  int t1, splits;
  int64_t rps;
  gram_partition(n, D, &t1, &splits, &rps);
  double* partial = static_cast<double*>(workspace);
  double* zpart = partial + static_cast<size_t>(splits) * D * D;
  gram_f64_kernel<<<dim3(t1 * (t1 + 1) / 2, splits), 256, 0, st>>>(x, ld, n, D, mean, scale, rps, t1, partial, zpart);
  EMR2A_LAUNCH_CHECK("gram_f64_kernel");
  const int64_t total = static_cast<int64_t>(D) * D;
  gram_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(partial, zpart, splits, D, gram, zsum);
  EMR2A_LAUNCH_CHECK("gram_finish_kernel");
  return EMR2A_OK;
}
