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
// Both read every element once with 128-bit loads (a warp covers 512 contiguous bytes of a row) and keep
// the per-column constants in registers.  The covariance and projection GEMMs of the PCA are plain library
// GEMMs issued by the host layer (emr2a_b200/preprocess.py).
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
