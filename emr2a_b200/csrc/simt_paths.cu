// CUDA-core (fp32 FMA) paths of libemr2a.so:
//   - emr2a_scores            full score matrix (API surfaces that return every score)
//   - simt_topk_search        EMR2A_PREC_FP32 arm of emr2a_topk_search: exact-order fp32
//                             arithmetic, any shape/alignment; also the on-device checker
//                             of the tcgen05 arm in the parity tests
//   - emr2a_topk_from_scores, emr2a_euclid_scores, emr2a_late_fuse_scores
//
// Every dot product is a single fp32 accumulator advanced in ascending-k order, so a
// score is independent of tile shape, split count and GPU count.
#include "common.cuh"

namespace emr2a {

constexpr int S_BM = 64;    // queries per tile
constexpr int S_BN = 128;   // database rows per tile
constexpr int S_BK = 16;
constexpr int S_THREADS = 256;
constexpr int S_SPAD = S_BN + 1;

struct SimtParams {
  const float* q;
  const float* db;
  int64_t Q, N;
  int D;
  int64_t ldq, lddb;
  // scores mode
  float* out;
  int64_t ld_out;
  // projection mode (emr2a_project): q rows are standardised on load, (q - a_mean) / a_scale in IEEE fp32, and
  // out_bias[c] is subtracted from column c of the result
  const float* a_mean;
  const float* a_scale;
  const float* out_bias;
  // top-k mode
  const uint8_t* q_fold;
  const uint8_t* db_fold;
  int64_t idx_base;
  int K;
  uint64_t* keys_out;       // [splits][Q][K]
  int64_t tiles_per_split;
};

__device__ __forceinline__ void simt_tile_mma(const SimtParams& p, int64_t m0, int64_t n0,
                                              float (*As)[S_BM + 4], float (*Bs)[S_BN + 4],
                                              float acc[4][8]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.D; k0 += S_BK) {
    // A tile: 64 rows x 16 k  -> 1024 values, 4 per thread (k fastest for coalescing)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + it * S_THREADS;
      const int r = e >> 4, k = e & 15;
      const int64_t gr = m0 + r;
      float v = 0.f;
      if (gr < p.Q && k0 + k < p.D) {
        v = __ldg(p.q + gr * p.ldq + k0 + k);
        if (p.a_mean != nullptr) v = __fdiv_rn(__fsub_rn(v, __ldg(p.a_mean + k0 + k)), __ldg(p.a_scale + k0 + k));
      }
      As[k][r] = v;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int e = tid + it * S_THREADS;
      const int r = e >> 4, k = e & 15;
      const int64_t gr = n0 + r;
      float v = 0.f;
      if (gr < p.N && k0 + k < p.D) v = __ldg(p.db + gr * p.lddb + k0 + k);
      Bs[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < S_BK; ++k) {
      float a[4], b[8];
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(S_THREADS) simt_scores_kernel(const SimtParams p) {
  __shared__ __align__(16) float As[S_BK][S_BM + 4];
  __shared__ __align__(16) float Bs[S_BK][S_BN + 4];
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * S_BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * S_BN;
  float acc[4][8];
  simt_tile_mma(p, m0, n0, As, Bs, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = m0 + ty * 4 + i;
    if (r >= p.Q) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (c < p.N) p.out[r * p.ld_out + c] = p.out_bias != nullptr ? __fsub_rn(acc[i][j], __ldg(p.out_bias + c)) : acc[i][j];
    }
  }
}

// Top-K arm: a block owns one query tile and a contiguous range of database tiles; the
// per-query sorted lists live in shared memory across the whole range.
__global__ void __launch_bounds__(S_THREADS) simt_topk_kernel(const SimtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*As)[S_BM + 4] = reinterpret_cast<float (*)[S_BM + 4]>(smem_raw);
  float (*Bs)[S_BN + 4] = reinterpret_cast<float (*)[S_BN + 4]>(smem_raw + sizeof(float) * S_BK * (S_BM + 4));
  float* S = reinterpret_cast<float*>(smem_raw + sizeof(float) * S_BK * (S_BM + 4 + S_BN + 4));
  uint64_t* lists = reinterpret_cast<uint64_t*>(S + S_BM * S_SPAD + 1 /*keep 8B alignment below*/);
  lists = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(lists) + 7) & ~static_cast<uintptr_t>(7));

  const int tid = threadIdx.x;
  const int K = p.K;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * S_BM;
  const int64_t n_tiles = (p.N + S_BN - 1) / S_BN;
  const int64_t t_begin = static_cast<int64_t>(blockIdx.y) * p.tiles_per_split;
  int64_t t_end = t_begin + p.tiles_per_split;
  if (t_end > n_tiles) t_end = n_tiles;

  for (int e = tid; e < S_BM * K; e += S_THREADS) lists[e] = 0ull;
  uint64_t kth = 0ull;                       // current K-th best key of this thread's query
  const int64_t my_q = m0 + tid;
  const int my_fold = (tid < S_BM && my_q < p.Q && p.q_fold) ? p.q_fold[my_q] : -1;
  __syncthreads();

  const int ty = tid >> 4, tx = tid & 15;
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t n0 = t * S_BN;
    float acc[4][8];
    simt_tile_mma(p, m0, n0, As, Bs, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        S[(ty * 4 + i) * S_SPAD + c] = acc[i][j];
      }
    __syncthreads();
    if (tid < S_BM && my_q < p.Q) {
      uint64_t* mine = lists + static_cast<int64_t>(tid) * K;
      const int64_t lim = (p.N - n0 < S_BN) ? (p.N - n0) : S_BN;
      for (int c = 0; c < lim; ++c) {
        if (my_fold >= 0 && p.db_fold[n0 + c] == my_fold) continue;
        const uint64_t key = pack_key(S[tid * S_SPAD + c], static_cast<uint32_t>(n0 + c + p.idx_base));
        if (key > kth) {
          int pos = K - 1;
          while (pos > 0 && mine[pos - 1] < key) { mine[pos] = mine[pos - 1]; --pos; }
          mine[pos] = key;
          kth = mine[K - 1];
        }
      }
    }
    __syncthreads();
  }
  // write the partial lists
  for (int e = tid; e < S_BM * K; e += S_THREADS) {
    const int r = e / K, j = e - r * K;
    if (m0 + r < p.Q)
      p.keys_out[(static_cast<int64_t>(blockIdx.y) * p.Q + m0 + r) * K + j] = lists[e];
  }
}

static size_t simt_topk_smem(int K) {
  return sizeof(float) * S_BK * (S_BM + 4 + S_BN + 4) + sizeof(float) * (S_BM * S_SPAD + 1) + 8 +
         sizeof(uint64_t) * S_BM * K;
}

int simt_pick_splits(int64_t Q, int64_t N) {
  const int64_t m_tiles = (Q + S_BM - 1) / S_BM;
  const int64_t n_tiles = (N + S_BN - 1) / S_BN;
  int64_t want = (2LL * sm_count() + m_tiles - 1) / m_tiles;   // ~2 CTAs per SM
  if (want < 1) want = 1;
  if (want > n_tiles) want = n_tiles;
  if (want > 64) want = 64;
  return static_cast<int>(want < 1 ? 1 : want);
}

size_t simt_topk_workspace_bytes(int64_t Q, int64_t N, int K) {
  const int splits = simt_pick_splits(Q, N);
  return splits > 1 ? sizeof(uint64_t) * static_cast<size_t>(splits) * Q * K : 0;
}

int simt_topk_search(const float* q, const float* db, int64_t Q, int64_t N, int D, int64_t ldq, int64_t lddb,
                     const uint8_t* q_fold, const uint8_t* db_fold, int64_t idx_base, int K,
                     uint64_t* out_keys, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (K > 128) return fail(EMR2A_ERR_UNSUPPORTED, "topk_search(fp32): K=%d > 128", K);
  const int splits = simt_pick_splits(Q, N);
  const size_t need = simt_topk_workspace_bytes(Q, N, K);
  if (need > ws_bytes || (need && !workspace))
    return fail(EMR2A_ERR_WORKSPACE, "topk_search(fp32): workspace %zu < %zu", ws_bytes, need);
  const int64_t n_tiles = (N + S_BN - 1) / S_BN;
  SimtParams p{};
  p.q = q; p.db = db; p.Q = Q; p.N = N; p.D = D; p.ldq = ldq; p.lddb = lddb;
  p.q_fold = q_fold; p.db_fold = db_fold; p.idx_base = idx_base; p.K = K;
  p.keys_out = splits > 1 ? static_cast<uint64_t*>(workspace) : out_keys;
  p.tiles_per_split = (n_tiles + splits - 1) / splits;
  const size_t smem = simt_topk_smem(K);
  EMR2A_CUDA_TRY(cudaFuncSetAttribute(simt_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
  dim3 grid(static_cast<unsigned>((Q + S_BM - 1) / S_BM), static_cast<unsigned>(splits));
  simt_topk_kernel<<<grid, S_THREADS, smem, st>>>(p);
  EMR2A_LAUNCH_CHECK("simt_topk_kernel");
  if (splits > 1)
    return emr2a_topk_merge(static_cast<const uint64_t*>(workspace), splits, Q, K, Q * K, K, K, out_keys, st);
  return EMR2A_OK;
}

// ---- Top-K of a given score matrix: K rounds of block-wide arg-max ---------------------
__global__ void __launch_bounds__(256) topk_from_scores_kernel(const float* __restrict__ scores, int64_t N,
                                                               int64_t ld, int K, uint64_t* __restrict__ out) {
  __shared__ uint64_t red[8];
  __shared__ uint64_t winner;
  const float* row = scores + static_cast<int64_t>(blockIdx.x) * ld;
  uint64_t bound = ~0ull;                       // keys are unique: take the best key strictly below the last winner
  for (int r = 0; r < K; ++r) {
    uint64_t best = 0ull;
    for (int64_t c = threadIdx.x; c < N; c += blockDim.x) {
      const uint64_t key = pack_key(row[c], static_cast<uint32_t>(c));
      if (key < bound && key > best) best = key;
    }
    best = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint64_t v = threadIdx.x < 8 ? red[threadIdx.x] : 0ull;
      v = warp_max_u64(v);
      if (threadIdx.x == 0) { winner = v; out[static_cast<int64_t>(blockIdx.x) * K + r] = v; }
    }
    __syncthreads();
    bound = winner;
    if (bound == 0ull) {                        // fewer than K admissible entries: rest stay empty
      for (int r2 = r + 1 + threadIdx.x; r2 < K; r2 += blockDim.x) out[static_cast<int64_t>(blockIdx.x) * K + r2] = 0ull;
      break;
    }
  }
}

// ---- euclidean similarity ------------------------------------------------------------
__global__ void __launch_bounds__(256) euclid_dist_kernel(const float* __restrict__ q, const float* __restrict__ db,
                                                          int64_t N, int D, int64_t lddb, float* __restrict__ out,
                                                          unsigned int* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  float ss = 0.f;
  for (int e = lane; e < D; e += 32) {
    const float d = __fsub_rn(__ldg(db + row * lddb + e), __ldg(q + e));
    ss = fmaf(d, d, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) {
    const float dist = __fsqrt_rn(ss);
    out[row] = dist;
    atomicMax(max_bits, __float_as_uint(dist));     // dist >= 0: uint order == float order
  }
}
__global__ void __launch_bounds__(256) euclid_finish_kernel(int64_t N, float* __restrict__ out,
                                                            const unsigned int* __restrict__ max_bits) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float mx = __uint_as_float(*max_bits);
  const float d = out[i];
  out[i] = mx > 0.f ? __fsub_rn(1.0f, __fdiv_rn(d, mx)) : __fsub_rn(1.0f, d);
}

// ---- late fusion of score rows ---------------------------------------------------------
__device__ double block_sum_d(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  return t;
}
__device__ float block_minmax_f(float v, bool want_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_max ? fmaxf(v, t) : fminf(v, t);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) t = want_max ? fmaxf(t, red[w]) : fminf(t, red[w]);
  return t;
}

// shift/denominator of normalize_scores for one row (retrieval/fusion.py:31-42)
__device__ void row_norm_consts(const float* row, int64_t N, int mode, float& shift, float& denom,
                                double* red_d, float* red_f) {
  shift = 0.f; denom = 1.f;
  if (mode == EMR2A_SCORE_ZSCORE) {
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < N; c += blockDim.x) s += static_cast<double>(row[c]);
    const float mean = static_cast<float>(block_sum_d(s, red_d) / static_cast<double>(N));
    double v = 0.0;
    for (int64_t c = threadIdx.x; c < N; c += blockDim.x) {
      const double d = static_cast<double>(row[c]) - static_cast<double>(mean);
      v += d * d;
    }
    const float sd = static_cast<float>(sqrt(block_sum_d(v, red_d) / static_cast<double>(N)));
    shift = mean;
    denom = static_cast<float>(static_cast<double>(sd) + 1e-8);
  } else if (mode == EMR2A_SCORE_MINMAX) {
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t c = threadIdx.x; c < N; c += blockDim.x) { const float x = row[c]; lo = fminf(lo, x); hi = fmaxf(hi, x); }
    lo = block_minmax_f(lo, false, red_f);
    hi = block_minmax_f(hi, true, red_f);
    shift = lo;
    denom = static_cast<float>(static_cast<double>(hi) - static_cast<double>(lo) + 1e-8);
  }
}

__global__ void __launch_bounds__(256) late_fuse_kernel(const float* __restrict__ ts, const float* __restrict__ is_,
                                                        int64_t N, int64_t ld, float w, float omw, int mode,
                                                        float* __restrict__ out, int64_t ld_out) {
  __shared__ double red_d[8];
  __shared__ float red_f[8];
  const float* trow = ts + static_cast<int64_t>(blockIdx.x) * ld;
  const float* irow = is_ + static_cast<int64_t>(blockIdx.x) * ld;
  float* orow = out + static_cast<int64_t>(blockIdx.x) * ld_out;
  float st, dt, si, di;
  row_norm_consts(trow, N, mode, st, dt, red_d, red_f);
  row_norm_consts(irow, N, mode, si, di, red_d, red_f);
  for (int64_t c = threadIdx.x; c < N; c += blockDim.x) {
    float a = trow[c], b = irow[c];
    if (mode != EMR2A_SCORE_NONE) {
      a = __fdiv_rn(__fsub_rn(a, st), dt);
      b = __fdiv_rn(__fsub_rn(b, si), di);
    }
    orow[c] = __fadd_rn(__fmul_rn(w, a), __fmul_rn(omw, b));     // no FMA contraction: numpy rounds each product
  }
}

// ---- mean over the slices of each patient (ingest) ---------------------------------------
// One block per patient; threads stride over the feature dimension (coalesced), slices are added in
// order in fp32 and divided by the count -- numpy's `arr.mean(axis=0)` arithmetic.  HBM-bound:
// total_slices * D * 4 bytes read + n * D * 4 written.
__global__ void __launch_bounds__(128) segment_mean_kernel(const float* __restrict__ x, int64_t ld,
                                                           const int64_t* __restrict__ offsets, int D,
                                                           float* __restrict__ out, int64_t ld_out) {
  const int64_t seg = blockIdx.x;
  const int64_t r0 = offsets[seg], r1 = offsets[seg + 1];
  const float cnt = static_cast<float>(r1 - r0);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc = __fadd_rn(acc, __ldg(x + r * ld + c));
    out[seg * ld_out + c] = r1 > r0 ? __fdiv_rn(acc, cnt) : 0.f;
  }
}

}  // namespace emr2a

using namespace emr2a;

extern "C" int emr2a_segment_mean(const float* x, int64_t ld, const int64_t* offsets, int64_t n_segments, int D,
                                  float* out, int64_t ld_out, void* stream) {
  if (!x || !offsets || !out || n_segments < 0 || D <= 0 || ld < D || ld_out < D)
    return fail(EMR2A_ERR_INVALID, "segment_mean: bad arguments");
  if (n_segments == 0) return EMR2A_OK;
  segment_mean_kernel<<<static_cast<unsigned>(n_segments), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, offsets, D, out, ld_out);
  EMR2A_LAUNCH_CHECK("segment_mean_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_scores(const float* q, const float* db, int64_t Q, int64_t N, int D, int64_t ldq,
                            int64_t lddb, float* out, int64_t ld_out, void* stream) {
  if (Q < 0 || N < 0 || D <= 0 || !q || !db || !out) return fail(EMR2A_ERR_INVALID, "scores: bad arguments");
  if (ldq < D || lddb < D || ld_out < N) return fail(EMR2A_ERR_INVALID, "scores: leading dimension too small");
  if (Q == 0 || N == 0) return EMR2A_OK;
  SimtParams p{};
  p.q = q; p.db = db; p.Q = Q; p.N = N; p.D = D; p.ldq = ldq; p.lddb = lddb; p.out = out; p.ld_out = ld_out;
  const int64_t gy = (Q + S_BM - 1) / S_BM;
  if (gy > 65535) return fail(EMR2A_ERR_UNSUPPORTED, "scores: Q=%lld too large for one call", (long long)Q);
  dim3 grid(static_cast<unsigned>((N + S_BN - 1) / S_BN), static_cast<unsigned>(gy));
  simt_scores_kernel<<<grid, S_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  EMR2A_LAUNCH_CHECK("simt_scores_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_project(const float* x, int64_t ld, int64_t n, int D, const float* mean, const float* scale,
                             const float* w, int64_t ldw, int P, const float* bias, float* out, int64_t ld_out,
                             void* stream) {
  if (n < 0 || D <= 0 || P <= 0 || !x || !w || !out || ld < D || ldw < D || ld_out < P || ((mean == nullptr) != (scale == nullptr)))
    return fail(EMR2A_ERR_INVALID, "project: bad arguments");
  if (n == 0) return EMR2A_OK;
  SimtParams p{};
  p.q = x; p.db = w; p.Q = n; p.N = P; p.D = D; p.ldq = ld; p.lddb = ldw; p.out = out; p.ld_out = ld_out;
  p.a_mean = mean; p.a_scale = scale; p.out_bias = bias;
  const int64_t gy = (n + S_BM - 1) / S_BM;
  for (int64_t y0 = 0; y0 < gy; y0 += 65535) {                 // the row-tile index lives in gridDim.y
    const int64_t rows0 = y0 * S_BM;
    SimtParams c = p;
    c.q = x + rows0 * ld; c.out = out + rows0 * ld_out; c.Q = n - rows0 < 65535LL * S_BM ? n - rows0 : 65535LL * S_BM;
    dim3 grid(static_cast<unsigned>((P + S_BN - 1) / S_BN), static_cast<unsigned>((c.Q + S_BM - 1) / S_BM));
    simt_scores_kernel<<<grid, S_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(c);
    EMR2A_LAUNCH_CHECK("simt_scores_kernel(project)");
  }
  return EMR2A_OK;
}

extern "C" int emr2a_topk_from_scores(const float* scores, int64_t Q, int64_t N, int64_t ld, int K,
                                      uint64_t* out_keys, void* stream) {
  if (Q < 0 || N < 0 || K <= 0 || !scores || !out_keys || ld < N) return fail(EMR2A_ERR_INVALID, "topk_from_scores: bad arguments");
  if (Q == 0) return EMR2A_OK;
  topk_from_scores_kernel<<<static_cast<unsigned>(Q), 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, N, ld, K, out_keys);
  EMR2A_LAUNCH_CHECK("topk_from_scores_kernel");
  return EMR2A_OK;
}

extern "C" size_t emr2a_euclid_workspace_bytes(int64_t) { return 16; }

extern "C" int emr2a_euclid_scores(const float* q, const float* db, int64_t N, int D, int64_t lddb, float* out,
                                   void* workspace, size_t ws_bytes, void* stream) {
  if (N < 0 || D <= 0 || !q || !db || !out || lddb < D) return fail(EMR2A_ERR_INVALID, "euclid_scores: bad arguments");
  if (!workspace || ws_bytes < 16) return fail(EMR2A_ERR_WORKSPACE, "euclid_scores: workspace too small");
  if (N == 0) return EMR2A_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EMR2A_CUDA_TRY(cudaMemsetAsync(workspace, 0, 16, st));
  const int64_t blocks = (N * 32 + 255) / 256;
  euclid_dist_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(q, db, N, D, lddb, out, static_cast<unsigned int*>(workspace));
  EMR2A_LAUNCH_CHECK("euclid_dist_kernel");
  euclid_finish_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, st>>>(N, out, static_cast<unsigned int*>(workspace));
  EMR2A_LAUNCH_CHECK("euclid_finish_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_late_fuse_scores(const float* text_scores, const float* image_scores, int64_t Q, int64_t N,
                                      int64_t ld, float w_text, float one_minus_w, int mode, float* out,
                                      int64_t ld_out, void* stream) {
  if (Q < 0 || N <= 0 || !text_scores || !image_scores || !out || ld < N || ld_out < N)
    return fail(EMR2A_ERR_INVALID, "late_fuse_scores: bad arguments");
  if (mode < EMR2A_SCORE_NONE || mode > EMR2A_SCORE_MINMAX) return fail(EMR2A_ERR_INVALID, "late_fuse_scores: unknown mode %d", mode);
  if (Q == 0) return EMR2A_OK;
  late_fuse_kernel<<<static_cast<unsigned>(Q), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      text_scores, image_scores, N, ld, w_text, one_minus_w, mode, out, ld_out);
  EMR2A_LAUNCH_CHECK("late_fuse_kernel");
  return EMR2A_OK;
}
