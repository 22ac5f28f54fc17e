// K3: merge of partial Top-K lists; K4: label gather + majority/weighted vote + metrics counters.
//
// Both are HBM-bound streams over packed keys:
//   K3 algorithmic bytes/query = parts * K_in * 8 (read) + K_out * 8 (write)
//   K4 algorithmic bytes/query = K * 8 (keys) + K * 4 (label gather) + 4 (true label)
//                                + requested per-query outputs
#include "common.cuh"

namespace emr2a {

// ---- K3 ---------------------------------------------------------------------------
// One warp per query.  Keys of all parts are spread over the lanes (E per lane, in
// registers); K_out rounds of warp arg-max pop the winners.  Keys are unique (an index
// occurs in exactly one part), 0 = empty.
template <int E>
__global__ void __launch_bounds__(256) topk_merge_reg_kernel(const uint64_t* __restrict__ in, int parts, int64_t Q,
                                                             int K_in, int64_t part_stride, int64_t q_stride,
                                                             int K_out, uint64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const int total = parts * K_in;
  uint64_t k[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int t = lane + 32 * e;
    k[e] = 0ull;
    if (t < total) {
      const int p = t / K_in, j = t - p * K_in;
      k[e] = in[p * part_stride + q * q_stride + j];
    }
  }
  uint64_t mine = 0ull;            // lane j keeps output j (K_out <= 32 per pass)
  for (int r0 = 0; r0 < K_out; r0 += 32) {
    for (int r = r0; r < K_out && r < r0 + 32; ++r) {
      uint64_t best = k[0];
#pragma unroll
      for (int e = 1; e < E; ++e) best = k[e] > best ? k[e] : best;
      best = warp_max_u64(best);
      if (best != 0ull) {
#pragma unroll
        for (int e = 0; e < E; ++e) if (k[e] == best) k[e] = 0ull;
      }
      if (lane == r - r0) mine = best;
    }
    if (r0 + lane < K_out) out[q * K_out + r0 + lane] = mine;
  }
}

// P-way merge of SORTED lists (parts <= 32, K_in <= 32): lane p owns list p, staged in shared memory
// ([slot][lane] layout: conflict-free), and exposes its head; each round is one warp arg-max over the
// heads, the winning lane advances.  ~40 instructions per output key instead of ~150 for the
// register kernel above, which re-scans every key of the query each round.
constexpr int PW_WARPS = 8;
__global__ void __launch_bounds__(PW_WARPS * 32) topk_merge_pway_kernel(const uint64_t* __restrict__ in, int parts, int64_t Q,
                                                                        int K_in, int64_t part_stride, int64_t q_stride,
                                                                        int K_out, uint64_t* __restrict__ out) {
  extern __shared__ uint64_t pw_keys[];                       // [PW_WARPS][K_in][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * PW_WARPS + warp;
  if (q >= Q) return;
  uint64_t* mine_list = pw_keys + static_cast<size_t>(warp) * K_in * 32;
  if (lane < parts) {
    const uint64_t* src = in + lane * part_stride + q * q_stride;
    for (int j = 0; j < K_in; ++j) mine_list[j * 32 + lane] = src[j];
  }
  __syncwarp();
  int ptr = 0;
  uint64_t head = lane < parts ? mine_list[lane] : 0ull;
  uint64_t keep = 0ull;
  for (int r = 0; r < K_out; ++r) {
    const uint64_t best = warp_max_u64(head);
    if ((r & 31) == lane) keep = best;
    if (best != 0ull && head == best) {                       // keys are unique: exactly one lane advances
      ++ptr;
      head = ptr < K_in ? mine_list[ptr * 32 + lane] : 0ull;
    }
    if ((r & 31) == 31 || r == K_out - 1) {
      const int base = r & ~31;
      if (base + lane <= r) out[q * K_out + base + lane] = keep;
    }
  }
}

// Generic fallback (parts * K_in > 256): re-scan from memory each round.
__global__ void __launch_bounds__(256) topk_merge_scan_kernel(const uint64_t* __restrict__ in, int parts, int64_t Q,
                                                              int K_in, int64_t part_stride, int64_t q_stride,
                                                              int K_out, uint64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const int total = parts * K_in;
  uint64_t bound = ~0ull;
  for (int r = 0; r < K_out; ++r) {
    uint64_t best = 0ull;
    if (bound != 0ull) {
      for (int t = lane; t < total; t += 32) {
        const int p = t / K_in, j = t - p * K_in;
        const uint64_t key = in[p * part_stride + q * q_stride + j];
        if (key < bound && key > best) best = key;
      }
      best = warp_max_u64(best);
    }
    bound = best;
    if (lane == 0) out[q * K_out + r] = best;
  }
}

// ---- K4 ---------------------------------------------------------------------------
struct VoteParams {
  const uint64_t* keys;
  int64_t Q;
  int K;
  const int32_t* db_labels;
  int64_t label_base;
  const int32_t* q_labels;
  const uint8_t* q_group;
  int n_groups, C;
  int k_list[16];
  int nk;
  int wacc_f32;
  int64_t* top_idx;
  float* top_scores;
  int32_t* top_labels;
  int32_t* pred_top1;
  int32_t* pred_vote;
  int32_t* pred_weighted;
  unsigned long long* hit_counts;
  unsigned long long* vote_counts;
  unsigned long long* confusion;
  unsigned long long* group_sizes;
  int smem_small;     // hits/votes/sizes counters staged in shared memory
  int smem_conf;      // confusion matrices staged in shared memory
};

template <int KMAX>
__global__ void __launch_bounds__(128) vote_metrics_kernel(const VoteParams p) {
  extern __shared__ unsigned int sm_cnt[];
  // layout: [n_groups*(nk+4)] small counters, then [n_groups*2*C*C] confusion
  const int n_small = p.smem_small ? p.n_groups * (p.nk + 4) : 0;
  const int n_conf = p.smem_conf ? p.n_groups * 2 * p.C * p.C : 0;
  for (int i = threadIdx.x; i < n_small + n_conf; i += blockDim.x) sm_cnt[i] = 0u;
  __syncthreads();
  unsigned int* sm_small = sm_cnt;
  unsigned int* sm_conf = sm_cnt + n_small;

  const int K = p.K;
  constexpr int UNR = KMAX <= 16 ? KMAX : 1;     // small K: lists stay in registers
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.Q;
       q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int lab[KMAX];
    float sc[KMAX];
    int valid = 0;
#pragma unroll(UNR)
    for (int j = 0; j < KMAX; ++j) {
      lab[j] = -1; sc[j] = 0.f;
      if (j < K) {
        const uint64_t key = p.keys[q * K + j];
        int64_t idx = -1;
        if (key != 0ull) {
          idx = static_cast<int64_t>(key_index(key));
          sc[j] = key_score(key);
          lab[j] = __ldg(p.db_labels + (idx - p.label_base));
          valid = j + 1;
        }
        if (p.top_idx) p.top_idx[q * K + j] = idx;
        if (p.top_scores) p.top_scores[q * K + j] = sc[j];
        if (p.top_labels) p.top_labels[q * K + j] = lab[j];
      }
    }
    const int truth = p.q_labels[q];
    const int g = p.q_group ? p.q_group[q] : 0;
    int top1 = -1, maj = -1, wv = -1;
    if (valid > 0) {
      top1 = lab[0];
      int best_n = 0;
      double best_s = 0.0;
      bool have_s = false;
#pragma unroll(UNR)
      for (int i = 0; i < KMAX; ++i) {
        if (i < valid) {
          int n = 0;
          double s64 = 0.0;
          float s32 = 0.f;
          bool first = true;
#pragma unroll(UNR)
          for (int j = 0; j < KMAX; ++j) {
            if (j < valid && lab[j] == lab[i]) {
              if (j < i) first = false;
              ++n;
              s64 += static_cast<double>(sc[j]);
              s32 = __fadd_rn(s32, sc[j]);
            }
          }
          if (first) {        // evaluate each label once, at its first (best-ranked) occurrence
            if (n > best_n) { best_n = n; maj = lab[i]; }
            const double s = p.wacc_f32 ? static_cast<double>(s32) : s64;
            if (!have_s || s > best_s) { best_s = s; wv = lab[i]; have_s = true; }
          }
        }
      }
    }
    if (p.pred_top1) p.pred_top1[q] = top1;
    if (p.pred_vote) p.pred_vote[q] = maj;
    if (p.pred_weighted) p.pred_weighted[q] = wv;

    // counters
    const int small_base = g * (p.nk + 4);
    for (int t = 0; t < p.nk; ++t) {
      const int kk = p.k_list[t] < valid ? p.k_list[t] : valid;
      bool hit = false;
#pragma unroll(UNR)
      for (int j = 0; j < KMAX; ++j) if (j < kk && lab[j] == truth) hit = true;
      if (hit) {
        if (p.smem_small) atomicAdd(&sm_small[small_base + t], 1u);
        else atomicAdd(&p.hit_counts[g * p.nk + t], 1ull);
      }
    }
    const bool c1 = top1 == truth && valid > 0, c2 = maj == truth && valid > 0, c3 = wv == truth && valid > 0;
    if (p.smem_small) {
      if (c1) atomicAdd(&sm_small[small_base + p.nk + 0], 1u);
      if (c2) atomicAdd(&sm_small[small_base + p.nk + 1], 1u);
      if (c3) atomicAdd(&sm_small[small_base + p.nk + 2], 1u);
      atomicAdd(&sm_small[small_base + p.nk + 3], 1u);
    } else {
      if (c1) atomicAdd(&p.vote_counts[g * 3 + 0], 1ull);
      if (c2) atomicAdd(&p.vote_counts[g * 3 + 1], 1ull);
      if (c3) atomicAdd(&p.vote_counts[g * 3 + 2], 1ull);
      atomicAdd(&p.group_sizes[g], 1ull);
    }
    if (truth >= 0 && truth < p.C) {
      if (top1 >= 0 && top1 < p.C) {
        const int o = ((g * 2 + 0) * p.C + truth) * p.C + top1;
        if (p.smem_conf) atomicAdd(&sm_conf[o], 1u); else atomicAdd(&p.confusion[o], 1ull);
      }
      if (maj >= 0 && maj < p.C) {
        const int o = ((g * 2 + 1) * p.C + truth) * p.C + maj;
        if (p.smem_conf) atomicAdd(&sm_conf[o], 1u); else atomicAdd(&p.confusion[o], 1ull);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_small; i += blockDim.x) {
    const unsigned int v = sm_small[i];
    if (!v) continue;
    const int g = i / (p.nk + 4), t = i - g * (p.nk + 4);
    if (t < p.nk) atomicAdd(&p.hit_counts[g * p.nk + t], static_cast<unsigned long long>(v));
    else if (t < p.nk + 3) atomicAdd(&p.vote_counts[g * 3 + (t - p.nk)], static_cast<unsigned long long>(v));
    else atomicAdd(&p.group_sizes[g], static_cast<unsigned long long>(v));
  }
  for (int i = threadIdx.x; i < n_conf; i += blockDim.x) {
    const unsigned int v = sm_conf[i];
    if (v) atomicAdd(&p.confusion[i], static_cast<unsigned long long>(v));
  }
}

}  // namespace emr2a

using namespace emr2a;

extern "C" int emr2a_topk_merge(const uint64_t* keys_in, int parts, int64_t Q, int K_in, int64_t part_stride,
                                int64_t q_stride, int K_out, uint64_t* keys_out, void* stream) {
  if (!keys_in || !keys_out || parts <= 0 || Q < 0 || K_in <= 0 || K_out <= 0)
    return fail(EMR2A_ERR_INVALID, "topk_merge: bad arguments");
  if (Q == 0) return EMR2A_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int total = parts * K_in;
  const unsigned blocks = static_cast<unsigned>((Q * 32 + 255) / 256);
  if (parts <= 32 && K_in <= 32 && total > 32) {
    const size_t smem = sizeof(uint64_t) * PW_WARPS * K_in * 32;          // up to 64 KB
    EMR2A_CUDA_TRY(cudaFuncSetAttribute(topk_merge_pway_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    topk_merge_pway_kernel<<<static_cast<unsigned>((Q + PW_WARPS - 1) / PW_WARPS), PW_WARPS * 32, smem, st>>>(
        keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
    EMR2A_LAUNCH_CHECK("topk_merge_pway_kernel");
    return EMR2A_OK;
  }
  if (total <= 32) topk_merge_reg_kernel<1><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 64) topk_merge_reg_kernel<2><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 128) topk_merge_reg_kernel<4><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 256) topk_merge_reg_kernel<8><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 512) topk_merge_reg_kernel<16><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 1024) topk_merge_reg_kernel<32><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else if (total <= 2048) topk_merge_reg_kernel<64><<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  else topk_merge_scan_kernel<<<blocks, 256, 0, st>>>(keys_in, parts, Q, K_in, part_stride, q_stride, K_out, keys_out);
  EMR2A_LAUNCH_CHECK("topk_merge kernel");
  return EMR2A_OK;
}

namespace emr2a {
__global__ void __launch_bounds__(256) keys_map_rows_kernel(uint64_t* __restrict__ keys, int64_t total,
                                                            const uint32_t* __restrict__ row_ids, int64_t n_rows,
                                                            int64_t idx_base) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint64_t k = keys[i];
  if (k == 0ull) return;
  const int64_t local = static_cast<int64_t>(key_index(k)) - idx_base;
  if (local < 0 || local >= n_rows) return;                  // not an index of this shard: left as it is
  keys[i] = (k & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(0xFFFFFFFFu - row_ids[local]);
}
}  // namespace emr2a

extern "C" int emr2a_keys_map_rows(uint64_t* keys, int64_t n_keys, const uint32_t* row_ids, int64_t n_rows,
                                   int64_t idx_base, void* stream) {
  if (!keys || !row_ids || n_keys < 0 || n_rows < 0 || idx_base < 0) return fail(EMR2A_ERR_INVALID, "keys_map_rows: bad arguments");
  if (n_keys == 0) return EMR2A_OK;
  keys_map_rows_kernel<<<static_cast<unsigned>((n_keys + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keys, n_keys, row_ids, n_rows, idx_base);
  EMR2A_LAUNCH_CHECK("keys_map_rows_kernel");
  return EMR2A_OK;
}

extern "C" int emr2a_vote_metrics(const uint64_t* keys, int64_t Q, int K, const int32_t* db_labels,
                                  int64_t label_base, const int32_t* q_labels, const uint8_t* q_group,
                                  int n_groups, int C, const int32_t* k_list, int nk, int wacc_f32,
                                  int64_t* top_idx, float* top_scores, int32_t* top_labels, int32_t* pred_top1,
                                  int32_t* pred_vote, int32_t* pred_weighted, unsigned long long* hit_counts,
                                  unsigned long long* vote_counts, unsigned long long* confusion,
                                  unsigned long long* group_sizes, void* stream) {
  if (!keys || !db_labels || !q_labels || Q < 0 || K <= 0 || n_groups <= 0 || C <= 0)
    return fail(EMR2A_ERR_INVALID, "vote_metrics: bad arguments");
  if (K > 128) return fail(EMR2A_ERR_UNSUPPORTED, "vote_metrics: K=%d > 128", K);
  if (nk < 0 || nk > 16 || (nk > 0 && !k_list)) return fail(EMR2A_ERR_INVALID, "vote_metrics: nk must be in [0,16]");
  if (!hit_counts || !vote_counts || !confusion || !group_sizes)
    return fail(EMR2A_ERR_INVALID, "vote_metrics: counter arrays are required");
  if (Q == 0) return EMR2A_OK;
  VoteParams p{};
  p.keys = keys; p.Q = Q; p.K = K; p.db_labels = db_labels; p.label_base = label_base; p.q_labels = q_labels;
  p.q_group = q_group; p.n_groups = n_groups; p.C = C; p.nk = nk; p.wacc_f32 = wacc_f32;
  for (int i = 0; i < nk; ++i) p.k_list[i] = k_list[i];          // k_list is a HOST array (tiny, copied by value)
  p.top_idx = top_idx; p.top_scores = top_scores; p.top_labels = top_labels;
  p.pred_top1 = pred_top1; p.pred_vote = pred_vote; p.pred_weighted = pred_weighted;
  p.hit_counts = hit_counts; p.vote_counts = vote_counts; p.confusion = confusion; p.group_sizes = group_sizes;
  const int64_t n_small = static_cast<int64_t>(n_groups) * (nk + 4);
  const int64_t n_conf = static_cast<int64_t>(n_groups) * 2 * C * C;
  p.smem_small = n_small <= 2048;
  p.smem_conf = n_conf <= 8192;
  const size_t smem = sizeof(unsigned int) * ((p.smem_small ? n_small : 0) + (p.smem_conf ? n_conf : 0));
  int64_t blocks = (Q + 127) / 128;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K <= 8) vote_metrics_kernel<8><<<static_cast<unsigned>(blocks), 128, smem, st>>>(p);
  else if (K <= 16) vote_metrics_kernel<16><<<static_cast<unsigned>(blocks), 128, smem, st>>>(p);
  else if (K <= 32) vote_metrics_kernel<32><<<static_cast<unsigned>(blocks), 128, smem, st>>>(p);
  else vote_metrics_kernel<128><<<static_cast<unsigned>(blocks), 128, smem, st>>>(p);
  EMR2A_LAUNCH_CHECK("vote_metrics_kernel");
  return EMR2A_OK;
}
