// K2 (tensor-core arm): query x database similarity as a tcgen05/TMEM GEMM fed by TMA, with the
// Top-K selection fused into the epilogue -- the score matrix never reaches HBM.
//
// Mapping (sm_100a, one CTA per SM, 256 threads, warp-specialised):
//   M (TMEM lanes)   = 128 queries      -> each epilogue thread owns ONE query row
//   N (TMEM columns) = 256 database rows per tile, two accumulators (2 x 256 = all 512 columns)
//   K                = 64 bf16 per stage (one 128-byte swizzle row), UMMA_K = 16
//   warp 0  TMA producer   : cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1  MMA issuer     : tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32 in TMEM)
//                            BF16X3: per k-step hi*hi + hi*lo + lo*hi into the same accumulator
//   warp 2  TMEM allocator
//   warps 4-7 epilogue     : tcgen05.ld 32x32b.x32 -> 32 scores of the thread's own query per chunk,
//                            running max against the thread's K-th best (compare-mostly), rare sorted
//                            insertion into a register-resident Top-K list; the epilogue of tile t
//                            overlaps the MMAs of tile t+1 through the second accumulator.
// A CTA walks work units (query tile, database split) statically; each unit ends with the thread
// writing its list as packed keys to the partial-list workspace, merged afterwards by K3.
//
// Roofline: tensor pipe.  Algorithmic FLOPs = 2 * D * Q * N_admissible; BF16X3 issues 3x that.
#include <cuda.h>
#include "common.cuh"

namespace emr2a {

constexpr int T_BM = 128;
constexpr int T_BN = 256;
constexpr int T_BK = 64;
constexpr int T_THREADS = 256;
constexpr uint32_t T_A_BYTES = T_BM * T_BK * 2;   // 16 KB
constexpr uint32_t T_B_BYTES = T_BN * T_BK * 2;   // 32 KB

struct TcParams {
  int64_t Q, N;
  int k_chunks;            // ceil(D / 64)
  int64_t m_tiles, n_tiles;
  int splits;
  int64_t tiles_per_split;
  const uint8_t* q_fold;
  const uint8_t* db_fold;  // padded to n_tiles * 256 bytes
  int fold_sorted;         // both fold vectors are non-decreasing: single-fold tiles of the query's own fold are skipped
  int64_t idx_base;
  int K;
  uint64_t* keys_out;      // [splits][Q][K]
  uint32_t* tau;           // [Q] order-preserving image of a lower bound of the query's global KCAP-th best score
  float* debug_scores;     // optional [Q][N] dump of every score (bring-up / tests)
};

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address
  d |= static_cast<uint64_t>(0) << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// instruction descriptor: D=F32, A=B=BF16, both K-major, N=256, M=128
constexpr uint32_t T_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((T_BN >> 3) << 17) | ((T_BM >> 4) << 24);

// ---- register-resident sorted Top-K list ------------------------------------------------
template <int KCAP>
struct RegTopK {
  float s[KCAP];
  uint32_t i[KCAP];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) { s[j] = -INFINITY; i[j] = 0xFFFFFFFFu; }
  }
  __device__ __forceinline__ float threshold() const { return s[KCAP - 1]; }
  // candidates reach a thread in ascending index order, so strict '>' keeps the lower index on ties
  __device__ __forceinline__ void insert(float c, uint32_t ci) {
#pragma unroll
    for (int j = KCAP - 1; j > 0; --j) {
      const bool up = c > s[j - 1];
      const bool here = !up && (c > s[j]);
      s[j] = up ? s[j - 1] : (here ? c : s[j]);
      i[j] = up ? i[j - 1] : (here ? ci : i[j]);
    }
    const bool top = c > s[0];
    s[0] = top ? c : s[0];
    i[0] = top ? ci : i[0];
  }
};

__device__ __forceinline__ float select32(const float (&v)[32], int idx) {
  // 5-level select tree on the bits of idx (keeps v[] in registers)
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (idx & 1) ? v[2 * j + 1] : v[2 * j];
  float b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = (idx & 2) ? a[2 * j + 1] : a[2 * j];
  float c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = (idx & 4) ? b[2 * j + 1] : b[2 * j];
  const float d0 = (idx & 8) ? c[1] : c[0];
  const float d1 = (idx & 8) ? c[3] : c[2];
  return (idx & 16) ? d1 : d0;
}

// CV rule at tile granularity (fold-sorted inputs): if every query of the tile is in fold f and every
// database row of the tile is in fold f, all 128 x 256 pairs are inadmissible -- no TMA, no MMA, no epilogue.
// The three warp roles evaluate the same predicate from the same global bytes.
__device__ __forceinline__ int unit_fold(const TcParams& p, int64_t mt) {
  if (!p.fold_sorted || p.q_fold == nullptr) return -1;
  const int64_t a = mt * T_BM, b = (a + T_BM - 1 < p.Q) ? a + T_BM - 1 : p.Q - 1;
  const int lo = __ldg(p.q_fold + a), hi = __ldg(p.q_fold + b);
  return lo == hi ? lo : -1;
}
__device__ __forceinline__ bool tile_skipped(const TcParams& p, int ufold, int64_t t) {
  if (ufold < 0) return false;
  const int64_t a = t * T_BN, b = (a + T_BN - 1 < p.N) ? a + T_BN - 1 : p.N - 1;
  return __ldg(p.db_fold + a) == ufold && __ldg(p.db_fold + b) == ufold;
}

template <int PASSES, int KCAP, bool HAS_FOLD>
__global__ void __launch_bounds__(T_THREADS, 1)
tc_topk_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
               const __grid_constant__ CUtensorMap tm_db_hi, const __grid_constant__ CUtensorMap tm_db_lo,
               const TcParams p) {
  constexpr int PLANES = PASSES == 3 ? 2 : 1;
  constexpr uint32_t STAGE_BYTES = PLANES * (T_A_BYTES + T_B_BYTES);
  constexpr int STAGES = PASSES == 3 ? 2 : 4;

  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tiles_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_db_hi)) : "memory");
    if (PASSES == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_q_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_db_lo)) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bar_tfull[a]), 1); mbar_init(smem_u32(&bar_tempty[a]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int64_t n_units = p.m_tiles * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int64_t split = u / p.m_tiles, mt = u - split * p.m_tiles;
        const int m0 = static_cast<int>(mt * T_BM);
        const int64_t t0 = split * p.tiles_per_split;
        const int64_t t1 = (t0 + p.tiles_per_split < p.n_tiles) ? t0 + p.tiles_per_split : p.n_tiles;
        const int ufold = HAS_FOLD ? unit_fold(p, mt) : -1;
        for (int64_t t = t0; t < t1; ++t) {
          if (HAS_FOLD && tile_skipped(p, ufold, t)) continue;
          const int n0 = static_cast<int>(t * T_BN);
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_arrive_expect_tx(full, STAGE_BYTES);
            const uint32_t sb = tiles_base + stage * STAGE_BYTES;
            tma_load_2d(sb, &tm_q_hi, full, kc * T_BK, m0);
            tma_load_2d(sb + PLANES * T_A_BYTES, &tm_db_hi, full, kc * T_BK, n0);
            if (PASSES == 3) {
              tma_load_2d(sb + T_A_BYTES, &tm_q_lo, full, kc * T_BK, m0);
              tma_load_2d(sb + PLANES * T_A_BYTES + T_B_BYTES, &tm_db_lo, full, kc * T_BK, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int64_t split = u / p.m_tiles;
        const int64_t t0 = split * p.tiles_per_split;
        const int64_t t1 = (t0 + p.tiles_per_split < p.n_tiles) ? t0 + p.tiles_per_split : p.n_tiles;
        const int ufold = HAS_FOLD ? unit_fold(p, u - split * p.m_tiles) : -1;
        for (int64_t t = t0; t < t1; ++t) {
          if (HAS_FOLD && tile_skipped(p, ufold, t)) continue;
          mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * T_BN);
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tcgen05_fence_after();
            const uint32_t sb = tiles_base + stage * STAGE_BYTES;
            const uint64_t a_hi = make_smem_desc(sb);
            const uint64_t b_hi = make_smem_desc(sb + PLANES * T_A_BYTES);
            const uint64_t a_lo = make_smem_desc(sb + T_A_BYTES);
            const uint64_t b_lo = make_smem_desc(sb + PLANES * T_A_BYTES + T_B_BYTES);
#pragma unroll
            for (int k = 0; k < T_BK / 16; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * 32) >> 4);     // 16 bf16 = 32 bytes along the swizzled row
              if (PASSES == 3) {
                // small cross terms first, dominant term last
                umma_bf16(d_tmem, a_hi + koff, b_lo + koff, T_IDESC, (kc | k) != 0 ? 1u : 0u);
                umma_bf16(d_tmem, a_lo + koff, b_hi + koff, T_IDESC, 1u);
                umma_bf16(d_tmem, a_hi + koff, b_hi + koff, T_IDESC, 1u);
              } else {
                umma_bf16(d_tmem, a_hi + koff, b_hi + koff, T_IDESC, (kc | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(smem_u32(&bar_empty[stage]));       // smem slot free once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(smem_u32(&bar_tfull[acc]));           // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: fused Top-K =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    RegTopK<KCAP> top;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int64_t split = u / p.m_tiles, mt = u - split * p.m_tiles;
      const int64_t q = mt * T_BM + row;
      const int64_t t0 = split * p.tiles_per_split;
      const int64_t t1 = (t0 + p.tiles_per_split < p.n_tiles) ? t0 + p.tiles_per_split : p.n_tiles;
      const uint32_t my_fold = (HAS_FOLD && q < p.Q) ? p.q_fold[q] : 0xFFFFu;
      top.reset();
      // Units of the same query tile that finished earlier published their KCAP-th best score: the global
      // KCAP-th best is at least that, so rows strictly below it can be skipped here without changing the
      // merged list (">= bound" is kept, hence the step down by one ulp).  It removes the list warm-up --
      // a burst of warp-serialised insertions -- from every unit but the first of a query tile.
      float thr0 = -INFINITY;
      if (p.tau != nullptr && q < p.Q) {
        const uint32_t t = __ldcg(p.tau + q);
        if (t != 0u) thr0 = __uint_as_float(__float_as_uint(unorder_f32(t)) - ((unorder_f32(t) > 0.f) ? 1u : 0u) + ((unorder_f32(t) < 0.f) ? 1u : 0u));
      }
      float thr = thr0;
      const int ufold = HAS_FOLD ? unit_fold(p, mt) : -1;
      for (int64_t t = t0; t < t1; ++t) {
        if (HAS_FOLD && tile_skipped(p, ufold, t)) continue;
        const int64_t n0 = t * T_BN;
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tcgen05_fence_after();
        const bool edge = (n0 + T_BN > p.N);
#pragma unroll 1
        for (int ch = 0; ch < T_BN / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * T_BN + ch * 32), v);
          const int64_t c0 = n0 + ch * 32;
          if (p.debug_scores && q < p.Q) {
#pragma unroll
            for (int c = 0; c < 32; ++c) if (c0 + c < p.N) p.debug_scores[q * p.N + c0 + c] = v[c];
          }
          if (edge) {
#pragma unroll
            for (int c = 0; c < 32; ++c) if (c0 + c >= p.N) v[c] = -INFINITY;
          }
          if (HAS_FOLD) {
            const uint4* fp = reinterpret_cast<const uint4*>(p.db_fold + c0);
            const uint4 f0 = __ldg(fp), f1 = __ldg(fp + 1);
            const uint32_t w[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const uint32_t f = (w[c >> 2] >> (8 * (c & 3))) & 0xFFu;
              v[c] = (f == my_fold) ? -INFINITY : v[c];
            }
          }
          float mx = v[0];
#pragma unroll
          for (int c = 1; c < 32; ++c) mx = fmaxf(mx, v[c]);
          if (mx > thr) {
            uint32_t mask = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) mask |= (v[c] > thr) ? (1u << c) : 0u;
            while (mask) {
              const int c = __ffs(mask) - 1;
              mask &= mask - 1u;
              const float val = select32(v, c);
              if (val > thr) {
                top.insert(val, static_cast<uint32_t>(c0 + c + p.idx_base));
                thr = fmaxf(top.threshold(), thr0);
              }
            }
          }
          __syncwarp();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (q < p.Q) {
        if (p.tau != nullptr && top.i[KCAP - 1] != 0xFFFFFFFFu) atomicMax(p.tau + q, order_f32(top.s[KCAP - 1]));
        uint64_t* dst = p.keys_out + (split * p.Q + q) * p.K;
#pragma unroll
        for (int j = 0; j < KCAP; ++j)
          if (j < p.K) dst[j] = (top.i[j] == 0xFFFFFFFFu) ? 0ull : pack_key(top.s[j], top.i[j]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---- host side ------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

static int make_plane_map(CUtensorMap* map, const uint16_t* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(EMR2A_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(T_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EMR2A_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d): rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld);
  return EMR2A_OK;
}

// When the caller asks for it, the per-split lists are left unmerged and described here.
struct TcPartials {
  const uint64_t* parts;   // [splits][Q][K]
  int splits;
  const uint32_t* tau;     // [Q] max over splits of the split's K-th (last kept) score, 0 = no list was full; null if splits == 1
};

struct TcPlan {
  int64_t m_tiles, n_tiles, tiles_per_split;
  int splits;
  int grid;
  size_t keys_bytes, fold_bytes, tau_bytes;
};

static TcPlan tc_plan(int64_t Q, int64_t N, int K, bool has_fold) {
  TcPlan pl{};
  pl.m_tiles = (Q + T_BM - 1) / T_BM;
  pl.n_tiles = (N + T_BN - 1) / T_BN;
  const int sms = sm_count();
  // pick the split count whose unit count fills whole waves of CTAs best
  int best_s = 1; double best_eff = -1.0;
  const int64_t max_s = pl.n_tiles < 64 ? pl.n_tiles : 64;
  for (int64_t s = 1; s <= max_s; ++s) {
    const int64_t tps = (pl.n_tiles + s - 1) / s;
    const int64_t s_eff = (pl.n_tiles + tps - 1) / tps;       // splits that actually get tiles
    if (s_eff != s) continue;
    const int64_t units = pl.m_tiles * s;
    const int64_t waves = (units + sms - 1) / sms;
    // per-CTA time ~ waves * tps tiles (+ ~2 tiles of list warm-up per unit)
    const double cost = static_cast<double>(waves) * (static_cast<double>(tps) + 2.0);
    const double ideal = static_cast<double>(pl.m_tiles) * pl.n_tiles / sms;
    const double eff = ideal / cost;
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = static_cast<int>(s); }
  }
  pl.splits = best_s;
  pl.tiles_per_split = (pl.n_tiles + pl.splits - 1) / pl.splits;
  const int64_t units = pl.m_tiles * pl.splits;
  pl.grid = static_cast<int>(units < sms ? units : sms);
  pl.keys_bytes = sizeof(uint64_t) * static_cast<size_t>(pl.splits) * Q * K;
  pl.fold_bytes = has_fold ? static_cast<size_t>(pl.n_tiles) * T_BN : 0;
  pl.tau_bytes = (sizeof(uint32_t) * static_cast<size_t>(Q) + 255) & ~static_cast<size_t>(255);
  return pl;
}

size_t tc_topk_workspace_bytes(int64_t Q, int64_t N, int K) {
  TcPlan pl = tc_plan(Q, N, K, true);
  return ((pl.keys_bytes + 255) & ~static_cast<size_t>(255)) + pl.fold_bytes + pl.tau_bytes + 512;
}

template <int PASSES, int KCAP, bool HAS_FOLD>
static int tc_launch(const CUtensorMap& mq_hi, const CUtensorMap& mq_lo, const CUtensorMap& md_hi, const CUtensorMap& md_lo,
                     const TcParams& p, int grid, cudaStream_t st) {
  constexpr int PLANES = PASSES == 3 ? 2 : 1;
  constexpr int STAGES = PASSES == 3 ? 2 : 4;
  const size_t smem = static_cast<size_t>(STAGES) * PLANES * (T_A_BYTES + T_B_BYTES) + 1024;
  auto kern = tc_topk_kernel<PASSES, KCAP, HAS_FOLD>;
  EMR2A_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, T_THREADS, smem, st>>>(mq_hi, mq_lo, md_hi, md_lo, p);
  EMR2A_LAUNCH_CHECK("tc_topk_kernel");
  return EMR2A_OK;
}

template <int PASSES, int KCAP>
static int tc_launch_fold(bool has_fold, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& d,
                          const TcParams& p, int grid, cudaStream_t st) {
  return has_fold ? tc_launch<PASSES, KCAP, true>(a, b, c, d, p, grid, st)
                  : tc_launch<PASSES, KCAP, false>(a, b, c, d, p, grid, st);
}

int tc_topk_search(const uint16_t* q_hi, const uint16_t* q_lo, const uint16_t* db_hi, const uint16_t* db_lo,
                   int64_t Q, int64_t N, int D, int64_t ldq, int64_t lddb,
                   const uint8_t* q_fold, const uint8_t* db_fold, int64_t idx_base, int K, int passes,
                   uint64_t* out_keys, void* workspace, size_t ws_bytes, float* debug_scores, cudaStream_t st,
                   TcPartials* partials, int fold_sorted) {
  if (K > 32) return fail(EMR2A_ERR_UNSUPPORTED, "topk_search(bf16): K=%d > 32 (use EMR2A_PREC_FP32)", K);
  const int64_t Dp = (static_cast<int64_t>(D) + T_BK - 1) / T_BK * T_BK;
  if (ldq < Dp || lddb < Dp || (ldq % 8) || (lddb % 8))
    return fail(EMR2A_ERR_UNSUPPORTED, "topk_search(bf16): planes need ld >= round_up(D,64) and ld %% 8 == 0 (ldq=%lld lddb=%lld D=%d)", (long long)ldq, (long long)lddb, D);
  if (!q_hi || !db_hi || (passes == 3 && (!q_lo || !db_lo))) return fail(EMR2A_ERR_INVALID, "topk_search(bf16): missing operand plane");
  auto mis = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) != 0; };
  if (mis(q_hi) || mis(db_hi) || (passes == 3 && (mis(q_lo) || mis(db_lo))))
    return fail(EMR2A_ERR_UNSUPPORTED, "topk_search(bf16): operand planes must be 16-byte aligned");
  if (Q >= (1LL << 31) || N >= (1LL << 31) - T_BN) return fail(EMR2A_ERR_UNSUPPORTED, "topk_search(bf16): Q/N too large for one call");
  if (N + idx_base >= 0xFFFFFFFFLL) return fail(EMR2A_ERR_UNSUPPORTED, "topk_search: global index exceeds 32 bits");
  const bool has_fold = q_fold != nullptr;
  TcPlan pl = tc_plan(Q, N, K, has_fold);
  const size_t keys_off = 0;
  const size_t fold_off = (pl.keys_bytes + 255) & ~static_cast<size_t>(255);
  const size_t tau_off = (fold_off + pl.fold_bytes + 255) & ~static_cast<size_t>(255);
  const size_t need = tau_off + pl.tau_bytes;
  if (!workspace || ws_bytes < need) return fail(EMR2A_ERR_WORKSPACE, "topk_search(bf16): workspace %zu < %zu", ws_bytes, need);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(EMR2A_ERR_INVALID, "topk_search(bf16): workspace must be 256-byte aligned");

  CUtensorMap mq_hi, mq_lo, md_hi, md_lo;
  int rc;
  if ((rc = make_plane_map(&mq_hi, q_hi, Q, Dp, ldq, T_BM)) != EMR2A_OK) return rc;
  if ((rc = make_plane_map(&md_hi, db_hi, N, Dp, lddb, T_BN)) != EMR2A_OK) return rc;
  if (passes == 3) {
    if ((rc = make_plane_map(&mq_lo, q_lo, Q, Dp, ldq, T_BM)) != EMR2A_OK) return rc;
    if ((rc = make_plane_map(&md_lo, db_lo, N, Dp, lddb, T_BN)) != EMR2A_OK) return rc;
  } else {
    mq_lo = mq_hi; md_lo = md_hi;
  }
  TcParams p{};
  p.Q = Q; p.N = N; p.k_chunks = static_cast<int>(Dp / T_BK);
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splits = pl.splits; p.tiles_per_split = pl.tiles_per_split;
  p.idx_base = idx_base; p.K = K; p.debug_scores = debug_scores;
  p.keys_out = (pl.splits > 1 || partials) ? reinterpret_cast<uint64_t*>(ws + keys_off) : out_keys;
  if (pl.splits > 1) {     // sharing only pays (and is only needed) when a query tile is searched by several units
    p.tau = reinterpret_cast<uint32_t*>(ws + tau_off);
    EMR2A_CUDA_TRY(cudaMemsetAsync(p.tau, 0, sizeof(uint32_t) * static_cast<size_t>(Q), st));
  }
  if (has_fold) {
    uint8_t* fpad = ws + fold_off;
    EMR2A_CUDA_TRY(cudaMemsetAsync(fpad, 0xFF, pl.fold_bytes, st));
    EMR2A_CUDA_TRY(cudaMemcpyAsync(fpad, db_fold, static_cast<size_t>(N), cudaMemcpyDeviceToDevice, st));
    p.q_fold = q_fold; p.db_fold = fpad; p.fold_sorted = fold_sorted;
  }
  const int kcap = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
  if (passes == 3) {
    if (kcap == 8) rc = tc_launch_fold<3, 8>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
    else if (kcap == 16) rc = tc_launch_fold<3, 16>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
    else rc = tc_launch_fold<3, 32>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
  } else {
    if (kcap == 8) rc = tc_launch_fold<1, 8>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
    else if (kcap == 16) rc = tc_launch_fold<1, 16>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
    else rc = tc_launch_fold<1, 32>(has_fold, mq_hi, mq_lo, md_hi, md_lo, p, pl.grid, st);
  }
  if (rc != EMR2A_OK) return rc;
  if (partials) {
    partials->parts = p.keys_out;
    partials->splits = pl.splits;
    partials->tau = p.tau;
    return EMR2A_OK;
  }
  if (pl.splits > 1)
    return emr2a_topk_merge(reinterpret_cast<const uint64_t*>(ws + keys_off), pl.splits, Q, K, Q * K, K, K, out_keys, st);
  return EMR2A_OK;
}

}  // namespace emr2a
