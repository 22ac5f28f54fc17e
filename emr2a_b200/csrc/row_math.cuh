// Row arithmetic shared by K1 (normalize_fuse.cu) and the kernels that RE-CREATE K1's fp32 output rows on the fly
// from the raw rows (rescore.cu, "deferred fp32 rows"): loaders, the correctly rounded row division, and LazyRows.
#pragma once
#include "common.cuh"

namespace emr2a {

template <typename InT> struct Loader;
template <> struct Loader<float> {
  typedef float4 Raw;
  static __device__ __forceinline__ Raw load_raw(const void* base, int64_t elem) {
    return ldg_stream_f4(reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem));
  }
  static __device__ __forceinline__ float4 widen(const Raw& r) { return r; }
  static __device__ __forceinline__ float4 load4(const void* base, int64_t elem) {
    return ldg_stream_f4(reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem));
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t elem) {
    return __ldg(static_cast<const float*>(base) + elem);
  }
};
template <> struct Loader<__nv_bfloat16> {
  typedef uint2 Raw;
  static __device__ __forceinline__ Raw load_raw(const void* base, int64_t elem) {
    return ldg_stream_u2(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + elem));
  }
  static __device__ __forceinline__ float4 widen(const Raw& r) {
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xFFFF0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xFFFF0000u);
    return f;
  }
  static __device__ __forceinline__ float4 load4(const void* base, int64_t elem) {
    uint2 r = ldg_stream_u2(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + elem));
    float4 f;
    f.x = __uint_as_float(r.x << 16);
    f.y = __uint_as_float(r.x & 0xFFFF0000u);
    f.z = __uint_as_float(r.y << 16);
    f.w = __uint_as_float(r.y & 0xFFFF0000u);
    return f;
  }
  static __device__ __forceinline__ float load1(const void* base, int64_t elem) {
    uint16_t r = __ldg(static_cast<const uint16_t*>(base) + elem);
    return __uint_as_float(static_cast<uint32_t>(r) << 16);
  }
};

__device__ __forceinline__ float sq4(const float4& a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }
// Division of a whole row by one divisor: r = RN(1/d) once per row, then per element
// q0 = x*r, rem = x - q0*d (exact in an FMA), q = q0 + rem*r -- the correctly rounded quotient
// x/d (Markstein) without the per-element special-case branches of __fdiv_rn, which made K1
// ALU-bound (3160 instructions per row).  Inputs are finite and d >= 1e-8.
struct RowDiv {
  float d, r;
  __device__ __forceinline__ float operator()(float x) const {
    const float q0 = x * r;
    const float rem = fmaf(-q0, d, x);
    return fmaf(rem, r, q0);
  }
};
__device__ __forceinline__ RowDiv row_div(float d) {
  RowDiv v;
  v.d = d;
  v.r = __frcp_rn(d);
  return v;
}
__device__ __forceinline__ void div4(float4& a, const RowDiv& dv) {
  a.x = dv(a.x); a.y = dv(a.y); a.z = dv(a.z); a.w = dv(a.w);
}
__device__ __forceinline__ void mul4(float4& a, float w) { a.x *= w; a.y *= w; a.z *= w; a.w *= w; }


// ---- deferred fp32 rows ---------------------------------------------------------------------------------------
// The RESCORE arm needs the fp32 rows K1 produces only for the few candidates it re-scores (and for the rare exact
// re-scan).  Writing them for the whole database costs 4 bytes per element of HBM traffic in K1 (C2: 4.1 of 10.2 GB)
// and as much HBM capacity.  Instead K1 can record, per row, the divisors it used (row_div_out [n][4]):
//     [0] ||seg0|| + 1e-8   [1] ||seg1|| + 1e-8   (1.0 without SEGNORM)
//     [2] the row divisor (||row|| + 1e-8, or ||row|| with ZERO_GUARD)   [3] 1.0 if that division was applied, else 0.0
// and the consumers apply the same operations to the raw element -- x -> RN(x / n_s) -> * w_s -> RN(. / n_row) -- which
// gives bit for bit the value K1 would have stored: every step is a correctly rounded IEEE operation on the same inputs.
struct LazyRows {
  const void* seg0;
  const void* seg1;
  int d0, d1;
  int64_t ld0, ld1;
  float w0, w1;
  int flags;
  const float* row_div;
};

struct LazyRowCtx {
  RowDiv n0, n1, dv;
  bool seg, weight, row;
};

__device__ __forceinline__ LazyRowCtx lazy_ctx_from(const LazyRows& L, const float4& d) {
  LazyRowCtx c;
  c.n0 = row_div(d.x); c.n1 = row_div(d.y); c.dv = row_div(d.z);
  c.seg = (L.flags & EMR2A_NF_SEGNORM) != 0;
  c.weight = L.w0 != 1.0f || L.w1 != 1.0f;
  c.row = d.w != 0.f;
  return c;
}
__device__ __forceinline__ LazyRowCtx lazy_row_ctx(const LazyRows& L, int64_t row) {
  return lazy_ctx_from(L, __ldg(reinterpret_cast<const float4*>(L.row_div) + row));
}

// Raw elements [e, e + 4) of row `row` (e % 4 == 0; d0 % 4 == 0 so that a chunk never straddles the segments).  Plain
// read-only loads, NOT the `asm volatile` streaming loads of K1: the gather kernels issue the loads of several rows
// before the arithmetic of the first one, and volatile asm statements may not be reordered against each other.
template <typename InT> struct RawChunk;
template <> struct RawChunk<float> {
  typedef float4 T;
  static __device__ __forceinline__ T ld(const void* base, int64_t elem) {
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem));
  }
  static __device__ __forceinline__ float4 widen(const T& r) { return r; }
};
template <> struct RawChunk<__nv_bfloat16> {
  typedef uint2 T;
  static __device__ __forceinline__ T ld(const void* base, int64_t elem) {
    return __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + elem));
  }
  static __device__ __forceinline__ float4 widen(const T& r) { return Loader<__nv_bfloat16>::widen(r); }
};
template <typename InT>
__device__ __forceinline__ typename RawChunk<InT>::T lazy_raw4(const LazyRows& L, int64_t row, int e) {
  return e < L.d0 ? RawChunk<InT>::ld(L.seg0, row * L.ld0 + e) : RawChunk<InT>::ld(L.seg1, row * L.ld1 + (e - L.d0));
}
// K1's arithmetic on a raw chunk: the fp32 values K1 would have stored for elements [e, e + 4) of the row
template <typename InT>
__device__ __forceinline__ float4 lazy_apply4(const LazyRows& L, const LazyRowCtx& c, const typename RawChunk<InT>::T& raw, int e) {
  const bool s0 = e < L.d0;
  float4 v = RawChunk<InT>::widen(raw);
  if (c.seg) div4(v, s0 ? c.n0 : c.n1);
  if (c.weight) mul4(v, s0 ? L.w0 : L.w1);
  if (c.row) div4(v, c.dv);
  return v;
}
template <typename InT>
__device__ __forceinline__ float4 lazy_load4(const LazyRows& L, const LazyRowCtx& c, int64_t row, int e) {
  return lazy_apply4<InT>(L, c, lazy_raw4<InT>(L, row, e), e);
}

}  // namespace emr2a
