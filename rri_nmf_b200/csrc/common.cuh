// common.cuh -- shared device helpers for the RRI sweep kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RRI_DEVINL __device__ __forceinline__

namespace rri {

constexpr int WARP = 32;

template <typename T> struct Vec;            // 16-byte vector type per scalar
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T> RRI_DEVINL T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (evict-first) 16-byte loads of X: the data matrix is read once per pass and must not
// displace the factor tiles that live in L2
RRI_DEVINL float4 ld_stream(const float4* p) { return __ldcs(p); }
RRI_DEVINL double2 ld_stream(const double2* p) { return __ldcs(p); }
RRI_DEVINL float ld_stream(const float* p) { return __ldcs(p); }
RRI_DEVINL double ld_stream(const double* p) { return __ldcs(p); }

RRI_DEVINL void unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
RRI_DEVINL void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }

template <typename T> RRI_DEVINL T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> RRI_DEVINL T tmin(T a, T b) { return a < b ? a : b; }

// balanced contiguous partition of [0,n) into `parts`: range of part p
RRI_DEVINL void part_range(int64_t n, int parts, int p, int64_t& b, int64_t& e) {
    int64_t q = n / parts, r = n % parts;
    b = q * p + (p < r ? p : r);
    e = b + q + (p < r ? 1 : 0);
}

// Scalar-c projected solve of optimization.py:51-67 applied to one element:
//   numer = (statistic - reg_l1)  [= -w of qf_min], denom = c.
//   c > 0 : x = max(numer,0)/(c+eps)                       (:53-55; ub ignored)
//   c <= 0: with ub: x = ub where (c - numer) < 0, else 0    (:60-65)
//           without ub the reference raises unconditionally  (:66-67 -> :105-107) -> flag
template <typename T>
RRI_DEVINL T solve_scalar_c(T numer, T denom, T eps, T ub, bool has_ub, bool& unbounded) {
    if (denom > T(0)) return tmax(numer, T(0)) / (denom + eps);
    if (!has_ub) { unbounded = true; return T(0); }
    return ((denom - numer) < T(0)) ? ub : T(0);
}

// Vector-c solve of optimization.py:75-84 for one element (masked WRRI): zero where c<=0, clip to ub.
template <typename T>
RRI_DEVINL T solve_vector_c(T numer, T denom, T eps, T ub, bool has_ub, bool& unbounded) {
    T x = T(0);
    if (denom > T(0)) x = tmax(numer, T(0)) / (denom + eps);
    else if (denom < T(0) && !has_ub) unbounded = true;      // :76-77
    if (has_ub) x = tmin(x, ub);
    return x;
}

}  // namespace rri
