// Forward-mode derivative values for the normal pass (device only).
//
// A dval is a float4 {d/dx, d/dy, d/dz, value}, the layout the reference's
// mpr::Deriv uses (reference inc/gpu_deriv.hpp:18-29).
//
// Rounding contract.  The reference is compiled with nvcc's default
// -fmad=true, so `a*b + c*d` patterns in gpu_deriv.hpp are contracted by
// ptxas.  Reading the SASS of the reference's eval_pixels_d built for sm_100a
// shows one consistent rule: the FIRST product is fused and the SECOND is
// rounded on its own, i.e.  a*b + c*d -> fma(a, b, rn(c*d))  and
// a*b - c*d -> fma(a, b, -rn(c*d));  v*v + 1 -> fma(v, v, 1);
// 1 - v*v -> fma(-v, v, 1).  powf(v, 2) in the quotient rule is a real
// libdevice powf call (not folded to v*v).  The formulas below spell those
// choices out with explicit _rn intrinsics (which ptxas never re-associates or
// contracts), so normals match the reference independent of how this
// translation unit happens to be scheduled.
#pragma once
#include <cuda_runtime.h>

namespace mprb {

typedef float4 dval;  // .x .y .z = gradient, .w = value

__device__ __forceinline__ dval dv(float v, float dx, float dy, float dz) {
    return make_float4(dx, dy, dz, v);
}
__device__ __forceinline__ dval dv_const(float v) { return make_float4(0.0f, 0.0f, 0.0f, v); }

// gpu_deriv.hpp:42-44
__device__ __forceinline__ dval dv_neg(dval a) { return dv(-a.w, -a.x, -a.y, -a.z); }

// gpu_deriv.hpp:48-61
__device__ __forceinline__ dval dv_add(dval a, dval b) {
    return dv(__fadd_rn(a.w, b.w), __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y),
              __fadd_rn(a.z, b.z));
}
__device__ __forceinline__ dval dv_add(dval a, float c) {
    return dv(__fadd_rn(a.w, c), a.x, a.y, a.z);
}

// gpu_deriv.hpp:65-81   a.d*b.v + b.d*a.v
__device__ __forceinline__ dval dv_mul(dval a, dval b) {
    return dv(__fmul_rn(a.w, b.w),
              __fmaf_rn(a.x, b.w, __fmul_rn(b.x, a.w)),
              __fmaf_rn(a.y, b.w, __fmul_rn(b.y, a.w)),
              __fmaf_rn(a.z, b.w, __fmul_rn(b.z, a.w)));
}
__device__ __forceinline__ dval dv_mul(dval a, float c) {
    return dv(__fmul_rn(a.w, c), __fmul_rn(a.x, c), __fmul_rn(a.y, c), __fmul_rn(a.z, c));
}

// gpu_deriv.hpp:85-104   (b.v*a.d - a.v*b.d) / b.v^2
__device__ __forceinline__ dval dv_div(dval a, dval b) {
    const float d = powf(b.w, 2);
    return dv(__fdiv_rn(a.w, b.w),
              __fdiv_rn(__fmaf_rn(b.w, a.x, -__fmul_rn(a.w, b.x)), d),
              __fdiv_rn(__fmaf_rn(b.w, a.y, -__fmul_rn(a.w, b.y)), d),
              __fdiv_rn(__fmaf_rn(b.w, a.z, -__fmul_rn(a.w, b.z)), d));
}
__device__ __forceinline__ dval dv_div(dval a, float c) {
    return dv(__fdiv_rn(a.w, c), __fdiv_rn(a.x, c), __fdiv_rn(a.y, c), __fdiv_rn(a.z, c));
}
__device__ __forceinline__ dval dv_div(float c, dval b) {
    const float d = powf(b.w, 2);
    return dv(__fdiv_rn(c, b.w),
              __fdiv_rn(__fmul_rn(-c, b.x), d),
              __fdiv_rn(__fmul_rn(-c, b.y), d),
              __fdiv_rn(__fmul_rn(-c, b.z), d));
}

// gpu_deriv.hpp:108-132: ties go to the second operand for min, first for max.
__device__ __forceinline__ dval dv_min(dval a, dval b) { return (a.w < b.w) ? a : b; }
__device__ __forceinline__ dval dv_min(dval a, float c) { return (a.w < c) ? a : dv_const(c); }
__device__ __forceinline__ dval dv_max(dval a, dval b) { return (a.w >= b.w) ? a : b; }
__device__ __forceinline__ dval dv_max(dval a, float c) { return (a.w >= c) ? a : dv_const(c); }

// gpu_deriv.hpp:143-149
__device__ __forceinline__ dval dv_abs(dval a) { return (a.w < 0.0f) ? dv_neg(a) : a; }

// gpu_deriv.hpp:153-166
__device__ __forceinline__ dval dv_sub(dval a, dval b) {
    return dv(__fsub_rn(a.w, b.w), __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y),
              __fsub_rn(a.z, b.z));
}
__device__ __forceinline__ dval dv_sub(dval a, float c) {
    return dv(__fsub_rn(a.w, c), a.x, a.y, a.z);
}
__device__ __forceinline__ dval dv_sub(float c, dval b) {
    return dv(__fsub_rn(c, b.w), -b.x, -b.y, -b.z);
}

// gpu_deriv.hpp:168-203
__device__ __forceinline__ dval dv_sqrt(dval a) {
    const float s = sqrtf(a.w);
    const float d = __fadd_rn(s, s);   // 2 * sqrt(v), exact either way
    return dv(s, __fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d));
}
__device__ __forceinline__ dval dv_atan(dval a) {
    const float d = __fmaf_rn(a.w, a.w, 1.0f);
    return dv(atanf(a.w), __fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d));
}
__device__ __forceinline__ dval dv_acos(dval a) {
    const float d = -sqrtf(__fmaf_rn(-a.w, a.w, 1.0f));
    return dv(acosf(a.w), __fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d));
}
__device__ __forceinline__ dval dv_asin(dval a) {
    const float d = sqrtf(__fmaf_rn(-a.w, a.w, 1.0f));
    return dv(asinf(a.w), __fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d));
}
__device__ __forceinline__ dval dv_exp(dval a) {
    const float v = expf(a.w);
    return dv(v, __fmul_rn(v, a.x), __fmul_rn(v, a.y), __fmul_rn(v, a.z));
}
__device__ __forceinline__ dval dv_cos(dval a) {
    const float s = -sinf(a.w);
    return dv(cosf(a.w), __fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z));
}
__device__ __forceinline__ dval dv_sin(dval a) {
    const float c = cosf(a.w);
    return dv(sinf(a.w), __fmul_rn(c, a.x), __fmul_rn(c, a.y), __fmul_rn(c, a.z));
}
__device__ __forceinline__ dval dv_log(dval a) {
    const float v = a.w;
    return dv(logf(v), __fdiv_rn(a.x, v), __fdiv_rn(a.y, v), __fdiv_rn(a.z, v));
}

}  // namespace mprb
