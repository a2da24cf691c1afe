// Interval arithmetic for the tile evaluator (device only).
//
// An interval is a float2 {lo, hi}.  Semantics follow the reference's
// mpr::Interval operator set (reference inc/gpu_interval.hpp, line numbers
// cited per function) bit for bit: outward directed rounding on + - * / sqrt,
// strict comparisons in the sign-class tests, min/max that report which
// operand dominated, cos/sin that always answer [-1, 1], log that clamps its
// lower bound at 0.  The transcendental bounds are the single-precision
// libdevice functions: the reference writes `__double2float_rd(::acos(x))`
// with a float x, which overload resolution turns into acosf followed by an
// exact float->double->float round trip (checked in the reference's PTX:
// only cvt.f64.f32 / cvt.rm|rp.f32.f64, no f64 arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace mprb {

typedef float2 ival;  // .x = lower bound, .y = upper bound

__device__ __forceinline__ ival iv(float lo, float hi) { return make_float2(lo, hi); }

// gpu_interval.hpp:65-67
__device__ __forceinline__ ival iv_neg(ival a) { return iv(-a.y, -a.x); }

// gpu_interval.hpp:71-81
__device__ __forceinline__ ival iv_add(ival a, ival b) {
    return iv(__fadd_rd(a.x, b.x), __fadd_ru(a.y, b.y));
}
__device__ __forceinline__ ival iv_add(ival a, float c) {
    return iv(__fadd_rd(a.x, c), __fadd_ru(a.y, c));
}

// gpu_interval.hpp:284-294
__device__ __forceinline__ ival iv_sub(ival a, ival b) {
    return iv(__fsub_rd(a.x, b.y), __fsub_ru(a.y, b.x));
}
__device__ __forceinline__ ival iv_sub(ival a, float c) {
    return iv(__fsub_rd(a.x, c), __fsub_ru(a.y, c));
}
__device__ __forceinline__ ival iv_sub(float c, ival b) {
    return iv(__fsub_rd(c, b.y), __fsub_ru(c, b.x));
}

// Sign class of an interval under the reference's strict tests:
//   bit0 = (lo < 0), bit1 = (hi > 0)  ->  0: zero-ish, 1: negative, 2: positive, 3: mixed
// NaN bounds fail both tests and land in class 0, exactly as the nested ifs
// of gpu_interval.hpp:85-146 do.
__device__ __forceinline__ int iv_class(ival a) {
    return (a.x < 0.0f ? 1 : 0) | (a.y > 0.0f ? 2 : 0);
}

// gpu_interval.hpp:85-146 (nine sign cases; any zero-ish operand gives [0, 0])
__device__ __forceinline__ ival iv_mul(ival a, ival b) {
    const int ca = iv_class(a), cb = iv_class(b);
    if (ca == 0 || cb == 0) {
        return iv(0.0f, 0.0f);
    }
    // Pick the operand pair whose product bounds each side.
    float l0, l1, h0, h1;
    if (ca == 3) {                      // a spans zero
        if (cb == 3) {                  // mixed * mixed: two candidates per side
            return iv(fminf(__fmul_rd(a.x, b.y), __fmul_rd(a.y, b.x)),
                      fmaxf(__fmul_ru(a.x, b.x), __fmul_ru(a.y, b.y)));
        } else if (cb == 1) {           // mixed * negative
            l0 = a.y; l1 = b.x; h0 = a.x; h1 = b.x;
        } else {                        // mixed * positive
            l0 = a.x; l1 = b.y; h0 = a.y; h1 = b.y;
        }
    } else if (ca == 1) {               // a negative
        if (cb == 3)      { l0 = a.x; l1 = b.y; h0 = a.x; h1 = b.x; }
        else if (cb == 1) { l0 = a.y; l1 = b.y; h0 = a.x; h1 = b.x; }
        else              { l0 = a.x; l1 = b.y; h0 = a.y; h1 = b.x; }
    } else {                            // a positive
        if (cb == 3)      { l0 = a.y; l1 = b.x; h0 = a.y; h1 = b.y; }
        else if (cb == 1) { l0 = a.y; l1 = b.x; h0 = a.x; h1 = b.y; }
        else              { l0 = a.x; l1 = b.x; h0 = a.y; h1 = b.y; }
    }
    return iv(__fmul_rd(l0, l1), __fmul_ru(h0, h1));
}

// gpu_interval.hpp:148-154
__device__ __forceinline__ ival iv_mul(ival a, float c) {
    return (c < 0.0f) ? iv(__fmul_rd(a.y, c), __fmul_ru(a.x, c))
                      : iv(__fmul_rd(a.x, c), __fmul_ru(a.y, c));
}

// gpu_interval.hpp:162-190
__device__ __forceinline__ ival iv_div(ival a, ival b) {
    if (b.x <= 0.0f && b.y >= 0.0f) {
        return iv(-CUDART_INF_F, CUDART_INF_F);
    }
    const bool bneg = b.y < 0.0f;
    if (a.y < 0.0f) {
        return bneg ? iv(__fdiv_rd(a.y, b.x), __fdiv_ru(a.x, b.y))
                    : iv(__fdiv_rd(a.x, b.x), __fdiv_ru(a.y, b.y));
    } else if (a.x < 0.0f) {
        return bneg ? iv(__fdiv_rd(a.y, b.y), __fdiv_ru(a.x, b.y))
                    : iv(__fdiv_rd(a.x, b.x), __fdiv_ru(a.y, b.x));
    } else {
        return bneg ? iv(__fdiv_rd(a.y, b.y), __fdiv_ru(a.x, b.x))
                    : iv(__fdiv_rd(a.x, b.y), __fdiv_ru(a.y, b.x));
    }
}

// gpu_interval.hpp:192-200
__device__ __forceinline__ ival iv_div(ival a, float c) {
    if (c < 0.0f) return iv(__fdiv_rd(a.y, c), __fdiv_ru(a.x, c));
    if (c > 0.0f) return iv(__fdiv_rd(a.x, c), __fdiv_ru(a.y, c));
    return iv(-CUDART_INF_F, CUDART_INF_F);
}
// gpu_interval.hpp:202-204
__device__ __forceinline__ ival iv_div(float c, ival b) { return iv_div(iv(c, c), b); }

// min / max with the 2-bit verdict: 0 = undecided, 1 = first operand, 2 = second.
// gpu_interval.hpp:208-252
__device__ __forceinline__ ival iv_min(ival a, ival b, int& choice) {
    if (a.y < b.x) { choice = 1; return a; }
    if (b.y < a.x) { choice = 2; return b; }
    choice = 0;
    return iv(fminf(a.x, b.x), fminf(a.y, b.y));
}
__device__ __forceinline__ ival iv_max(ival a, ival b, int& choice) {
    if (a.x > b.y) { choice = 1; return a; }
    if (b.x > a.y) { choice = 2; return b; }
    choice = 0;
    return iv(fmaxf(a.x, b.x), fmaxf(a.y, b.y));
}

// gpu_interval.hpp:256-266
__device__ __forceinline__ ival iv_square(ival a) {
    if (a.y < 0.0f) return iv(__fmul_rd(a.y, a.y), __fmul_ru(a.x, a.x));
    if (a.x > 0.0f) return iv(__fmul_rd(a.x, a.x), __fmul_ru(a.y, a.y));
    if (-a.x > a.y) return iv(0.0f, __fmul_ru(a.x, a.x));
    return iv(0.0f, __fmul_ru(a.y, a.y));
}

// gpu_interval.hpp:268-276
__device__ __forceinline__ ival iv_abs(ival a) {
    if (a.x >= 0.0f) return a;
    if (a.y < 0.0f) return iv_neg(a);
    return iv(0.0f, fmaxf(-a.x, a.y));
}

// gpu_interval.hpp:296-304
__device__ __forceinline__ ival iv_sqrt(ival a) {
    if (a.y < 0.0f) return iv(CUDART_NAN_F, CUDART_NAN_F);
    if (a.x <= 0.0f) return iv(0.0f, __fsqrt_ru(a.y));
    return iv(__fsqrt_rd(a.x), __fsqrt_ru(a.y));
}

// gpu_interval.hpp:306-336, 382-391.  See the header comment for why these
// are the float libdevice entry points.
__device__ __forceinline__ ival iv_acos(ival a) {
    if (a.y < -1.0f || a.x > 1.0f) return iv(CUDART_NAN_F, CUDART_NAN_F);
    return iv(acosf(a.y), acosf(a.x));
}
__device__ __forceinline__ ival iv_asin(ival a) {
    if (a.y < -1.0f || a.x > 1.0f) return iv(CUDART_NAN_F, CUDART_NAN_F);
    return iv(asinf(a.x), asinf(a.y));
}
__device__ __forceinline__ ival iv_atan(ival a) { return iv(atanf(a.x), atanf(a.y)); }
__device__ __forceinline__ ival iv_exp(ival a) { return iv(expf(a.x), expf(a.y)); }
__device__ __forceinline__ ival iv_log(ival a) {
    if (a.y < 0.0f) return iv(CUDART_NAN_F, CUDART_NAN_F);
    if (a.x <= 0.0f) return iv(0.0f, logf(a.y));
    return iv(logf(a.x), logf(a.y));
}
// gpu_interval.hpp:346-380: the reference returns before any range reduction.
__device__ __forceinline__ ival iv_cos(ival) { return iv(-1.0f, 1.0f); }
__device__ __forceinline__ ival iv_sin(ival) { return iv(-1.0f, 1.0f); }

}  // namespace mprb
