/*
 * ek_math.cuh -- device math with the numerics of the reference's CPU path.
 *
 * The reference's own GPU backend maps exp/log/sin/cos/rcp/rsqrt to PTX
 * `.approx.ftz` instructions (include/enoki/cuda.h:433-467).  BASELINE.json's
 * north star asks for parity with the reference's *CPU* path instead, so these
 * are the Cephes-derived polynomials of include/enoki/array_math.h evaluated with
 * the same Estrin groupings (array_math.h:25-105) and explicit IEEE add/mul/fma
 * intrinsics (never contracted, no flush-to-zero), which makes the results
 * bit-identical to an AVX2+FMA build of the reference compiled with
 * -ffp-contract=off.
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace ekm {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fbits(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t ubits(float f) { return __float_as_uint(f); }

/* x86 cvttps2dq semantics: out-of-range and NaN give the "integer indefinite" 0x80000000 */
__device__ __forceinline__ int32_t cvtt_f32_i32(float x) {
    return (x >= -2147483648.f && x < 2147483648.f) ? __float2int_rz(x) : (int32_t) 0x80000000;
}
/* Saturating conversion (cvt.rzi.s32.f32) where the difference to cvttps2dq cannot be observed:
   - sincos: the argument is >= 0; positive overflow saturates to 0x7fffffff and (j + 1) & ~1 maps it to the
     same 0x80000000 the CPU path produces; NaN ends in NaN either way;
   - exp: the argument is an integer in [-128, 128] unless the overflow/underflow masks or a NaN override it. */
__device__ __forceinline__ int32_t cvtt_sat(float x) { return __float2int_rz(x); }

/* array_math.h:25-33 */
__device__ __forceinline__ float poly2(float x, float c0, float c1, float c2) {
    float x2 = fmul(x, x);
    return ffma(x2, c2, ffma(x, c1, c0));
}
/* array_math.h:50-58 */
__device__ __forceinline__ float poly5(float x, float c0, float c1, float c2, float c3, float c4, float c5) {
    float x2 = fmul(x, x), x4 = fmul(x2, x2);
    return ffma(x2, ffma(x, c3, c2), ffma(x4, ffma(x, c5, c4), ffma(x, c1, c0)));
}
/* array_math.h:80-88 */
__device__ __forceinline__ float poly8(float x, float c0, float c1, float c2, float c3, float c4,
                                       float c5, float c6, float c7, float c8) {
    float x2 = fmul(x, x), x4 = fmul(x2, x2), x8 = fmul(x4, x4);
    return ffma(x4, ffma(x2, ffma(x, c7, c6), ffma(x, c5, c4)),
                ffma(x2, ffma(x, c3, c2), fadd(ffma(x, c1, c0), fmul(c8, x8))));
}

/* array_math.h:261-367 sincos_approx<Sin, Cos>, single precision branch */
template <bool Sin, bool Cos>
__device__ __forceinline__ void sincos(float x, float &s_out, float &c_out) {
    float xa = fabsf(x);
    int32_t j = cvtt_sat(fmul(xa, 1.2732395447351626862f));
    j = (j + 1) & ~1;
    float y = __int2float_rn(j);
    uint32_t sign_sin = ((uint32_t) j << 29) ^ ubits(x);
    uint32_t sign_cos = (uint32_t) (~(j - 2)) << 29;
    y = fsub(fsub(fsub(xa, fmul(y, 0.78515625f)), fmul(y, 2.4187564849853515625e-4f)),
             fmul(y, 3.77489497744594108e-8f));
    float z = fmul(y, y);
    if (xa == __int_as_float(0x7f800000)) z = fbits(0xffffffffu);
    float s = fmul(poly2(z, -1.6666654611e-1f, 8.3321608736e-3f, -1.9515295891e-4f), z);
    float c = fmul(poly2(z, 4.166664568298827e-2f, -1.388731625493765e-3f, 2.443315711809948e-5f), z);
    s = ffma(s, y, y);
    c = ffma(c, z, ffma(z, -0.5f, 1.f));
    bool polymask = (j & 2) == 0;
    if (Sin) s_out = fbits(ubits(polymask ? s : c) ^ (sign_sin & 0x80000000u));
    if (Cos) c_out = fbits(ubits(polymask ? c : s) ^ (sign_cos & 0x80000000u));
}
__device__ __forceinline__ float sin_f32(float x) { float s, c; sincos<true, false>(x, s, c); return s; }
__device__ __forceinline__ float cos_f32(float x) { float s, c; sincos<false, true>(x, s, c); return c; }

/* array_math.h:711-776 exp, single precision branch (+ ldexp :677-680) */
__device__ __forceinline__ float exp_f32(float x) {
    const float max_range = +88.3762588501f, min_range = -88.3762588501f;
    bool overflow = x > max_range, underflow = x < min_range;
    float n = floorf(ffma(1.4426950408889634073599f, x, 0.5f));
    float xr = ffma(-n, 0.693359375f, x);
    xr = ffma(-n, -2.12194440e-4f, xr);
    float z = poly5(xr, 5.0000001201e-1f, 1.6666665459e-1f, 4.1665795894e-2f,
                        8.3334519073e-3f, 1.3981999507e-3f, 1.9875691500e-4f);
    z = ffma(z, fmul(xr, xr), fadd(xr, 1.f));
    uint32_t scale = (uint32_t) (cvtt_sat(n) + 0x7f) << 23;
    float r = fmul(z, fbits(scale));
    return overflow ? __int_as_float(0x7f800000) : (underflow ? 0.f : r);
}

/* array_math.h:778-898 log, single precision branch, non-AVX512 (frexp :682-709) */
__device__ __forceinline__ float log_f32(float x) {
    bool valid = x >= 0.f;
    uint32_t xi = ubits(x);
    uint32_t exponent_bits = xi & 0x7f800000u;
    bool is_normal = (x != 0.f) && (exponent_bits != 0x7f800000u);
    int32_t exponent_i = (int32_t) (exponent_bits >> 23) - 0x7f;
    uint32_t mantissa = (xi & ~0x7f800000u) | 0x3f000000u;
    float xm = fbits(is_normal ? mantissa : xi);
    float e = __int2float_rn(is_normal ? exponent_i : 0);

    bool ge = xm >= 0.70710678118654752440f;
    if (ge) e = fadd(e, 1.f);
    xm = fadd(xm, fsub(ge ? 0.f : xm, 1.f));

    float z = fmul(xm, xm);
    float y = poly8(xm, 3.3333331174e-1f, -2.4999993993e-1f, 2.0000714765e-1f, -1.6668057665e-1f,
                        1.4249322787e-1f, -1.2420140846e-1f, 1.1676998740e-1f, -1.1514610310e-1f,
                        7.0376836292e-2f);
    y = fmul(y, fmul(xm, z));
    y = ffma(e, -2.12194440e-4f, y);
    z = ffma(z, -0.5f, fadd(xm, y));
    float r = ffma(e, 0.693359375f, z);
    if (x == __int_as_float(0x7f800000)) r = __int_as_float(0x7f800000);
    if (x == 0.f) r = __int_as_float(0xff800000);
    return valid ? r : fbits(0xffffffffu);
}

/* safe_mul / safe_fmadd: src/autodiff/autodiff.cpp:1191-1221 */
__device__ __forceinline__ float mul_nz(float a, float b) {
    return (a == 0.f || b == 0.f) ? 0.f : fmul(a, b);
}
__device__ __forceinline__ float fma_nz(float a, float b, float c) {
    return (a == 0.f || b == 0.f) ? c : ffma(a, b, c);
}
__device__ __forceinline__ double mul_nz(double a, double b) {
    return (a == 0.0 || b == 0.0) ? 0.0 : __dmul_rn(a, b);
}
__device__ __forceinline__ double fma_nz(double a, double b, double c) {
    return (a == 0.0 || b == 0.0) ? c : __fma_rn(a, b, c);
}

/* min/max with the x86 vminps/vmaxps operand rule of the CPU path
   (array_avx.h min_/max_: second operand returned when either is NaN) */
__device__ __forceinline__ float min_x86(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float max_x86(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ double min_x86(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double max_x86(double a, double b) { return a > b ? a : b; }

} // namespace ekm
