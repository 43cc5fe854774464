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
#ifndef EK_HOST_EMU
#include <cuda_runtime.h>
#endif

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

/* ---------------- packed single precision (Blackwell FFMA2) ----------------
   sm_100 executes two independent IEEE fp32 fused multiply-adds per instruction (PTX fma.rn.f32x2, SASS FFMA2).
   The interpreter is bound by instruction issue, not by the fp32 pipe, so the polynomial bodies below process
   two elements per instruction.  Every packed operation is an FMA with exact identities, so each half is
   rounded exactly like the scalar code above:
       a * b  = fma(a, b, -0)        (the sum with -0 is exact, also for +-0 products)
       a + b  = fma(a, 1, b)         a - b = fma(b, -1, a)
   (mul.rn.f32x2 followed by add.rn.f32x2 is NOT used: ptxas 12.9 contracts that pair into one FFMA2 --
   tools/micro/f32x2.cu -- which would change the rounding.) */
struct f2 { float x, y; };
__device__ __forceinline__ f2 mk2(float v) { return f2{ v, v }; }
#ifdef EK_HOST_EMU          /* tests/cpu_kernel: both halves round like the scalar fma */
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) { return f2{ fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y) }; }
#else
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("{\n .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd;\n}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#endif
__device__ __forceinline__ f2 fmul2(f2 a, f2 b) { return ffma2(a, b, mk2(-0.f)); }
__device__ __forceinline__ f2 fadd2(f2 a, f2 b) { return ffma2(a, mk2(1.f), b); }
__device__ __forceinline__ f2 fsub2(f2 a, f2 b) { return ffma2(b, mk2(-1.f), a); }
__device__ __forceinline__ f2 ffma2(f2 a, float b, float c) { return ffma2(a, mk2(b), mk2(c)); }
__device__ __forceinline__ f2 ffma2(f2 a, float b, f2 c) { return ffma2(a, mk2(b), c); }
__device__ __forceinline__ f2 fmul2(f2 a, float b) { return fmul2(a, mk2(b)); }

__device__ __forceinline__ f2 poly2(f2 x, float c0, float c1, float c2) {
    f2 x2 = fmul2(x, x);
    return ffma2(x2, c2, ffma2(x, c1, c0));
}
__device__ __forceinline__ f2 poly5(f2 x, float c0, float c1, float c2, float c3, float c4, float c5) {
    f2 x2 = fmul2(x, x), x4 = fmul2(x2, x2);
    return ffma2(x2, ffma2(x, c3, c2), ffma2(x4, ffma2(x, c5, c4), ffma2(x, c1, c0)));
}
/* two elements of sincos_approx (same steps as sincos<> above) */
template <bool Sin, bool Cos>
__device__ __forceinline__ void sincos2(f2 x, f2 &s_out, f2 &c_out) {
    f2 xa = { fabsf(x.x), fabsf(x.y) };
    f2 q = fmul2(xa, 1.2732395447351626862f);
    int32_t j0 = (cvtt_sat(q.x) + 1) & ~1, j1 = (cvtt_sat(q.y) + 1) & ~1;
    f2 y = { __int2float_rn(j0), __int2float_rn(j1) };
    y = fsub2(fsub2(fsub2(xa, fmul2(y, 0.78515625f)), fmul2(y, 2.4187564849853515625e-4f)),
              fmul2(y, 3.77489497744594108e-8f));
    f2 z = fmul2(y, y);
    if (xa.x == __int_as_float(0x7f800000)) z.x = fbits(0xffffffffu);
    if (xa.y == __int_as_float(0x7f800000)) z.y = fbits(0xffffffffu);
    f2 s = fmul2(poly2(z, -1.6666654611e-1f, 8.3321608736e-3f, -1.9515295891e-4f), z);
    f2 c = fmul2(poly2(z, 4.166664568298827e-2f, -1.388731625493765e-3f, 2.443315711809948e-5f), z);
    s = ffma2(s, y, y);
    c = ffma2(c, z, ffma2(z, -0.5f, 1.f));
    bool p0 = (j0 & 2) == 0, p1 = (j1 & 2) == 0;
    if (Sin) {
        s_out.x = fbits(ubits(p0 ? s.x : c.x) ^ ((((uint32_t) j0 << 29) ^ ubits(x.x)) & 0x80000000u));
        s_out.y = fbits(ubits(p1 ? s.y : c.y) ^ ((((uint32_t) j1 << 29) ^ ubits(x.y)) & 0x80000000u));
    }
    if (Cos) {
        c_out.x = fbits(ubits(p0 ? c.x : s.x) ^ (((uint32_t) (~(j0 - 2)) << 29) & 0x80000000u));
        c_out.y = fbits(ubits(p1 ? c.y : s.y) ^ (((uint32_t) (~(j1 - 2)) << 29) & 0x80000000u));
    }
}
__device__ __forceinline__ f2 sin_f32x2(f2 x) { f2 s, c; sincos2<true, false>(x, s, c); return s; }
__device__ __forceinline__ f2 cos_f32x2(f2 x) { f2 s, c; sincos2<false, true>(x, s, c); return c; }
/* two elements of exp_f32 */
__device__ __forceinline__ f2 exp_f32x2(f2 x) {
    const float max_range = +88.3762588501f, min_range = -88.3762588501f;
    f2 t = ffma2(x, 1.4426950408889634073599f, 0.5f);
    f2 n = { floorf(t.x), floorf(t.y) };
    f2 nn = { -n.x, -n.y };
    f2 xr = ffma2(nn, 0.693359375f, x);
    xr = ffma2(nn, -2.12194440e-4f, xr);
    f2 z = poly5(xr, 5.0000001201e-1f, 1.6666665459e-1f, 4.1665795894e-2f,
                     8.3334519073e-3f, 1.3981999507e-3f, 1.9875691500e-4f);
    z = ffma2(z, fmul2(xr, xr), fadd2(xr, mk2(1.f)));
    f2 sc = { fbits((uint32_t) (cvtt_sat(n.x) + 0x7f) << 23), fbits((uint32_t) (cvtt_sat(n.y) + 0x7f) << 23) };
    f2 r = fmul2(z, sc);
    r.x = x.x > max_range ? __int_as_float(0x7f800000) : (x.x < min_range ? 0.f : r.x);
    r.y = x.y > max_range ? __int_as_float(0x7f800000) : (x.y < min_range ? 0.f : r.y);
    return r;
}

/* ---------------- double precision (array_math.h, the `!Single` branches) ---------------- */
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double dbits(uint64_t u) { return __longlong_as_double((long long) u); }
__device__ __forceinline__ uint64_t ubits(double d) { return (uint64_t) __double_as_longlong(d); }
/* x86 cvttpd2qq / scalar cvttsd2si semantics */
__device__ __forceinline__ long long cvtt_f64_i64(double x) {
    return (x >= -9223372036854775808.0 && x < 9223372036854775808.0) ? __double2ll_rz(x) : (long long) 0x8000000000000000ull;
}
__device__ __forceinline__ double poly2(double x, double c0, double c1, double c2) {
    double x2 = dmul(x, x);
    return dfma(x2, c2, dfma(x, c1, c0));
}
__device__ __forceinline__ double poly3(double x, double c0, double c1, double c2, double c3) {
    double x2 = dmul(x, x);
    return dfma(x2, dfma(x, c3, c2), dfma(x, c1, c0));
}
__device__ __forceinline__ double poly5(double x, double c0, double c1, double c2, double c3, double c4, double c5) {
    double x2 = dmul(x, x), x4 = dmul(x2, x2);
    return dfma(x2, dfma(x, c3, c2), dfma(x4, dfma(x, c5, c4), dfma(x, c1, c0)));
}

/* array_math.h:261-367 sincos_approx, double branch */
template <bool Sin, bool Cos>
__device__ __forceinline__ void sincos(double x, double &s_out, double &c_out) {
    double xa = fabs(x);
    long long j = cvtt_f64_i64(dmul(xa, 1.2732395447351626862));
    j = (j + 1) & 0xfffffffell;   /* the reference masks with Int(~1u): a 32-bit constant widened to 64 */
    double y = __ll2double_rn(j);
    uint64_t sign_sin = ((uint64_t) j << 61) ^ ubits(x);
    uint64_t sign_cos = (uint64_t) (~(j - 2)) << 61;
    y = dsub(dsub(dsub(xa, dmul(y, 7.85398125648498535156e-1)), dmul(y, 3.77489470793079817668e-8)),
             dmul(y, 2.69515142907905952645e-15));
    double z = dmul(y, y);
    if (xa == dbits(0x7ff0000000000000ull)) z = dbits(~0ull);
    double s = dmul(poly5(z, -1.66666666666666307295e-1, 8.33333333332211858878e-3, -1.98412698295895385996e-4,
                             2.75573136213857245213e-6, -2.50507477628578072866e-8, 1.58962301576546568060e-10), z);
    double c = dmul(poly5(z, 4.16666666666665929218e-2, -1.38888888888730564116e-3, 2.48015872888517045348e-5,
                             -2.75573141792967388112e-7, 2.08757008419747316778e-9, -1.13585365213876817300e-11), z);
    s = dfma(s, y, y);
    c = dfma(c, z, dfma(z, -0.5, 1.0));
    bool polymask = (j & 2) == 0;
    if (Sin) s_out = dbits(ubits(polymask ? s : c) ^ (sign_sin & 0x8000000000000000ull));
    if (Cos) c_out = dbits(ubits(polymask ? c : s) ^ (sign_cos & 0x8000000000000000ull));
}
__device__ __forceinline__ double sin_f64(double x) { double s, c; sincos<true, false>(x, s, c); return s; }
__device__ __forceinline__ double cos_f64(double x) { double s, c; sincos<false, true>(x, s, c); return c; }

/* array_math.h:711-776 exp, double branch */
__device__ __forceinline__ double exp_f64(double x) {
    const double max_range = +7.0943613930310391424428e2, min_range = -7.0943613930310391424428e2;
    bool overflow = x > max_range, underflow = x < min_range;
    double n = floor(dfma(1.4426950408889634073599, x, 0.5));
    double xr = dfma(-n, 6.93145751953125e-1, x);
    xr = dfma(-n, 1.42860682030941723212e-6, xr);
    double z = dmul(xr, xr);
    double p = dmul(poly2(z, 9.99999999999999999910e-1, 3.02994407707441961300e-2, 1.26177193074810590878e-4), xr);
    double q = poly3(z, 2.00000000000000000009e0, 2.27265548208155028766e-1, 2.52448340349684104192e-3, 3.00198505138664455042e-6);
    double pq = __ddiv_rn(p, dsub(q, p));
    z = dadd(dadd(pq, pq), 1.0);
    uint64_t scale = (uint64_t) (cvtt_f64_i64(n) + 0x3ff) << 52;
    double r = dmul(z, dbits(scale));
    return overflow ? dbits(0x7ff0000000000000ull) : (underflow ? 0.0 : r);
}

/* array_math.h:778-898 log, double branch (both sub-branches evaluated and selected, as the IsCuda path does) */
__device__ __forceinline__ double log_f64(double x) {
    bool valid = x >= 0.0;
    uint64_t xi = ubits(x);
    uint64_t exponent_bits = xi & 0x7ff0000000000000ull;
    bool is_normal = (x != 0.0) && (exponent_bits != 0x7ff0000000000000ull);
    long long exponent_i = (long long) (exponent_bits >> 52) - 0x3ff;
    uint64_t mantissa = (xi & ~0x7ff0000000000000ull) | 0x3fe0000000000000ull;
    double xm = dbits(is_normal ? mantissa : xi);
    double e = __ll2double_rn(is_normal ? exponent_i : 0);
    bool e_big = fabs(e) > 2.0;
    bool ge = xm >= 0.70710678118654752440;
    if (ge) e = dadd(e, 1.0);
    /* big exponent: log(x) = z + z^3 P(z)/Q(z), z = 2(x-1)/(x+1) */
    double zb = dsub(xm, 0.5);
    if (ge) zb = dsub(zb, 0.5);
    double yb = dadd(dmul(0.5, ge ? xm : zb), 0.5);
    double x2b = __ddiv_rn(zb, yb);
    double zz = dmul(x2b, x2b);
    zz = dmul(x2b, __ddiv_rn(dmul(zz, poly2(zz, -6.41409952958715622951e1, 1.63866645699558079767e1, -7.89580278884799154124e-1)),
                             poly3(zz, -7.69691943550460008604e2, 3.12093766372244180303e2, -3.56722798256324312549e1, 1.00000000000000000000e0)));
    double r_big = dadd(dfma(-e, 2.121944400546905827679e-4, zz), x2b);
    /* small exponent: log(1+x) = x - x^2/2 + x^3 P(x)/Q(x) */
    double x2s = dsub(ge ? xm : dadd(xm, xm), 1.0);
    double zs = dmul(x2s, x2s);
    double ys = dmul(x2s, __ddiv_rn(dmul(zs, poly5(x2s, 7.70838733755885391666e0, 1.79368678507819816313e1, 1.44989225341610930846e1,
                                                     4.70579119878881725854e0, 4.97494994976747001425e-1, 1.01875663804580931796e-4)),
                                   poly5(x2s, 2.31251620126765340583e1, 7.11544750618563894466e1, 8.29875266912776603211e1,
                                         4.52279145837532221105e1, 1.12873587189167450590e1, 1.00000000000000000000e0)));
    ys = dfma(-e, 2.121944400546905827679e-4, ys);
    double r_small = dadd(x2s, dfma(-0.5, zs, ys));
    double r = e_big ? r_big : r_small;
    r = dfma(e, 0.693359375, r);
    if (x == dbits(0x7ff0000000000000ull)) r = dbits(0x7ff0000000000000ull);
    if (x == 0.0) r = dbits(0xfff0000000000000ull);
    return valid ? r : dbits(~0ull);
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
