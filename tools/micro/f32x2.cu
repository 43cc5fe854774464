// Microbenchmark: issue rate of FFMA vs FFMA2 (packed f32x2) on sm_100a, and a rounding check that
// mul.rn.f32x2 followed by add.rn.f32x2 is NOT contracted into one fused multiply-add.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE> __global__ void bench(float *out, int iters, float s) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], s, 0.25f);
    } else {
        u64 p[8], ss = pk(s, s), cc = pk(0.25f, 0.25f);
        for (int i = 0; i < 8; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, cc);
        for (int i = 0; i < 8; ++i) upk(p[i], x[2 * i], x[2 * i + 1]);
    }
    float acc = 0; for (int i = 0; i < 16; ++i) acc += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void rounding(const float *a, const float *b, const float *c, uint32_t *bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    u64 r = add2(mul2(pk(a[2 * i], a[2 * i + 1]), pk(b[2 * i], b[2 * i + 1])), pk(c[2 * i], c[2 * i + 1]));
    float lo, hi; upk(r, lo, hi);
    float e0 = __fadd_rn(__fmul_rn(a[2 * i], b[2 * i]), c[2 * i]), e1 = __fadd_rn(__fmul_rn(a[2 * i + 1], b[2 * i + 1]), c[2 * i + 1]);
    if (__float_as_uint(lo) != __float_as_uint(e0) || __float_as_uint(hi) != __float_as_uint(e1)) atomicAdd(bad, 1u);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        int iters = 20000;
        cudaEventRecord(e0);
        if (mode == 0) bench<0><<<148 * 8, 256>>>(out, iters, 0.999f); else bench<1><<<148 * 8, 256>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * iters * 148 * 8 * 256;
        printf("%s: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, flops / ms * 1e-9);
    }
    int n = 1 << 20; float *a, *b, *c; uint32_t *bad;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&bad, 4);
    uint32_t st = 12345; auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (float) ((st >> 8) * (1.0 / 16777216.0) * 8.0 - 4.0); };
    for (int i = 0; i < n; ++i) { a[i] = rnd(); b[i] = rnd(); c[i] = -a[i] * b[i] * (1.0f + 1e-7f * (i & 7)); }
    *bad = 0; rounding<<<n / 2 / 256, 256>>>(a, b, c, bad, n); cudaDeviceSynchronize();
    printf("mul.rn.f32x2 + add.rn.f32x2 vs separate roundings: %u mismatches of %d pairs (0 = not contracted)\n", *bad, n / 2);
    return 0;
}
