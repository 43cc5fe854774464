/*
 * ek_adjoint.cu -- level-batched reverse-mode adjoint sweep (sm_100a).
 *
 * Replaces the per-edge array passes of Tape::backward()
 * (src/autodiff/autodiff.cpp:863-888):
 *     source.grad = safe_fmadd(edge.weight, target.grad, source.grad)
 * The host (ek_tape.cpp) assigns every reachable node a level (longest distance
 * from the root over out-edges) and emits, per level, one "job" per source node:
 * the list of its out-edges (weight, target gradient) ordered by descending
 * target id -- the reference's accumulation order, so fp32 results are
 * bit-identical to the CPU tape.  One launch per level: every job is cut into
 * chunks of CHUNK elements, CTAs grid-stride over (job, chunk) pairs, each thread
 * owns 4 consecutive elements (128-bit coalesced loads of weight and adjoint,
 * zero-guarded fma, one 128-bit store).  Algorithmic traffic: every weight once,
 * every adjoint written once and read once per out-edge.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "ek_adjoint.h"
#include "ek_math.cuh"

namespace {

template <typename T> struct Vec4;
template <> struct Vec4<float>  { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T> __device__ __forceinline__ T ld_scalar(const EkAdjTerm &t, bool weight) {
    uint32_t kind = weight ? (t.flags & 3u) : ((t.flags >> 2) & 3u);
    uint64_t bits = weight ? t.w : t.g;
    if (kind == EK_ADJ_IMM) {
        if (sizeof(T) == 4) return (T) __uint_as_float((uint32_t) bits);
        return (T) __longlong_as_double((long long) bits);
    }
    return __ldg(reinterpret_cast<const T *>(bits));
}

template <typename T>
__global__ void __launch_bounds__(256)
ek_adjoint_kernel(const EkAdjJob *__restrict__ jobs, const EkAdjTerm *__restrict__ terms,
                  const uint32_t *__restrict__ chunk_start, uint32_t n_jobs, uint32_t n_chunks) {
    constexpr uint32_t CHUNK = EK_ADJ_CHUNK;
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        /* binary search: last job whose first chunk is <= c */
        uint32_t lo = 0, hi = n_jobs - 1;
        while (lo < hi) {
            uint32_t mid = (lo + hi + 1) >> 1;
            if (__ldg(chunk_start + mid) <= c) lo = mid; else hi = mid - 1;
        }
        const EkAdjJob job = jobs[lo];
        const uint32_t base = (c - __ldg(chunk_start + lo)) * CHUNK;
        const uint32_t n = job.size;
        T *dst = reinterpret_cast<T *>(job.dst);
        const EkAdjTerm *tt = terms + job.first_term;

        constexpr int PER = 16 / sizeof(T);                  /* elements per 128-bit access */
        for (uint32_t e0 = base + threadIdx.x * PER; e0 < min(base + CHUNK, n); e0 += blockDim.x * PER) {
            const bool full = e0 + PER <= n && job.aligned;
            T acc[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) acc[j] = (T) 0;
            for (uint32_t k = 0; k < job.n_terms; ++k) {
                const EkAdjTerm t = tt[k];
                T w[PER], g[PER];
                const uint32_t wk = t.flags & 3u, gk = (t.flags >> 2) & 3u;
                if (wk == EK_ADJ_ARRAY) {
                    const T *p = reinterpret_cast<const T *>(t.w) + e0;
                    if (full) {
                        if (sizeof(T) == 4) { float4 v = __ldcs(reinterpret_cast<const float4 *>(p)); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
                        else { double2 v = __ldcs(reinterpret_cast<const double2 *>(p)); w[0] = v.x; w[1] = v.y; }
                    } else {
#pragma unroll
                        for (int j = 0; j < PER; ++j) w[j] = e0 + j < n ? p[j] : (T) 0;
                    }
                } else {
                    T s = ld_scalar<T>(t, true);
#pragma unroll
                    for (int j = 0; j < PER; ++j) w[j] = s;
                }
                if (gk == EK_ADJ_ARRAY) {
                    const T *p = reinterpret_cast<const T *>(t.g) + e0;
                    if (full) {
                        if (sizeof(T) == 4) { float4 v = __ldg(reinterpret_cast<const float4 *>(p)); g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w; }
                        else { double2 v = __ldg(reinterpret_cast<const double2 *>(p)); g[0] = v.x; g[1] = v.y; }
                    } else {
#pragma unroll
                        for (int j = 0; j < PER; ++j) g[j] = e0 + j < n ? p[j] : (T) 0;
                    }
                } else {
                    T s = ld_scalar<T>(t, false);
#pragma unroll
                    for (int j = 0; j < PER; ++j) g[j] = s;
                }
                /* autodiff.cpp:873-876: first contribution safe_mul, then safe_fmadd */
                if (k == 0) {
#pragma unroll
                    for (int j = 0; j < PER; ++j) acc[j] = ekm::mul_nz(w[j], g[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < PER; ++j) acc[j] = ekm::fma_nz(w[j], g[j], acc[j]);
                }
            }
            if (full) {
                if (sizeof(T) == 4) *reinterpret_cast<float4 *>(dst + e0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                else *reinterpret_cast<double2 *>(dst + e0) = make_double2(acc[0], acc[1]);
            } else {
#pragma unroll
                for (int j = 0; j < PER; ++j) if (e0 + j < n) dst[e0 + j] = acc[j];
            }
        }
    }
}

} // namespace

cudaError_t ek_launch_adjoint(bool f64, const EkAdjJob *jobs, const EkAdjTerm *terms,
                              const uint32_t *chunk_start, uint32_t n_jobs, uint32_t n_chunks,
                              unsigned grid, cudaStream_t stream) {
    if (n_jobs == 0 || n_chunks == 0) return cudaSuccess;
    if (f64) ek_adjoint_kernel<double><<<grid, 256, 0, stream>>>(jobs, terms, chunk_start, n_jobs, n_chunks);
    else     ek_adjoint_kernel<float><<<grid, 256, 0, stream>>>(jobs, terms, chunk_start, n_jobs, n_chunks);
    return cudaGetLastError();
}
