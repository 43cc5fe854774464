/*
 * cuda_shim.h -- just enough of the CUDA device programming model to compile enoki_b200/csrc/ek_sweep_fast.cu as
 * HOST code and run it with one POSIX thread per CUDA thread (tests/cpu_kernel/emu_fast.cpp).  TEST INFRASTRUCTURE:
 * the repository had no GPU access while the fast kernel was written, so its control flow, operand addressing, case
 * bodies, private-bin arithmetic and reduction epilogue are executed on the CPU against the same expectations as the
 * GPU tests.  What the shim cannot show: timing, memory-model races, and the PTX-only paths it replaces (TMA bulk
 * copies -> memcpy, mbarrier -> a phase word, packed fma.rn.f32x2 -> two fmaf, rsqrt.approx -> 1/sqrtf).
 */
#pragma once
#define EK_HOST_EMU 1
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(x) alignas(x)

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
struct EmuDim { unsigned x = 1, y = 1, z = 1; };

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidConfiguration = 9 };

namespace emu {
extern thread_local EmuDim tid_;            /* threadIdx of the calling POSIX thread */
extern EmuDim bid_, bdim_, gdim_;           /* CTAs run one after the other */
extern uint8_t *smem_;                      /* dynamic shared memory of the running CTA */
constexpr uint32_t SMEM_WINDOW = 0x400u;    /* shared-space addresses start here, like on the device */
void cta_barrier();
/* warp-level rendezvous of the lanes in `mask`; `mine` is published, the return value is the array of all 32 values */
const uint64_t *warp_exchange(unsigned mask, uint64_t mine);
void warp_release(unsigned mask);
[[noreturn]] void trap(const char *why);
}
#define threadIdx (emu::tid_)
#define blockIdx (emu::bid_)
#define blockDim (emu::bdim_)
#define gridDim (emu::gdim_)

inline void __syncthreads() { emu::cta_barrier(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __trap() { emu::trap("__trap()"); }
inline long long clock64() { return 0; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t) ((const uint8_t *) p - emu::smem_) + emu::SMEM_WINDOW; }

/* ---- bit casts / conversions (round-to-nearest-even is the host default; no x87, no contraction: see the Makefile) */
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
inline double __hiloint2double(int hi, int lo) { uint64_t v = ((uint64_t) (uint32_t) hi << 32) | (uint32_t) lo; double d; memcpy(&d, &v, 8); return d; }
inline int __double2loint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int) (uint32_t) v; }
inline int __double2hiint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int) (uint32_t) (v >> 32); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __frcp_rn(float a) { return 1.f / a; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __int2float_rn(int v) { return (float) v; }
inline float __uint2float_rn(uint32_t v) { return (float) v; }
inline double __ll2double_rn(long long v) { return (double) v; }
inline int __float2int_rz(float x) { return (int) truncf(x); }
inline int __float2int_rd(float x) { return (int) floorf(x); }
inline int __float2int_ru(float x) { return (int) ceilf(x); }
inline int __float2int_rn(float x) { return (int) rintf(x); }
inline uint32_t __float2uint_rz(float x) { return (uint32_t) truncf(x); }
inline long long __float2ll_rz(float x) { return (long long) truncf(x); }
inline long long __float2ll_rd(float x) { return (long long) floorf(x); }
inline long long __float2ll_ru(float x) { return (long long) ceilf(x); }
inline long long __float2ll_rn(float x) { return (long long) rintf(x); }
inline int __double2int_rz(double x) { return (int) trunc(x); }
inline int __double2int_rd(double x) { return (int) floor(x); }
inline int __double2int_ru(double x) { return (int) ceil(x); }
inline int __double2int_rn(double x) { return (int) rint(x); }
inline long long __double2ll_rz(double x) { return (long long) trunc(x); }
inline long long __double2ll_rd(double x) { return (long long) floor(x); }
inline long long __double2ll_ru(double x) { return (long long) ceil(x); }
inline long long __double2ll_rn(double x) { return (long long) rint(x); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned) v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline int __mulhi(int a, int b) { return (int) (((long long) a * (long long) b) >> 32); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((unsigned long long) a * b) >> 32); }
inline int __ffs(unsigned v) { return __builtin_ffs((int) v); }
using std::max;
using std::min;

/* ---- memory */
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
inline float atomicAdd(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    do { float f; memcpy(&f, &old, 4); f += v; memcpy(&want, &f, 4); } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    float r; memcpy(&r, &old, 4); return r;
}
inline double atomicAdd(double *p, double v) {
    uint64_t *u = reinterpret_cast<uint64_t *>(p), old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
    do { double f; memcpy(&f, &old, 8); f += v; memcpy(&want, &f, 8); } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    double r; memcpy(&r, &old, 8); return r;
}
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

/* ---- warp primitives: all lanes named in `mask` call together (that is the CUDA contract as well) */
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int m) {
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    const uint64_t *all = emu::warp_exchange(mask, bits);
    uint64_t got = all[(threadIdx.x & 31u) ^ (unsigned) m];
    emu::warp_release(mask);
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src) {
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    const uint64_t *all = emu::warp_exchange(mask, bits);
    uint64_t got = all[(unsigned) src & 31u];
    emu::warp_release(mask);
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
inline unsigned __ballot_sync(unsigned mask, bool pred) {
    const uint64_t *all = emu::warp_exchange(mask, pred ? 1u : 0u);
    unsigned r = 0; for (unsigned l = 0; l < 32; ++l) if (((mask >> l) & 1u) && all[l]) r |= 1u << l;
    emu::warp_release(mask);
    return r;
}
inline bool __all_sync(unsigned mask, bool pred) { return (__ballot_sync(mask, pred) & mask) == mask; }
inline bool __any_sync(unsigned mask, bool pred) { return (__ballot_sync(mask, pred) & mask) != 0u; }
inline unsigned __match_any_sync(unsigned mask, unsigned long long v) {
    const uint64_t *all = emu::warp_exchange(mask, (uint64_t) v);
    unsigned r = 0; for (unsigned l = 0; l < 32; ++l) if (((mask >> l) & 1u) && all[l] == (uint64_t) v) r |= 1u << l;
    emu::warp_release(mask);
    return r;
}
