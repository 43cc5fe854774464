/* Virtual-call dispatch check (reference: include/enoki/array_call.h:124-193, tests/call.cpp):
   the same polymorphic "shader" classes are called through a pointer array on the reference CPU path
   (DynamicArray<Packet<..,8>>) and on this backend (CUDAArray<Base *> -> CUDAArray::partition_() ->
   ek_partition -> per-instance gather / scatter).  Results have to agree bit for bit: each instance evaluates the
   same elementwise expression on the lanes that refer to it.  Written from scratch. */
#include <enoki/cuda.h>
#include <enoki/dynamic.h>
#include <enoki/array.h>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace enoki;

template <typename Float> struct Shader {
    using Mask = mask_t<Float>;
    virtual ~Shader() = default;
    virtual Float eval(const Float &x, const Float &y, Mask active) const = 0;
    virtual void accumulate(const Float &x, Mask active) = 0;
};
template <typename Float> struct Scale : Shader<Float> {
    using Mask = mask_t<Float>;
    float k; Float sum = Float(0.f);
    Scale(float k) : k(k) { }
    Float eval(const Float &x, const Float &y, Mask) const override { return fmadd(x, Float(k), y); }
    void accumulate(const Float &x, Mask active) override { sum = sum + hsum(select(active, x, Float(0.f))); }
};
template <typename Float> struct Wave : Shader<Float> {
    using Mask = mask_t<Float>;
    Float total = Float(0.f);
    Float eval(const Float &x, const Float &y, Mask) const override { return sin(x) * y + sqrt(abs(y)); }
    void accumulate(const Float &x, Mask active) override { total = total + hsum(select(active, x * x, Float(0.f))); }
};

using FloatX = DynamicArray<Packet<float, 8>>;
using FloatC = CUDAArray<float>;
using ShaderX = Shader<FloatX>;
using ShaderC = Shader<FloatC>;

ENOKI_CALL_SUPPORT_BEGIN(ShaderX)
ENOKI_CALL_SUPPORT_METHOD(eval)
ENOKI_CALL_SUPPORT_METHOD(accumulate)
ENOKI_CALL_SUPPORT_END(ShaderX)

ENOKI_CALL_SUPPORT_BEGIN(ShaderC)
ENOKI_CALL_SUPPORT_METHOD(eval)
ENOKI_CALL_SUPPORT_METHOD(accumulate)
ENOKI_CALL_SUPPORT_END(ShaderC)

template <typename Float, typename PtrArray>
void run(size_t n, std::vector<float> &out, float sums[2]) {
    using ShaderT = Shader<Float>;
    using UInt = uint32_array_t<Float>;
    Scale<Float> s0(1.5f), s1(-0.25f);
    Wave<Float> w0;
    ShaderT *table[4] = { &s0, &w0, nullptr, &s1 };
    /* instance of element i: a fixed pseudo-random pattern (with null entries, which dispatch must skip) */
    std::vector<ShaderT *> host(n);
    uint32_t st = 17u;
    for (size_t i = 0; i < n; ++i) { st = st * 1664525u + 1013904223u; host[i] = table[(st >> 13) & 3u]; }
    PtrArray ptrs = PtrArray::copy(host.data(), n);
    /* (not linspace(): the reference's CPU and CUDA backends evaluate it differently, dynamic.h:923-938 vs cuda.h:655-663) */
    Float fi = Float(arange<UInt>(n));
    Float x = fmadd(fi, Float(6.f / float(n)), Float(-3.f)), y = fmadd(fi, Float(2.f / float(n)), Float(0.5f));
    Float r = ptrs->eval(x, y, true);
    ptrs->accumulate(r, true);
    out.resize(n);
    if constexpr (is_cuda_array_v<Float>) { r.eval(); cuda_memcpy_from_device(out.data(), r.data(), n * 4); }
    else memcpy(out.data(), r.data(), n * 4);
    sums[0] = (s0.sum + s1.sum).coeff(0);
    sums[1] = w0.total.coeff(0);
}

int main(int argc, char **argv) {
    size_t n = argc > 1 ? (size_t) atoll(argv[1]) : 100003;
    if (ek_device_count() == 0) { fprintf(stderr, "no CUDA device\n"); return 2; }
    std::vector<float> a, b; float sa[2], sb[2];
    run<FloatX, DynamicArray<Packet<ShaderX *, 8>>>(n, a, sa);
    run<FloatC, CUDAArray<ShaderC *>>(n, b, sb);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) if (memcmp(&a[i], &b[i], 4) != 0) { if (bad < 5) printf("  [%zu] cpu %.9g gpu %.9g\n", i, a[i], b[i]); ++bad; }
    printf("call_check: n=%zu, %zu mismatching results; accumulators cpu (%.7g, %.7g) gpu (%.7g, %.7g)\n", n, bad, sa[0], sa[1], sb[0], sb[1]);
    int fail = bad != 0;
    for (int k = 0; k < 2; ++k) if (std::fabs(sa[k] - sb[k]) > 2e-5f * std::max(1.f, std::fabs(sa[k]))) fail = 1;
    printf(fail ? "call_check: FAILED\n" : "call_check: all checks passed\n");
    return fail;
}
