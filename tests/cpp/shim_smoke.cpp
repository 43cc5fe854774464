/* Compile-and-run check of the C++ drop-in boundary: the reference's generic headers + this repo's
   <enoki/cuda.h> / <enoki/autodiff_b200.h>, with the type aliases of tests/autodiff.cpp swapped to
   CUDA types (the way the reference author tested the GPU path, tests/autodiff.cpp:447-448).
   Built by tests/cpp/Makefile in the dev container (needs /root/reference/include); the binary
   travels to the GPU box.  Exit code 0 = all checks passed. */
#include <enoki/autodiff.h>
#include <enoki/cuda.h>
#include <enoki/random.h>
#include <cstdio>
#include <cmath>
#include <vector>

using namespace enoki;
using FloatC = CUDAArray<float>;
using UInt32C = CUDAArray<uint32_t>;
using FloatD = DiffArray<FloatC>;
using UInt32D = DiffArray<UInt32C>;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); ++failures; } } while (0)

static bool close(float a, float b, float tol = 1e-5f) { return std::fabs(a - b) <= tol * std::max(1.f, std::fabs(b)); }

int main() {
    if (ek_device_count() == 0) { fprintf(stderr, "no CUDA device\n"); return 2; }

    /* --- evaluator through the router: operators, math, horizontal ops --- */
    {
        FloatC x = linspace<FloatC>(0.f, 1.f, 1001);
        FloatC y = x * x + sin(x) * 2.f;
        float s = hsum(y).coeff(0);
        double ref = 0;
        for (int i = 0; i < 1001; ++i) { double v = i / 1000.0; ref += v * v + 2 * std::sin(v); }
        CHECK(close(s, (float) ref, 1e-5f));
        CHECK(y.size() == 1001);
        CHECK(close(y.coeff(1000), 1.f + 2.f * std::sin(1.f)));
        CHECK(all(x >= 0.f) && any(x > 0.5f) && count(x > 0.5f) == 500);
        /* functions WITHOUT a CUDAArray member are traced op-by-op from array_math.h */
        FloatC t = tan(x), a = atan2(x, x + 1.f);
        CHECK(close(t.coeff(500), std::tan(0.5f), 2e-6f));
        CHECK(close(a.coeff(1000), std::atan2(1.f, 2.f), 2e-6f));
    }
    /* --- gather / scatter_add with array operands (array_struct.h:8-123) --- */
    {
        UInt32C idx = arange<UInt32C>(1000);
        FloatC table = linspace<FloatC>(0.f, 9.f, 10);
        FloatC g = gather<FloatC>(table, idx % 10u);
        CHECK(close(hsum(g).coeff(0), 4500.f));
        FloatC bins = zero<FloatC>(10);
        scatter_add(bins, FloatC(1.f), idx % 10u);
        CHECK(close(bins.coeff(3), 100.f));
    }
    /* --- PCG32 traced through CUDAArray<uint64_t> (random.h) --- */
    {
        using RNG = PCG32<FloatC>;
        RNG rng(PCG32_DEFAULT_STATE, arange<CUDAArray<uint64_t>>(4));
        CUDAArray<uint32_t> v = rng.next_uint32();
        PCG32<float> r0(PCG32_DEFAULT_STATE, 0), r3(PCG32_DEFAULT_STATE, 3);
        CHECK(v.coeff(0) == r0.next_uint32());
        CHECK(v.coeff(3) == r3.next_uint32());
    }
    /* --- DiffArray<CUDAArray<float>>: tests/autodiff.cpp style checks --- */
    {
        FloatD x = linspace<FloatC>(0.1f, 2.f, 64);
        set_requires_gradient(x);
        FloatD y = sin(x) * exp(x) + sqrt(x) / (x * x + 1.f);
        FloatD loss = hsum(y);
        backward(loss);
        FloatC g = gradient(x);
        for (int i : { 0, 17, 63 }) {
            float v = 0.1f + 1.9f * i / 63.f;
            double d = std::exp(v) * (std::sin(v) + std::cos(v)) + (0.5 / std::sqrt(v) * (v * v + 1) - std::sqrt(v) * 2 * v) / ((v * v + 1) * (v * v + 1));
            CHECK(close(g.coeff(i), (float) d, 1e-4f));
        }
    }
    {   /* scalar leaf broadcast into a wide expression (tests/autodiff.cpp:533-548) */
        FloatD a = 2.f;
        set_requires_gradient(a);
        FloatD x = linspace<FloatC>(0.f, 1.f, 10);
        FloatD y = hsum(a * x * x);
        backward(y);
        float want = 0; for (int i = 0; i < 10; ++i) want += (i / 9.f) * (i / 9.f);
        CHECK(close(gradient(a).coeff(0), want, 1e-5f));
    }
    {   /* gather under AD: gradient is a scatter_add (autodiff.cpp:384-398) */
        FloatD src = linspace<FloatC>(1.f, 4.f, 4);
        set_requires_gradient(src);
        UInt32D idx = UInt32C(3u, 3u, 0u, 1u, 3u);
        FloatD y = gather<FloatD>(src, idx);
        backward(hsum(y * y));
        FloatC g = gradient(src);
        CHECK(close(g.coeff(0), 2.f) && close(g.coeff(1), 4.f) && close(g.coeff(2), 0.f) && close(g.coeff(3), 24.f));
    }
    cuda_sync();
    if (failures == 0) printf("shim_smoke: all checks passed\n");
    return failures == 0 ? 0 : 1;
}
