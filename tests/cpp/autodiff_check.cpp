/* Tape parity on the C++ API level: every scenario is a template that is run on the reference's CPU
   tape (DiffArray<DynamicArray<Packet<float,8>>>, reference autodiff.cpp from oracle/_ref) and on this
   backend (DiffArray<CUDAArray<float>> through <enoki/cuda.h> + <enoki/autodiff_b200.h>); values and
   gradients must agree.  Covers the op list of SURVEY Appendix B, the special edges
   (gather / scatter / scatter_add, backward and forward mode: the shapes of tests/autodiff.cpp:400-466,609-636
   with their published expected vectors), broadcasting of scalar leaves, forward mode and psum/reverse. */
#include <enoki/autodiff.h>
#include <enoki/cuda.h>
#include <enoki/dynamic.h>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <functional>

using namespace enoki;
using FloatX = DynamicArray<Packet<float, 8>>;
using FloatC = CUDAArray<float>;

template <typename Float> std::vector<float> to_host(const Float &v) {
    std::vector<float> r(v.size());
    if constexpr (is_cuda_array_v<Float>) { if (v.size()) { v.eval(); cuda_memcpy_from_device(r.data(), v.data(), r.size() * 4); } }
    else { for (size_t i = 0; i < r.size(); ++i) r[i] = v.coeff(i); }
    return r;
}
template <typename FloatD> void append(std::vector<float> &out, const FloatD &v) {
    auto h = to_host(detach(v));
    out.insert(out.end(), h.begin(), h.end());
}
template <typename Float> void append_raw(std::vector<float> &out, const Float &v) {
    auto h = to_host(v);
    out.insert(out.end(), h.begin(), h.end());
}

/* ---- scenarios: return concatenated values / gradients ---- */
#define SCENARIO(name) template <typename FloatD> std::vector<float> name()
#define TYPES using Float = std::decay_t<decltype(detach(std::declval<const FloatD &>()))>; using UInt32D = uint32_array_t<FloatD>; (void) sizeof(UInt32D);

SCENARIO(s_arith) { TYPES
    FloatD x = linspace<FloatD>(0.3f, 2.5f, 37), y = linspace<FloatD>(-1.5f, 1.7f, 37);
    set_requires_gradient(x); set_requires_gradient(y);
    FloatD z = fmadd(x, y, x - y) / (x * x + 1.f) + sqrt(x) * rcp(x + 2.f) - rsqrt(x + 0.5f) + abs(y) + max(x, y) * min(x, y);
    backward(hsum(z));
    std::vector<float> r; append(r, z); append_raw(r, gradient(x)); append_raw(r, gradient(y)); return r;
}
SCENARIO(s_trig) { TYPES
    FloatD x = linspace<FloatD>(-1.2f, 1.3f, 41);
    set_requires_gradient(x);
    FloatD z = sin(x) * cos(x) + tan(x * 0.5f) + exp(x) * log(x * x + 1.f) + atan2(x, x * x + 0.5f) + sinh(x) - tanh(x) + asin(x * 0.5f);
    backward(hsum(z));
    std::vector<float> r; append(r, z); append_raw(r, gradient(x)); return r;
}
SCENARIO(s_select) { TYPES
    FloatD x = linspace<FloatD>(-2.f, 2.f, 33);
    set_requires_gradient(x);
    FloatD z = select(x > 0.5f, x * x, -x) + select(x < -1.f, FloatD(3.f), sqrt(abs(x) + 1.f));
    backward(hsum(z * z));
    std::vector<float> r; append(r, z); append_raw(r, gradient(x)); return r;
}
SCENARIO(s_broadcast) { TYPES          /* scalar leaves feeding wide expressions (autodiff.cpp:867-871) */
    FloatD a = 1.5f, b = -0.25f;
    set_requires_gradient(a); set_requires_gradient(b);
    FloatD x = linspace<FloatD>(0.f, 1.f, 50);
    FloatD z = hsum(sin(a * x + b) * a);
    backward(z);
    std::vector<float> r; append(r, z); append_raw(r, gradient(a)); append_raw(r, gradient(b)); return r;
}
SCENARIO(s_hprod) { TYPES
    FloatD x = linspace<FloatD>(0.9f, 1.1f, 9);
    set_requires_gradient(x);
    FloatD z = hprod(x) + hsum(x) * 2.f;
    backward(z);
    std::vector<float> r; append(r, z); append_raw(r, gradient(x)); return r;
}
SCENARIO(s_scatter_add) { TYPES        /* tests/autodiff.cpp:400-431 */
    UInt32D idx1 = arange<UInt32D>(5), idx2 = arange<UInt32D>(4) + 3u;
    FloatD x = linspace<FloatD>(0, 1, 5), y = linspace<FloatD>(1, 2, 4);
    set_requires_gradient(x); set_requires_gradient(y);
    FloatD buf = zero<FloatD>(10);
    scatter_add(buf, x, idx1);
    scatter_add(buf, y, idx2);
    FloatD s = dot(buf, buf);
    backward(s);
    std::vector<float> r; append(r, buf); append_raw(r, gradient(x)); append_raw(r, gradient(y)); return r;
}
SCENARIO(s_scatter) { TYPES            /* tests/autodiff.cpp:433-466 */
    UInt32D idx1 = arange<UInt32D>(5), idx2 = arange<UInt32D>(4) + 3u;
    FloatD x = linspace<FloatD>(0, 1, 5), y = linspace<FloatD>(1, 2, 4);
    set_requires_gradient(x); set_requires_gradient(y);
    FloatD buf = zero<FloatD>(10);
    scatter(buf, x, idx1);
    if constexpr (is_cuda_array_v<FloatD>) cuda_eval();
    scatter(buf, y, idx2);
    FloatD s = dot(buf, buf);
    backward(s);
    std::vector<float> r; append(r, buf); append_raw(r, gradient(x)); append_raw(r, gradient(y)); return r;
}
SCENARIO(s_gather) { TYPES             /* tests/autodiff.cpp:609-617 */
    FloatD x = linspace<FloatD>(-1.f, 1.f, 10);
    set_requires_gradient(x);
    FloatD y = gather<FloatD>(x * x, UInt32D(1, 2, 3));
    backward(hsum(y));
    std::vector<float> r; append(r, y); append_raw(r, gradient(x)); return r;
}
SCENARIO(s_gather_fwd) { TYPES         /* tests/autodiff.cpp:619-626 */
    FloatD x = linspace<FloatD>(-1.f, 1.f, 10);
    set_requires_gradient(x);
    FloatD y = gather<FloatD>(x * x, UInt32D(1, 2, 3));
    forward(x);
    std::vector<float> r; append_raw(r, gradient(y)); return r;
}
SCENARIO(s_scatter_fwd) { TYPES        /* tests/autodiff.cpp:628-636 */
    FloatD x = linspace<FloatD>(-1.f, 1.f, 5);
    set_requires_gradient(x);
    FloatD y = zero<FloatD>(10);
    scatter(y, x * x, arange<UInt32D>(5) + 2);
    forward(x);
    std::vector<float> r; append_raw(r, gradient(y)); return r;
}
SCENARIO(s_forward) { TYPES
    FloatD x = linspace<FloatD>(0.2f, 1.8f, 21);
    set_requires_gradient(x);
    FloatD y = exp(x) * sin(x) + x / (x + 1.f);
    forward(x);
    std::vector<float> r; append(r, y); append_raw(r, gradient(y)); return r;
}
SCENARIO(s_psum_reverse) { TYPES
    FloatD x = linspace<FloatD>(0.5f, 1.5f, 12);
    set_requires_gradient(x);
    FloatD y = psum(x * x) * reverse(x);
    backward(hsum(y));
    std::vector<float> r; append(r, y); append_raw(r, gradient(x)); return r;
}
SCENARIO(s_descent) { TYPES            /* tests/autodiff.cpp:550-562: a few steps of gradient descent */
    FloatD x = zero<FloatD>(10);
    for (int i = 0; i < 8; ++i) {
        set_requires_gradient(x);
        FloatD loss = hsum(sqr(x - linspace<FloatD>(0.f, 1.f, 10)));
        backward(loss);
        x = detach(x) - gradient(x) * 0.25f;
    }
    std::vector<float> r; append(r, x); return r;
}

/* distance in units in the last place between two floats (0 for equal values incl. +-0; NaN vs NaN = 0) */
static double ulp_dist(float a, float b) {
    if (a == b || (std::isnan(a) && std::isnan(b))) return 0;
    if (std::isnan(a) || std::isnan(b) || std::isinf(a) || std::isinf(b)) return 1e30;
    int32_t ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    auto key = [](int32_t v) -> int64_t { return v < 0 ? (int64_t) 0x80000000ll - (int64_t) (uint32_t) v : (int64_t) v; };   /* monotone in the float order */
    return std::fabs((double) (key(ia) - key(ib)));
}

/* Gates (north star: 1 ulp for fp32 arithmetic, 4 ulp for transcendentals, reductions by reassociation bound):
     max_ulp >= 0 : every value within that many ulp of the reference CPU tape (values of magnitude below 1e-30 are
                    compared absolutely: a gradient that is exactly 0 on one side and 1e-40 on the other is not 1e9 ulp)
     max_ulp <  0 : the chain contains rcp / rsqrt (the CPU path is rcpps / rsqrtps + one Newton step, itself up to 3 ulp from
                    the correctly rounded value and CPU-vendor dependent, array_avx.h:324-395), composite functions built on
                    them (tan, sinh, tanh, atan2, asin: array_math.h), or float reductions folded in a different order
                    (hsum, psum): only the relative bound `tol` applies -- the observed max ulp is still printed. */
struct Case { const char *name; std::function<std::vector<float>()> cpu, gpu; float tol; double max_ulp; };
#define CASE(fn, tol, ulp) Case { #fn, fn<DiffArray<FloatX>>, fn<DiffArray<FloatC>>, tol, ulp }

int main() {
    if (ek_device_count() == 0) { fprintf(stderr, "no CUDA device\n"); return 2; }
    std::vector<Case> cases = {
        /* Round 2: every ulp gate is still "report only" (-1).  The scenarios build their inputs with linspace(), which the
           reference evaluates differently on its CPU and CUDA paths (dynamic.h:923-938 vs cuda.h:655-663), so an input may
           already differ by an ulp; the gates can only be set from the numbers this program prints on hardware, and the
           repository had no GPU access when the ulp report was added.  The Python tests (tests/test_gpu_eval.py,
           test_gpu_tape.py) feed identical inputs to both sides and ARE bit-exact gates. */
        CASE(s_arith, 2e-5f, -1), CASE(s_trig, 2e-5f, -1), CASE(s_select, 1e-6f, -1), CASE(s_broadcast, 2e-5f, -1), CASE(s_hprod, 2e-5f, -1),
        CASE(s_scatter_add, 1e-6f, -1), CASE(s_scatter, 1e-6f, -1), CASE(s_gather, 1e-6f, -1), CASE(s_gather_fwd, 1e-6f, -1),
        CASE(s_scatter_fwd, 1e-6f, -1), CASE(s_forward, 2e-5f, -1), CASE(s_psum_reverse, 2e-5f, -1), CASE(s_descent, 1e-5f, -1) };
    int failures = 0;
    for (auto &c : cases) {
        std::vector<float> a, b;
        try { a = c.cpu(); b = c.gpu(); }
        catch (const std::exception &e) { printf("%-16s EXCEPTION %s\n", c.name, e.what()); ++failures; continue; }
        bool ok = a.size() == b.size() && !a.empty();
        double maxd = 0, maxu = 0;
        for (size_t i = 0; ok && i < a.size(); ++i) {
            double d = std::fabs((double) a[i] - b[i]) / std::max(1.0, std::fabs((double) a[i]));
            if (!(d <= c.tol)) ok = false;
            maxd = std::max(maxd, d);
            double u = (std::fabs(a[i]) < 1e-30f && std::fabs(b[i]) < 1e-30f) ? 0.0 : ulp_dist(a[i], b[i]);
            maxu = std::max(maxu, u);
        }
        if (ok && c.max_ulp >= 0 && maxu > c.max_ulp) ok = false;
        printf("%-16s %s  n=%zu/%zu  max rel diff %.3g  max ulp %.0f (gate: %s)\n", c.name, ok ? "ok  " : "FAIL", a.size(), b.size(), maxd, maxu,
               c.max_ulp >= 0 ? "ulp" : "relative bound only");
        if (!ok) for (size_t i = 0; i < std::min(a.size(), b.size()); ++i) printf("    [%zu] cpu % .9g  gpu % .9g\n", i, a[i], b[i]);
        if (!ok) ++failures;
    }
    /* published expected vectors of the reference's own tests (tests/autodiff.cpp:414-430,620-635) */
    {
        auto r = s_scatter_add<DiffArray<FloatC>>();
        const float ref[] = { 0.f, .25f, .5f, 1.75f, 2.3333f, 1.6667f, 2.f, 0.f, 0.f, 0.f,  0.f, .5f, 1.f, 3.5f, 4.6667f,  3.5f, 4.6667f, 3.3333f, 4.f };
        for (size_t i = 0; i < 19; ++i) if (std::fabs(r[i] - ref[i]) > 1e-4f + 1e-4f * std::fabs(ref[i])) { printf("scatter_add known answer %zu: %f vs %f\n", i, r[i], ref[i]); ++failures; }
        auto g = s_gather_fwd<DiffArray<FloatC>>();
        const float refg[] = { -1.55556f, -1.11111f, -0.666667f };
        for (size_t i = 0; i < 3; ++i) if (std::fabs(g[i] - refg[i]) > 1e-4f) { printf("gather_fwd known answer %zu: %f vs %f\n", i, g[i], refg[i]); ++failures; }
        auto sf = s_scatter_fwd<DiffArray<FloatC>>();
        const float refs[] = { 0.f, 0.f, -2.f, -1.f, 0.f, 1.f, 2.f, 0.f, 0.f, 0.f };
        for (size_t i = 0; i < 10; ++i) if (std::fabs(sf[i] - refs[i]) > 1e-4f) { printf("scatter_fwd known answer %zu: %f vs %f\n", i, sf[i], refs[i]); ++failures; }
    }
    cuda_sync();
    printf(failures ? "autodiff_check: FAILED (%d)\n" : "autodiff_check: all checks passed\n", failures);
    return failures ? 1 : 0;
}
