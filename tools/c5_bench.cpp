/* c5_bench -- BASELINE.json configs[4] as a measured workload: differentiable ray-sphere render,
   res x res orthographic rays, forward + backward through enoki::DiffArray<enoki::CUDAArray<float>> exactly as a
   Mitsuba-style caller would write it (reference headers + this repo's <enoki/cuda.h> / <enoki/autodiff_b200.h>).
   Scene after the reference's tests/sphere.cpp:58-88; same template as tests/cpp/sphere_check.cpp (which holds the
   parity check against the reference CPU tape -- nothing of oracle/ is linked here).
   Prints ONE JSON line:  c5_bench [resolution=4096] [iterations=5]
   A "step" = trace the render, cuda_eval() the image, loss = hsum((img - target)^2) / n, backward(loss), read the six
   scalar gradients.  Measured with graph simplification on (the default, autodiff.cpp:990-1074) and off. */
#include <enoki/autodiff.h>
#include <enoki/cuda.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace enoki;
using FloatC = CUDAArray<float>;
using FloatD = DiffArray<FloatC>;
using UIntC = CUDAArray<uint32_t>;
using Vec3D = Array<FloatD, 3>;

static FloatD render(const FloatD &px, const FloatD &py, const Vec3D &delta, const Vec3D &light) {
    Vec3D o = Vec3D(px, py, FloatD(-1.f)) + delta;
    Vec3D d(FloatD(0.f), FloatD(0.f), FloatD(1.f));
    FloatD a = dot(d, d), b = 2.f * dot(o, d), c = dot(o, o) - 1.f;
    FloatD disc = b * b - 4.f * a * c;
    FloatD t = (-b - sqrt(max(disc, 0.f))) / (2.f * a);
    Vec3D n = o + d * t;
    FloatD shade = 0.2f + max(dot(n, light), 0.f) * 0.9f;
    return select(disc >= 0.f, shade, FloatD(0.f));
}

struct StepResult { float loss; float grads[6]; };

static StepResult step(size_t res, const FloatC &target) {
    size_t n = res * res;
    UIntC idx = arange<UIntC>(n);
    FloatC fx = FloatC(idx % (uint32_t) res), fy = FloatC(idx / (uint32_t) res);
    float st = 2.4f / float(res - 1);
    FloatD px = FloatD(fmadd(fx, st, -1.2f)), py = FloatD(fmadd(fy, st, -1.2f));
    Vec3D delta(FloatD(0.05f), FloatD(-0.03f), FloatD(0.02f));
    float il = 1.f / std::sqrt(6.f);
    Vec3D light(FloatD(-il), FloatD(-il), FloatD(-2.f * il));
    for (int k = 0; k < 3; ++k) { set_requires_gradient(delta[k]); set_requires_gradient(light[k]); }
    FloatD img = render(px, py, delta, light);
    FloatD diff = img - FloatD(target);
    FloatD loss = hsum(diff * diff) / float(n);
    StepResult r;
    r.loss = detach(loss).coeff(0);
    backward(loss);
    for (int k = 0; k < 3; ++k) { r.grads[k] = gradient(delta[k]).coeff(0); r.grads[3 + k] = gradient(light[k]).coeff(0); }
    return r;
}

int main(int argc, char **argv) {
    size_t res = argc > 1 ? (size_t) atoi(argv[1]) : 4096;
    int iters = argc > 2 ? atoi(argv[2]) : 5;
    if (ek_device_count() == 0) { fprintf(stderr, "c5_bench: no CUDA device\n"); return 2; }
    size_t n = res * res;
    /* target = the render at delta = 0: computed once, resident in HBM */
    FloatC target;
    {
        UIntC idx = arange<UIntC>(n);
        FloatC fx = FloatC(idx % (uint32_t) res), fy = FloatC(idx / (uint32_t) res);
        float st = 2.4f / float(res - 1);
        float il = 1.f / std::sqrt(6.f);
        FloatD img = render(FloatD(fmadd(fx, st, -1.2f)), FloatD(fmadd(fy, st, -1.2f)), Vec3D(FloatD(0.f), FloatD(0.f), FloatD(0.f)),
                            Vec3D(FloatD(-il), FloatD(-il), FloatD(-2.f * il)));
        target = detach(img);
        target.eval();
    }
    cuda_sync();
    printf("{\"workload\": \"C5: differentiable ray-sphere render, %zux%zu rays, forward + backward, 6 scalar gradients\", \"rays\": %zu", res, res, n);
    for (int simplify = 1; simplify >= 0; --simplify) {
        ek_tape_set_graph_simplification(EK_FLOAT32, simplify);
        StepResult r = step(res, target); r = step(res, target);          /* warm-up */
        cuda_sync();
        ek_stats_reset();
        auto t0 = std::chrono::high_resolution_clock::now();
        for (int k = 0; k < iters; ++k) r = step(res, target);
        cuda_sync();
        auto t1 = std::chrono::high_resolution_clock::now();
        ek_stats st; ek_stats_get(&st);
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
        /* device time of the launches of one step (per-launch CUDA events; serialises host and device) */
        ek_set_timing(1); ek_stats_reset();
        r = step(res, target);
        cuda_sync();
        ek_stats stt; ek_stats_get(&stt);
        ek_set_timing(0);
        double sweep_bytes = (double) (st.bytes_in + st.bytes_out) / iters, adj_bytes = 10.0 * (double) st.edge_adjoints / iters;
        printf(", \"%s\": {\"ms_per_step\": %.4f, \"kernels_ms\": %.4f, \"launches_per_step\": %.1f, \"sweep_launches_per_step\": %.1f, "
               "\"fast_kernel_launches_per_step\": %.1f, \"adjoint_launches_per_step\": %.1f, \"sweep_bytes_per_step\": %.0f, "
               "\"edge_adjoints_per_step\": %.0f, \"adjoint_bytes_per_step\": %.0f, \"loss\": %.9g, "
               "\"grads\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g]}",
               simplify ? "simplify_on" : "simplify_off", ms, (double) stt.total_kernel_ms, (double) st.launches / iters,
               (double) st.sweep_launches / iters, (double) st.fast_launches / iters, (double) st.adjoint_launches / iters, sweep_bytes,
               (double) st.edge_adjoints / iters, adj_bytes, r.loss, r.grads[0], r.grads[1], r.grads[2], r.grads[3], r.grads[4], r.grads[5]);
    }
    ek_tape_set_graph_simplification(EK_FLOAT32, 1);
    printf("}\n");
    return 0;
}
