/* CPU-side check of the header shim (no GPU needed): records expressions through include/enoki/cuda.h exactly as user
   code would (reference router + array_math.h on top of enoki::CUDAArray<T>) over borrowed fake device pointers, and
   prints the sweep programs the planner would launch (ek_debug_plan).  tests/test_cpu_abi.py compares them with the
   programs recorded through the Python mirror.  Written from scratch. */
#include <enoki/cuda.h>
#include <enoki/array.h>
#include <cstdio>
#include <cstdlib>

using namespace enoki;
extern "C" char *ek_debug_plan(void);

static void dump(const char *tag) {
    char *p = ek_debug_plan();
    printf("== %s\n%s", tag, p ? p : "(error)\n");
    free(p);
}

int main() {
    using FloatC = CUDAArray<float>;
    using UIntC = CUDAArray<uint32_t>;
    const size_t n = 1 << 20;
    FloatC x0 = FloatC::map((void *) 0x7f0001000000ull, n), x1 = FloatC::map((void *) 0x7f0002000000ull, n),
           x2 = FloatC::map((void *) 0x7f0003000000ull, n), x3 = FloatC::map((void *) 0x7f0004000000ull, n);
    {   /* C2 */
        FloatC t = fmadd(x0, x1, x2);
        FloatC u = exp(-(t * t));
        FloatC v = sin(fmadd(x3, u, x0));
        FloatC out = fmadd(v, x1, sqrt(abs(t)));
        t = FloatC(); u = FloatC(); v = FloatC();
        dump("c2");
    }
    {   /* a reduction that feeds a later phase, an integer / conversion mix */
        FloatC y = x0 / hsum(x0 * x0);
        UIntC i = UIntC(abs(x1) * 1000.f) & 1023u;
        FloatC z = select(i < 512u, y, FloatC(i));
        y = FloatC(); i = UIntC();
        dump("phases");
    }
    return 0;
}
