/* Integer / indexing / Morton parity (BASELINE north star: bit-exact against the reference's CPU path).  The same
   templates run on DynamicArray<Packet<T, 8>> (reference headers, AVX2) and on CUDAArray<T> (this backend through
   include/enoki/cuda.h): arithmetic, shifts, mulhi, division / remainder, division by precomputed constants
   (array_idiv.h:150-251), popcnt / lzcnt / tzcnt, Morton encode / decode in 2 and 3 dimensions (morton.h:27-155),
   int <-> float conversions.  Every result array is compared bit for bit.  Written from scratch. */
#include <enoki/cuda.h>
#include <enoki/dynamic.h>
#include <enoki/morton.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include <string>

using namespace enoki;

static int g_fail = 0;

template <typename A> std::vector<scalar_t<A>> to_host(const A &a) {
    using S = scalar_t<A>;
    std::vector<S> v(a.size());
    if constexpr (is_cuda_array_v<A>) { a.eval(); cuda_memcpy_from_device(v.data(), a.data(), v.size() * sizeof(S)); }
    else memcpy(v.data(), a.data(), v.size() * sizeof(S));
    return v;
}
template <typename A, typename B> void cmp(const char *what, const char *type, const A &a, const B &b) {
    auto x = to_host(a); auto y = to_host(b);
    size_t bad = x.size() != y.size();
    for (size_t i = 0; i < std::min(x.size(), y.size()); ++i) if (memcmp(&x[i], &y[i], sizeof(x[i])) != 0) ++bad;
    printf("%-16s %-9s %s  n=%zu mismatches=%zu\n", what, type, bad ? "FAIL" : "ok  ", x.size(), bad);
    if (bad) g_fail = 1;
}

/* results of one template instantiated over the two backends, keyed by name */
template <typename Int, typename Float32A> struct Results { std::vector<std::pair<std::string, Int>> ints; std::vector<std::pair<std::string, Float32A>> floats; };

template <typename Int> auto run_int(size_t n) {
    using S = scalar_t<Int>;
    using U = std::make_unsigned_t<S>;
    using UIntA = replace_scalar_t<Int, U>;
    std::vector<std::pair<std::string, Int>> out;
    /* pseudo-random operands from index arithmetic (identical on both backends, wraps like the scalar type) */
    Int i = arange<Int>(n);
    Int x = i * S(2654435761u) + S(12345), y = (i ^ sr<3>(i)) * S(40503) + S(7);
    Int ynz = y | S(1);
    out.emplace_back("add", x + y); out.emplace_back("sub", x - y); out.emplace_back("mul", x * y);
    out.emplace_back("mulhi", mulhi(x, y));
    out.emplace_back("div", x / ynz); out.emplace_back("mod", x % ynz);
    out.emplace_back("and/or/xor", (x & y) ^ (x | y));
    out.emplace_back("not/neg", ~x - (-y));
    out.emplace_back("sl<5>", sl<5>(x)); out.emplace_back("sr<7>", sr<7>(x));
    Int amt = i & S(sizeof(S) * 8 - 1);
    out.emplace_back("sl var", x << amt); out.emplace_back("sr var", x >> amt);
    out.emplace_back("min/max", min(x, y) + max(x, y));
    out.emplace_back("abs", abs(x));
    out.emplace_back("popcnt", popcnt(x)); out.emplace_back("lzcnt", lzcnt(y)); out.emplace_back("tzcnt", tzcnt(x | S(64)));
    out.emplace_back("div const 7", x / divisor<S>(S(7)));
    out.emplace_back("div const 641", x / divisor<S>(S(641)));
    out.emplace_back("div const 16", x / divisor<S>(S(16)));
    out.emplace_back("select", select(x < y, x, y));
    if constexpr (std::is_unsigned_v<S>) {
        /* Morton codes: 2-D and 3-D encode, decode round trip */
        constexpr int B2 = sizeof(S) * 4, B3 = sizeof(S) * 8 / 3;
        UIntA a = UIntA(x) & U((U(1) << B2) - 1), b = UIntA(y) & U((U(1) << B2) - 1);
        UIntA m2 = morton_encode(Array<UIntA, 2>(a, b));
        out.emplace_back("morton2 enc", Int(m2));
        auto d2 = morton_decode<Array<UIntA, 2>>(m2);
        out.emplace_back("morton2 dec", Int(d2.x() + d2.y() * U(3)));
        UIntA a3 = UIntA(x) & U((U(1) << B3) - 1), b3 = UIntA(y) & U((U(1) << B3) - 1), c3 = UIntA(x ^ y) & U((U(1) << B3) - 1);
        UIntA m3 = morton_encode(Array<UIntA, 3>(a3, b3, c3));
        out.emplace_back("morton3 enc", Int(m3));
        auto d3 = morton_decode<Array<UIntA, 3>>(m3);
        out.emplace_back("morton3 dec", Int(d3.x() + d3.y() * U(3) + d3.z() * U(5)));
    }
    return out;
}

template <typename CpuInt, typename GpuInt> void check_int(const char *type, size_t n) {
    auto a = run_int<CpuInt>(n);
    auto b = run_int<GpuInt>(n);
    for (size_t k = 0; k < a.size(); ++k) cmp(a[k].first.c_str(), type, a[k].second, b[k].second);
}

template <typename FloatA, typename IntA, typename UIntA> auto run_cvt(size_t n) {
    std::vector<std::pair<std::string, IntA>> out;
    /* inputs from integer index arithmetic + one fmadd: identical on both backends (linspace() is not: the reference's
       CPU DynamicArray accumulates `value += step` per packet, its CUDA backend evaluates fmadd(index, step, min),
       dynamic.h:923-938 vs cuda.h:655-663) */
    IntA k = arange<IntA>(n) * 7919 - 1000000;
    FloatA f = fmadd(FloatA(k), FloatA(0.13371337f), FloatA(-0.75f));
    out.emplace_back("f32->i32", IntA(f));
    out.emplace_back("floor2int", floor2int<IntA>(f)); out.emplace_back("ceil2int", ceil2int<IntA>(f));
    out.emplace_back("f32->u32", IntA(UIntA(abs(f))));
    out.emplace_back("i32->f32 bits", reinterpret_array<IntA>(FloatA(k)));
    out.emplace_back("u32->f32 bits", reinterpret_array<IntA>(FloatA(UIntA(k) & 0x7fffffffu)));
    return out;
}

/* uint32 -> float above 2^31: the reference's AVX2 packets round twice (float(x & 0x7fffffff) + 2^31, array_avx.h:56-66),
   its scalar and AVX-512 paths round once.  This backend rounds once; the yardstick is the scalar conversion. */
static void check_u32_to_f32(size_t n) {
    using UIntC = CUDAArray<uint32_t>; using FloatC = CUDAArray<float>;
    UIntC u = arange<UIntC>(n) * 2654435761u + 0x80000000u;
    auto hu = to_host(u); auto hf = to_host(FloatC(u));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) { float want = (float) hu[i]; if (memcmp(&want, &hf[i], 4) != 0) ++bad; }
    printf("%-16s %-9s %s  n=%zu mismatches=%zu (vs scalar conversion)\n", "u32->f32 >=2^31", "convert", bad ? "FAIL" : "ok  ", n, bad);
    if (bad) g_fail = 1;
}

int main(int argc, char **argv) {
    size_t n = argc > 1 ? (size_t) atoll(argv[1]) : 100003;
    if (ek_device_count() == 0) { fprintf(stderr, "no CUDA device\n"); return 2; }
    check_int<DynamicArray<Packet<uint32_t, 8>>, CUDAArray<uint32_t>>("uint32", n);
    check_int<DynamicArray<Packet<int32_t, 8>>, CUDAArray<int32_t>>("int32", n);
    check_int<DynamicArray<Packet<uint64_t, 4>>, CUDAArray<uint64_t>>("uint64", n);
    check_int<DynamicArray<Packet<int64_t, 4>>, CUDAArray<int64_t>>("int64", n);
    {
        auto a = run_cvt<DynamicArray<Packet<float, 8>>, DynamicArray<Packet<int32_t, 8>>, DynamicArray<Packet<uint32_t, 8>>>(n);
        auto b = run_cvt<CUDAArray<float>, CUDAArray<int32_t>, CUDAArray<uint32_t>>(n);
        for (size_t k = 0; k < a.size(); ++k) cmp(a[k].first.c_str(), "convert", a[k].second, b[k].second);
    }
    check_u32_to_f32(n);
    printf(g_fail ? "int_check: FAILED\n" : "int_check: all checks passed\n");
    return g_fail;
}
