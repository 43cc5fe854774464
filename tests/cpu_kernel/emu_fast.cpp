/*
 * emu_fast.cpp -- runs enoki_b200/csrc/ek_sweep_fast.cu ON THE CPU: the kernel source is compiled as host code
 * (cuda_shim.h) and every CUDA thread of a CTA becomes a POSIX thread; CTAs run one after the other.
 * TEST INFRASTRUCTURE (tests/test_cpu_fast_kernel.py): see cuda_shim.h for what this can and cannot show.
 *
 * extern "C" emu_run_fast(): fills EkSweepArgs exactly like the launch code of ek_eval.cpp does (from the fields of
 * the `fast` section of ek_debug_program()) and "launches" ek_fast_kernel<T>.
 */
#include "cuda_shim.h"
#include <pthread.h>
#include <vector>

namespace emu {
thread_local EmuDim tid_;
EmuDim bid_, bdim_, gdim_;
uint8_t *smem_ = nullptr;
static pthread_barrier_t cta_bar_;
static std::atomic<int> failed_{ 0 };
static char fail_msg_[256];

void cta_barrier() { pthread_barrier_wait(&cta_bar_); }

[[noreturn]] void trap(const char *why) {
    int expect = 0;
    if (failed_.compare_exchange_strong(expect, 1)) snprintf(fail_msg_, sizeof(fail_msg_), "block %u thread %u: %s", bid_.x, tid_.x, why);
    fprintf(stderr, "emu_fast: TRAP (%s)\n", fail_msg_);
    abort();                                   /* (other threads may be parked in barriers: there is no clean unwind) */
}

/* per-warp rendezvous for an arbitrary lane mask: arrive, read everybody's value, leave.  Rendezvous with DIFFERENT
   masks may be in flight in one warp at the same time (independent thread scheduling: the lanes that left
   warp_agg_atomic_add early already wait in the next full-mask ballot while the others still match addresses), so
   every (warp, mask) pair has its own slot. */
struct WarpSlot {
    std::atomic<unsigned> key{ 0 };
    uint64_t val[32];
    std::atomic<unsigned> arrived{ 0 }, left{ 0 };
    std::atomic<unsigned> gen{ 0 };
};
static WarpSlot warps_[32][64];

static WarpSlot &slot_for(unsigned mask) {
    WarpSlot *row = warps_[tid_.x >> 5];
    for (unsigned h = (mask * 2654435761u) >> 26, probes = 0; probes < 64; ++probes, h = (h + 1) & 63u) {
        unsigned k = row[h].key.load();
        if (k == mask) return row[h];
        if (k == 0) { unsigned expect = 0; if (row[h].key.compare_exchange_strong(expect, mask) || expect == mask) return row[h]; }
    }
    trap("too many distinct lane masks in one warp");
}
const uint64_t *warp_exchange(unsigned mask, uint64_t mine) {
    const unsigned lane = tid_.x & 31u, want = (unsigned) __builtin_popcount(mask);
    if (!((mask >> lane) & 1u)) trap("lane calls a warp primitive with a mask that does not name it");
    WarpSlot &w = slot_for(mask);
    const unsigned g = w.gen.load();
    w.val[lane] = mine;
    if (w.arrived.fetch_add(1) + 1 == want) { w.arrived.store(0); w.gen.store(g + 1); }
    else { unsigned long spins = 0; while (w.gen.load() == g) { if (++spins > 400000000ul) trap("warp rendezvous never completed"); sched_yield(); } }
    return w.val;
}
void warp_release(unsigned mask) {
    WarpSlot &w = slot_for(mask);
    const unsigned want = (unsigned) __builtin_popcount(mask);
    const unsigned g = w.gen.load();
    if (w.left.fetch_add(1) + 1 == want) { w.left.store(0); w.gen.store(g + 1); }
    else { unsigned long spins = 0; while (w.gen.load() == g) { if (++spins > 400000000ul) trap("warp rendezvous never completed"); sched_yield(); } }
}
} // namespace emu

#include "../../enoki_b200/csrc/ek_sweep_fast.cu"

namespace {
struct ThreadArg { const EkSweepArgs *args; unsigned tid; unsigned T; };
void *thread_main(void *p) {
    const ThreadArg *a = (const ThreadArg *) p;
    emu::tid_.x = a->tid;
    if (a->T == 256) ek_fast_kernel<256>(*a->args); else ek_fast_kernel<128>(*a->args);
    return nullptr;
}
}

extern "C" int emu_run_fast(const uint32_t *prog_words, uint32_t n_init, uint32_t n_body, uint32_t n_fini,
                            const uint32_t *lits, uint32_t n_lit, const uint32_t *argw, uint32_t n_argw,
                            const uint64_t *scalar_ptr, const uint8_t *scalar_type, uint32_t n_scalar,
                            const uint64_t *staged_ptr, const uint16_t *staged_unit, const uint8_t *staged_esize, uint32_t n_staged,
                            uint32_t n, uint32_t T, uint32_t n_tmp, uint32_t n_in_units,
                            uint32_t off_bar, uint32_t off_extra, uint32_t off_slots, uint32_t smem_bytes,
                            uint32_t grid, uint32_t n_red) {
    if (T != 128 && T != 256) return -1;
    if (n_init + n_body + n_fini > EK_INLINE_PROG || n_lit > EK_MAX_LIT_INLINE || n_argw > EK_MAX_ARGW ||
        n_scalar > EK_MAX_SCALAR || n_staged > EK_MAX_STAGED) return -2;
    static EkSweepArgs args;                     /* (large: keep it off the stack) */
    memset(&args, 0, sizeof(args));
    memcpy(args.prog_inline, prog_words, (size_t) (n_init + n_body + n_fini) * 16);
    args.n_init = n_init; args.n_body = n_body; args.n_fini = n_fini;
    args.lit = nullptr; memcpy(args.lit_inline, lits, (size_t) n_lit * 4); args.n_lit = n_lit;
    memcpy(args.argw, argw, (size_t) n_argw * 4); args.n_argw = n_argw;
    args.n_scalar = n_scalar;
    for (uint32_t k = 0; k < n_scalar; ++k) { args.scalar_ptr[k] = (const void *) (uintptr_t) scalar_ptr[k]; args.scalar_type[k] = scalar_type[k]; }
    args.n_staged = n_staged; args.tma_ok = 1;
    for (uint32_t k = 0; k < n_staged; ++k) {
        args.staged_ptr[k] = (const void *) (uintptr_t) staged_ptr[k]; args.staged_unit[k] = staged_unit[k]; args.staged_esize[k] = staged_esize[k];
        if (staged_ptr[k] & 15u) args.tma_ok = 0;
    }
    args.n = n;
    const uint32_t tile = T * 16u;
    args.n_tiles = (n + tile - 1) / tile;
    args.n_tmp = n_tmp; args.n_in_units = n_in_units; args.n_stages = 1;
    args.smem_bar_off = off_bar; args.smem_extra_off = off_extra; args.smem_slots_off = off_slots;
    args.n_red = n_red;
    std::vector<uint64_t> partials((size_t) std::max(n_red, 1u) * grid, 0xdeadbeefdeadbeefull);
    std::vector<uint32_t> counters(std::max(n_red, 1u), 0u);
    args.red_partials = partials.data(); args.red_counters = counters.data();

    emu::gdim_.x = grid; emu::bdim_.x = T;
    void *raw = nullptr;
    if (posix_memalign(&raw, 1024, smem_bytes + 1024) != 0) return -3;
    emu::smem_ = (uint8_t *) raw;
    pthread_barrier_init(&emu::cta_bar_, nullptr, T);
    std::vector<pthread_t> th(T);
    std::vector<ThreadArg> ta(T);
    pthread_attr_t attr; pthread_attr_init(&attr); pthread_attr_setstacksize(&attr, 256 * 1024);
    for (uint32_t b = 0; b < grid; ++b) {
        emu::bid_.x = b;
        memset(emu::smem_, 0xCD, smem_bytes + 1024);       /* shared memory starts out as garbage on the device, too */
        for (uint32_t t = 0; t < T; ++t) { ta[t] = { &args, t, T }; pthread_create(&th[t], &attr, thread_main, &ta[t]); }
        for (uint32_t t = 0; t < T; ++t) pthread_join(th[t], nullptr);
    }
    pthread_attr_destroy(&attr);
    pthread_barrier_destroy(&emu::cta_bar_);
    free(raw);
    emu::smem_ = nullptr;
    for (uint32_t k = 0; k < n_red; ++k) if (counters[k] != 0) return -4;     /* tickets must be back at zero */
    return 0;
}
