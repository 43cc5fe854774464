/*
 * ek_dist.cpp -- multi-GPU entry points of the C ABI: one rank (process) per GPU, element-range sharding
 * of every wide array, and ONE all-reduce of the size-1 results (loss, gradients of size-1 leaves) per step -- the
 * local hsum happens where the reference has it (autodiff.cpp:867-871), the cross-GPU sum here.
 *
 * The reference has no multi-GPU support at all; these calls exist so that a C++ caller of the boundary (not only the
 * Python mirror through torch.distributed) can shard: rank 0 calls ek_dist_unique_id(), hands the 128 bytes to the other
 * ranks by whatever means it has (MPI, a file, a socket), every rank calls ek_dist_init(), and
 * ek_allreduce_scalars() / ek_allreduce() enqueue ncclAllReduce on the backend's stream (no host synchronisation).
 *
 * NCCL is loaded with dlopen() on first use (EK_NCCL_LIB, else libnccl.so.2): libenoki_b200.so keeps no link-time
 * dependency on it, single-GPU users never touch it.
 */
#include "ek_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {
struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
} g_nccl;

bool load_nccl() {
    if (g_nccl.lib) return true;
    const char *names[3] = { getenv("EK_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    std::string tried;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.lib) break;
        tried += std::string(nm) + " ";
    }
    if (!g_nccl.lib) { ek_set_error("ek_dist: could not load NCCL (tried " + tried + "; set EK_NCCL_LIB)"); return false; }
#define SYM(field, name) *(void **) (&g_nccl.field) = dlsym(g_nccl.lib, name); if (!g_nccl.field) { ek_set_error(std::string("ek_dist: NCCL symbol missing: ") + name); dlclose(g_nccl.lib); g_nccl.lib = nullptr; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(CommDestroy, "ncclCommDestroy") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}
bool nccl_ok(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return true;
    ek_set_error(std::string("ek_dist: ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
    return false;
}
ncclDataType_t nccl_type(ek_type t, bool &ok) {
    ok = true;
    switch (t) {
        case EK_FLOAT32: return ncclFloat32; case EK_FLOAT64: return ncclFloat64;
        case EK_INT32: return ncclInt32; case EK_UINT32: return ncclUint32;
        case EK_INT64: return ncclInt64; case EK_UINT64: return ncclUint64;
        default: ok = false; return ncclFloat32;
    }
}
} // namespace

extern "C" {

int ek_dist_unique_id(void *out128) {
    static_assert(sizeof(ncclUniqueId) == EK_DIST_ID_BYTES, "EK_DIST_ID_BYTES must equal sizeof(ncclUniqueId)");
    if (!load_nccl()) return -1;
    ncclUniqueId id;
    if (!nccl_ok(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
    memcpy(out128, &id, sizeof(id));
    return 0;
}

int ek_dist_init(int rank, int world, const void *id128) {
    if (world < 1 || rank < 0 || rank >= world) { ek_set_error("ek_dist_init(): bad rank / world size"); return -1; }
    if (g_nccl.comm) { ek_set_error("ek_dist_init(): already initialised"); return -1; }
    if (ek_init() != 0) return -1;
    g_nccl.rank = rank; g_nccl.world = world;
    if (world == 1) return 0;                      /* nothing to communicate with */
    if (!load_nccl()) return -1;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ek_cuda_check(cudaSetDevice(ek_ctx().device));
    return nccl_ok(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank), "ncclCommInitRank") ? 0 : -1;
}

int ek_dist_rank(void) { return g_nccl.rank; }
int ek_dist_world(void) { return g_nccl.world; }

/* in-place sum over all ranks of `count` values at device pointer `data`, on the backend's stream */
int ek_allreduce(ek_type type, void *data, size_t count) {
    if (g_nccl.world == 1) return 0;
    if (!g_nccl.comm) { ek_set_error("ek_allreduce(): call ek_dist_init() first"); return -1; }
    bool ok; ncclDataType_t t = nccl_type(type, ok);
    if (!ok) { ek_set_error("ek_allreduce(): unsupported type"); return -1; }
    return nccl_ok(g_nccl.AllReduce(data, data, count, t, ncclSum, g_nccl.comm, ek_ctx().stream), "ncclAllReduce") ? 0 : -1;
}

/* the step's cross-GPU exchange: every handle is evaluated if necessary, then summed in place over all
   ranks -- one NCCL group, enqueued on the backend's stream.  Size-1 variables are the intended use (loss, gradients of
   scalar leaves); wider ones work too (e.g. a replicated histogram after a sharded scatter_add). */
int ek_allreduce_scalars(const uint32_t *handles, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (ek_var_ptr(handles[i]) == nullptr && ek_eval_var(handles[i]) != 0) return -1;
        if (ek_var_ptr(handles[i]) == nullptr) { ek_set_error("ek_allreduce_scalars(): variable has no storage"); return -1; }
    }
    if (g_nccl.world == 1) return 0;
    if (!g_nccl.comm) { ek_set_error("ek_allreduce_scalars(): call ek_dist_init() first"); return -1; }
    if (!nccl_ok(g_nccl.GroupStart(), "ncclGroupStart")) return -1;
    bool good = true;
    for (size_t i = 0; i < n && good; ++i) {
        bool ok; ncclDataType_t t = nccl_type(ek_var_type(handles[i]), ok);
        if (!ok) { ek_set_error("ek_allreduce_scalars(): unsupported type"); good = false; break; }
        void *p = ek_var_ptr(handles[i]);
        good = nccl_ok(g_nccl.AllReduce(p, p, ek_var_size(handles[i]), t, ncclSum, g_nccl.comm, ek_ctx().stream), "ncclAllReduce");
    }
    bool ended = nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd");
    return good && ended ? 0 : -1;
}

void ek_dist_shutdown(void) {
    if (g_nccl.comm && g_nccl.CommDestroy) { cudaStreamSynchronize(ek_ctx().stream); g_nccl.CommDestroy(g_nccl.comm); }
    g_nccl.comm = nullptr; g_nccl.rank = 0; g_nccl.world = 1;
}

} /* extern "C" */
