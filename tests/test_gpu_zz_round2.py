"""Round-2 GPU tests.  Everything in this file was written AFTER the repository had lost its GPU access (round 2), so none
of it has run on hardware yet; the file name sorts last so that `pytest -x` reaches it only after the suite that passed
in round 1.

GPU tests of the user-visible runtime calls of the boundary: cuda_whos, cuda_make_managed,
cuda_fetch_element on unevaluated variables, pre-eval callbacks, cuda_malloc_trim / cuda_mem_get_info, the kernel
qualification verdict, and the cuda_partition device composition (opt-in path).
Reference behaviour: src/cuda/jit.cu:455-485 (managed), :1520-1538 (fetch_element), :1421-1422 (callbacks),
:1564-1634 (whos), :1715-1723 (trim + retry)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


F32 = 10
BIN = os.path.join(ROOT, "tests", "cpp", "shim_smoke")


def test_whos_lists_variables(gpu):
    ek = gpu
    a = ek.Float32.copy(np.arange(1000, dtype=np.float32)).set_label("whos_probe_input")
    b = (a * 2.0).set_label("whos_probe_pending")
    txt = ek.cuda_whos()
    assert "whos_probe_input" in txt and "whos_probe_pending" in txt
    line_in = next(l for l in txt.splitlines() if "whos_probe_input" in l)
    line_pe = next(l for l in txt.splitlines() if "whos_probe_pending" in l)
    assert "[x]" in line_in and "[ ]" in line_pe             # ready / scheduled (jit.cu:1597-1600)
    assert "4000" in line_in                                 # bytes
    b.eval()
    line_pe = next(l for l in ek.cuda_whos().splitlines() if "whos_probe_pending" in l)
    assert "[x]" in line_pe
    assert "Memory usage" in txt


def test_fetch_element_evaluates_on_demand(gpu):
    """cuda_fetch_element (jit.cu:1520-1538): reading an element of an unevaluated variable evaluates it first."""
    ek = gpu
    x = np.linspace(-2, 2, 5001).astype(np.float32)
    X = ek.Float32.copy(x)
    y = X * X + 1.0
    assert ek.lib().ek_var_ptr(y.index) is None
    assert y.coeff(17) == np.float32(x[17] * x[17]) + np.float32(1.0)
    assert ek.lib().ek_var_ptr(y.index) is not None
    s = ek.hsum(ek.UInt32.copy(np.arange(100, dtype=np.uint32)))
    assert int(s.coeff(0)) == 4950
    out = np.zeros(1, np.float32)
    assert ek.lib().ek_fetch_element(out.ctypes.data, y.index, 10**9, 4) != 0          # out of bounds -> error, no crash
    assert b"out of bounds" in ek.lib().ek_last_error()


def test_pre_eval_callbacks(gpu):
    """cuda_register_callback (jit.cu:1421-1422): callbacks run at the start of every cuda_eval(); unregistering an
    unknown entry is an error (jit.cu:1734-1741)."""
    ek = gpu
    L = ek.lib()
    CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
    hits = []
    cb = CB(lambda payload: hits.append(payload))
    L.ek_register_callback.argtypes = [CB, ctypes.c_void_p]; L.ek_unregister_callback.argtypes = [CB, ctypes.c_void_p]
    assert L.ek_register_callback(cb, ctypes.c_void_p(1234)) == 0
    try:
        a = ek.Float32.copy(np.ones(100, np.float32)) + 1.0
        ek.cuda_eval()
        assert hits == [1234]
        ek.cuda_eval()                                        # nothing pending: the callback still runs (jit.cu:1421)
        assert hits == [1234, 1234]
    finally:
        assert L.ek_unregister_callback(cb, ctypes.c_void_p(1234)) == 0
    assert L.ek_unregister_callback(cb, ctypes.c_void_p(1234)) != 0
    n = len(hits)
    b = a * 2.0
    ek.cuda_eval()
    assert len(hits) == n and b.coeff(3) == np.float32(4.0)


def test_kernel_qualification_verdict(gpu):
    """ek_init() decided between the fast and the general sweep kernel (csrc/ek_runtime.cpp "kernel qualification").
    Whatever the verdict, wide 32-bit sweeps must run, be bit-exact, and use the kernel the verdict names.  When the verdict
    is negative the helper is run once more in the foreground so that the test log shows which comparison disagreed."""
    ek = gpu
    L = ek.lib()
    mode = L.ek_fast_mode()
    x = np.random.default_rng(5).uniform(-3, 3, 300_001).astype(np.float32)
    L.ek_stats_reset()
    y = (ek.Float32.copy(x) * 1.25 + 0.5).numpy()
    assert (y == x * np.float32(1.25) + np.float32(0.5)).all()
    st = ek.stats()
    assert int(st.sweep_launches) >= 1
    assert (int(st.fast_launches) > 0) == bool(mode)
    if not mode and os.environ.get("EK_FAST") is None:
        helper = os.path.join(ROOT, "enoki_b200", "ek_qualify")
        r = subprocess.run([helper], capture_output=True, text=True, timeout=600, env=dict(os.environ, EK_FAST="0"))
        sys.stderr.write(r.stderr[-3000:])
        pytest.skip("the fast sweep kernel did NOT qualify on this GPU (general kernels are in use); ek_qualify says: " + r.stderr[-1500:])


@pytest.mark.parametrize("n,k", [(1, 1), (31, 3), (100_003, 7), (1 << 22, 40), (50_000, 1500)])
def test_partition(gpu, n, k):
    """cuda_partition (horiz.cu:35-122) feeds virtual-call dispatch: groups of indices per distinct pointer, pointers
    ascending, indices ascending inside a group (what the reference's stable radix sort + RLE produces).
    Default path: the host-side stable sort (csrc/ek_scan.cu); the device composition is test_partition_device_composition."""
    ek = gpu
    L = ek.lib()
    rng = np.random.default_rng(n + k)
    # k distinct 16-byte aligned "pointers" below 2^44.  (NEVER materialise the value range: the first version of this test
    # said rng.choice(np.arange(1, 1 << 40)) -- an 8 TiB host array -- and every GPU box that ran it was lost to the host's
    # OOM killer after ~2 minutes; that, not ek_partition, is what "hung" in rounds 1 and 2.)
    cand = np.unique(rng.integers(1, 1 << 40, size=4 * k + 16, dtype=np.uint64))
    assert len(cand) >= k
    table = np.sort(rng.permutation(cand)[:k]) * np.uint64(16)
    ptr = table[rng.integers(0, k, n)]
    P64 = ek.UInt64.copy(ptr)
    uniq = ctypes.c_void_p(); counts = ctypes.c_void_p(); perm = ctypes.c_void_p()
    assert L.ek_partition(n, P64.data(), ctypes.byref(uniq), ctypes.byref(counts), ctypes.byref(perm)) == 0, L.ek_last_error()
    cnt = np.ctypeslib.as_array(ctypes.cast(counts.value, ctypes.POINTER(ctypes.c_uint32)), shape=(1,))
    K = int(cnt[0])
    want_u, want_c = np.unique(ptr, return_counts=True)
    assert K == len(want_u)
    cnt = np.ctypeslib.as_array(ctypes.cast(counts.value, ctypes.POINTER(ctypes.c_uint32)), shape=(K + 1,)).copy()
    un = np.ctypeslib.as_array(ctypes.cast(uniq.value, ctypes.POINTER(ctypes.c_uint64)), shape=(K,)).copy()
    assert (un == want_u).all() and (cnt[1:] == want_c).all()
    pp = np.ctypeslib.as_array(ctypes.cast(perm.value, ctypes.POINTER(ctypes.c_uint64)), shape=(K,)).copy()
    order = np.argsort(ptr, kind="stable").astype(np.uint32)
    off = 0
    for i in range(K):
        got = ek.UInt32.map(int(pp[i]), int(cnt[i + 1]), True).numpy()
        assert (got == order[off:off + cnt[i + 1]]).all(), i
        off += int(cnt[i + 1])
    L.ek_host_free(uniq); L.ek_host_free(counts)
    libc = ctypes.CDLL(None); libc.free.argtypes = [ctypes.c_void_p]; libc.free(perm)


@pytest.mark.parametrize("n", ["1000", "100003", "4194304"])
def test_cpp_virtual_call_dispatch(gpu, n):
    """ENOKI_CALL_SUPPORT dispatch through CUDAArray<T *>::partition_() -> ek_partition, compared bit
    for bit with the same classes called on the reference CPU path (array_call.h:124-193)."""
    binp = os.path.join(os.path.dirname(BIN), "call_check")
    if not os.path.exists(binp):
        pytest.skip("tests/cpp/call_check not built (needs the reference headers at build time)")
    r = subprocess.run([binp, n], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_tape_fp64(gpu):
    """Tape<CUDAArray<double>> (autodiff.cpp:1239-1240) through the C ABI: layered graph with dyadic weights (every product
    and sum is exact in binary64, so the accumulation order does not matter) against numpy float64 -- bit for bit.
    Exercises ek_adjoint_kernel<double>, the fp64 zero guards (the reference's safe_* PTX hard-codes f32,
    autodiff.cpp:1200-1202; here the guards follow the tape's type) and fp64 reductions of the loss edges."""
    ek = gpu
    L = ek.lib()
    F64 = 11
    Lv, K, w = 6, 8, 5000
    rng = np.random.default_rng(64)
    vals = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, 0.25, -2.0, 1.5])
    L.ek_tape_set_graph_simplification(F64, 0)
    try:
        ids = [[L.ek_tape_append_leaf(F64, w) for _ in range(K)]]
        edges = []                                   # (src level, src k, dst level, dst k, weights)
        keep = []
        for lvl in range(1, Lv):
            row = []
            for k in range(K):
                nid = L.ek_tape_append_node(F64, w, b"n")
                for p in sorted(rng.choice(K, 2, replace=False).tolist()):
                    wt = vals[rng.integers(0, len(vals), w)].astype(np.float64)
                    wv = ek.Float64.copy(wt); keep.append(wv)
                    assert L.ek_tape_append_edge(F64, ids[lvl - 1][p], nid, wv.index) == 0
                    edges.append((lvl - 1, p, lvl, k, wt))
                row.append(nid)
            ids.append(row)
        one = ek.Float64.copy(np.ones(1, np.float64))
        loss = L.ek_tape_append_node(F64, 1, b"loss")
        for k in range(K):
            assert L.ek_tape_append_edge(F64, ids[Lv - 1][k], loss, one.index) == 0
        assert L.ek_tape_backward(F64, loss, 1) == 0, L.ek_last_error()
        got = []
        for k in range(K):
            h = L.ek_tape_gradient(F64, ids[0][k])
            if h == 0:
                got.append(np.zeros(w)); continue
            L.ek_inc_ref_ext(h)
            a = ek.Float64.from_index(h).numpy()
            got.append(np.broadcast_to(a, (w,)).copy() if a.size == 1 else a)
        for row in ids:
            for nid in row:
                L.ek_tape_dec_ref_ext(F64, nid)
        L.ek_tape_dec_ref_ext(F64, loss)
    finally:
        L.ek_tape_set_graph_simplification(F64, 1)
    grad = {(Lv - 1, k): np.ones(w) for k in range(K)}
    for lvl in range(Lv - 1, 0, -1):
        for (sl, sk, dl, dk, wt) in edges:
            if dl != lvl or (dl, dk) not in grad:
                continue
            grad[(sl, sk)] = grad.get((sl, sk), np.zeros(w)) + wt * grad[(dl, dk)]
    for k in range(K):
        want = grad.get((0, k), np.zeros(w))
        assert (got[k] == want).all(), (k, np.abs(got[k] - want).max())
    assert L.ek_tape_node_count(F64) == 0


def test_tape_full_size_c4_sampled_columns(gpu, oracle, P):
    """The BENCHED C4 tape (BASELINE configs[3]: 80 levels x 128 nodes, node width 131 072, 20 224 + 128 edges) against
    the oracle on a strided sample of element columns.  Columns are independent (every edge weight is element-wise), so
    the oracle runs the same graph at width 64 on columns 0, 2048, 4096, ... and has to reproduce the device's leaf
    gradients on those columns BIT FOR BIT -- at this size the level-batched descriptor uploads and the grid-stride path
    of the adjoint kernel are what runs (parity used to stop at 8 x 64 x 20 000)."""
    ek = gpu
    L = ek.lib()
    Lv, K, w, stride = 80, 128, 131072, 2048
    cols = np.arange(0, w, stride)
    nc = len(cols)
    rng = np.random.default_rng(2024)
    L.ek_tape_set_graph_simplification(F32, 0)
    try:
        ids = [[L.ek_tape_append_leaf(F32, w) for _ in range(K)]]
        src, dst, ws_small, keep = [], [], [], []
        for lvl in range(1, Lv):
            row = []
            for k in range(K):
                nid = L.ek_tape_append_node(F32, w, b"n")
                picks = sorted(rng.choice(K, 2, replace=False).tolist())
                for p in picks:
                    wt = rng.random(w, dtype=np.float32) + np.float32(0.5)
                    wt[rng.random(w, dtype=np.float32) < 0.01] = 0.0       # exact zeros: safe_mul / safe_fmadd
                    wv = ek.Float32.copy(wt)
                    keep.append(wv)
                    assert L.ek_tape_append_edge(F32, ids[lvl - 1][p], nid, wv.index) == 0
                    src.append((lvl - 1) * K + p + 1); dst.append(lvl * K + k + 1); ws_small.append(wt[cols].copy())
                row.append(nid)
            ids.append(row)
            ek.cuda_eval()
            del keep[:]                                                    # the tape holds the weights now
        one = ek.Float32.copy(np.ones(1, np.float32))
        loss = L.ek_tape_append_node(F32, 1, b"loss")
        for k in range(K):
            assert L.ek_tape_append_edge(F32, ids[Lv - 1][k], loss, one.index) == 0
            src.append((Lv - 1) * K + k + 1); dst.append(Lv * K + 1); ws_small.append(np.ones(1, np.float32))
        assert L.ek_tape_backward(F32, loss, 1) == 0, L.ek_last_error()
        got = []
        for k in range(K):
            h = L.ek_tape_gradient(F32, ids[0][k])
            assert h != 0
            L.ek_inc_ref_ext(h)
            got.append(ek.Float32.from_index(h).numpy()[cols])
        got = np.concatenate(got)
        for row in ids:
            for nid in row:
                L.ek_tape_dec_ref_ext(F32, nid)
        L.ek_tape_dec_ref_ext(F32, loss)
    finally:
        L.ek_tape_set_graph_simplification(F32, 1)
    # the same graph at width nc for the oracle
    n_nodes = Lv * K + 1
    node_size = np.full(n_nodes, nc, np.uint32); node_size[-1] = 1
    wsize = np.array([len(a) for a in ws_small], np.uint32)
    woff = np.concatenate([[0], np.cumsum(wsize[:-1], dtype=np.uint64)]).astype(np.uint64)
    weights = np.concatenate(ws_small)
    want_ids = np.arange(1, K + 1, dtype=np.uint32)
    want = np.zeros(K * nc, np.float32)
    src_a, dst_a = np.array(src, np.uint32), np.array(dst, np.uint32)
    rc = oracle.or_tape_backward(n_nodes, P(node_size), len(src_a), P(src_a), P(dst_a), P(weights), P(woff), P(wsize),
                                 n_nodes, K, P(want_ids), P(want))
    assert rc == 0
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), (np.abs(got - want).max(), int((got != want).sum()))
    assert L.ek_tape_node_count(F32) == 0


def test_malloc_trim_releases_cached_blocks(gpu):
    """cuda_malloc_trim (jit.cu:1871-1896, user-visible through tests/python/test_pytorch.py:16,28): freed blocks stay in the
    size-class free lists until a trim hands them back to the driver."""
    ek = gpu
    L = ek.lib()
    free0, total = ctypes.c_size_t(), ctypes.c_size_t()
    ek.cuda_sync(); L.ek_malloc_trim()
    L.ek_mem_get_info(ctypes.byref(free0), ctypes.byref(total))
    big = [ek.Float32.copy(np.zeros(1 << 24, np.float32)) for _ in range(4)]        # 4 x 64 MiB
    for b in big:
        b.eval()
    ek.cuda_sync()
    free1 = ctypes.c_size_t(); L.ek_mem_get_info(ctypes.byref(free1), ctypes.byref(total))
    assert free0.value - free1.value >= 200 << 20
    del big, b
    ek.cuda_sync()
    free2 = ctypes.c_size_t(); L.ek_mem_get_info(ctypes.byref(free2), ctypes.byref(total))
    assert free0.value - free2.value >= 200 << 20            # cached, not returned
    L.ek_malloc_trim()
    free3 = ctypes.c_size_t(); L.ek_mem_get_info(ctypes.byref(free3), ctypes.byref(total))
    assert free0.value - free3.value < 64 << 20              # handed back
    # the allocator keeps working after a trim
    assert float(ek.hsum(ek.Float32.copy(np.ones(1000, np.float32))).coeff(0)) == 1000.0


def test_make_managed_keeps_the_values(gpu):
    """cuda_make_managed (jit.cu:455-485): the array moves to managed memory; the host may read it after a sync."""
    ek = gpu
    L = ek.lib()
    x = np.random.default_rng(3).uniform(-1, 1, 4096).astype(np.float32)
    y = ek.Float32.copy(x) * 3.0
    assert L.ek_make_managed(y.index) == 0, L.ek_last_error()
    ek.cuda_sync()
    p = L.ek_var_ptr(y.index)
    host_view = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(4096,))
    assert (host_view == x * np.float32(3.0)).all()
    assert L.ek_make_managed(y.index) == 0                    # idempotent


def test_partition_device_composition(gpu):
    """EK_PARTITION_DEVICE=1: cuda_partition composed from hmin / select / compress on the device (csrc/ek_scan.cu) instead of
    the host sort that is the default; run in a child process because the switch is read from the environment."""
    code = r'''
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
import enoki_b200 as ek
L = ek.lib()
rng = np.random.default_rng(11)
n, k = 200_003, 9
cand = np.unique(rng.integers(1, 1 << 40, size=4 * k + 16, dtype=np.uint64))
table = np.sort(rng.permutation(cand)[:k]) * np.uint64(16)
ptr = table[rng.integers(0, k, n)]
P64 = ek.UInt64.copy(ptr)
uniq = ctypes.c_void_p(); counts = ctypes.c_void_p(); perm = ctypes.c_void_p()
assert L.ek_partition(n, P64.data(), ctypes.byref(uniq), ctypes.byref(counts), ctypes.byref(perm)) == 0, L.ek_last_error()
K = int(np.ctypeslib.as_array(ctypes.cast(counts.value, ctypes.POINTER(ctypes.c_uint32)), shape=(1,))[0])
assert K == k
cnt = np.ctypeslib.as_array(ctypes.cast(counts.value, ctypes.POINTER(ctypes.c_uint32)), shape=(K + 1,)).copy()
un = np.ctypeslib.as_array(ctypes.cast(uniq.value, ctypes.POINTER(ctypes.c_uint64)), shape=(K,)).copy()
want_u, want_c = np.unique(ptr, return_counts=True)
assert (un == want_u).all() and (cnt[1:] == want_c).all()
pp = np.ctypeslib.as_array(ctypes.cast(perm.value, ctypes.POINTER(ctypes.c_uint64)), shape=(K,)).copy()
order = np.argsort(ptr, kind="stable").astype(np.uint32)
off = 0
for i in range(K):
    got = ek.UInt32.map(int(pp[i]), int(cnt[i + 1]), True).numpy()
    assert (got == order[off:off + cnt[i + 1]]).all(), i
    off += int(cnt[i + 1])
print("device partition ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, EK_PARTITION_DEVICE="1"))
    assert r.returncode == 0 and "device partition ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_rank_sharded_matches_single_gpu(gpu):
    """SURVEY 8e on hardware (needs 2 GPUs, skipped otherwise): element-range sharding over two ranks + ONE all-reduce of the
    size-1 results reproduces the single-GPU loss and scalar gradients within the reassociation bound -- through
    torch.distributed (the round-1 transport) and through the library's own NCCL communicator (C ABI, csrc/ek_dist.cpp)."""
    import json
    try:
        import torch
        n_gpu = torch.cuda.device_count()
    except Exception:
        n_gpu = 0
    if n_gpu < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "two_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = next((l for l in r.stdout.splitlines() if l.startswith("TWO_RANK_RESULT ")), None)
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[len("TWO_RANK_RESULT "):])
    single = np.array(res["single"], np.float64)
    scale = np.maximum(np.abs(single), 1.0)
    assert (np.abs(np.array(res["torch"]) - single) <= 1e-5 * scale).all(), res
    assert res["native"] is not None, "the native NCCL communicator could not be created (see the worker's stderr): " + r.stderr[-1500:]
    assert (np.abs(np.array(res["native"]) - single) <= 1e-5 * scale).all(), res
