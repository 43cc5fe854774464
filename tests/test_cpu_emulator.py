"""CPU: the planner + assembler executed by a numpy interpreter of the sweep ISA (tests/ek_emulator.py) on random
expression DAGs -- the same generator as the GPU fuzz test, no GPU needed."""
import numpy as np
import pytest

from ek_emulator import Emulator, Unsupported
import test_gpu_fuzz as fuzz


@pytest.fixture(autouse=True, params=["assembled", "lowered"])
def program_form(request):
    """Every test of this module runs twice: on the assembler's program and on the program lowered for the 32-bit fast
    kernel (ek_eval.cpp lower_fast -> ek_sweep_fast.cu), raised back by ek_emulator.raise_fast."""
    Emulator.prefer_fast = request.param == "lowered"
    yield request.param
    Emulator.prefer_fast = False


class _Factory:
    """Stands in for Float32.copy / UInt32.copy: a fake device mapping plus the data the emulator reads."""
    def __init__(self, cls, table, base):
        self.cls, self.table, self.base = cls, table, base

    addresses = {}                              # device address -> variable index (all factories)

    def copy(self, a):
        address = self.base + 0x1000000 * (len(self.table) + 1)
        r = self.cls.map(address, len(a))
        self.table[r.index] = np.ascontiguousarray(a)
        _Factory.addresses[address] = r.index
        return r

    def __call__(self, *args):
        return self.cls(*args)

    def __getattr__(self, name):
        return getattr(self.cls, name)


def _case(ek, oracle, P, seed):
    """Returns None or a skip reason.  Runs in its own frame so that no array handle outlives the case (an
    unevaluated handle kept alive by a traceback would leak its trace into the next case's plan)."""
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([1, 33, 1000, 4097]))
    floats = [rng.uniform(-4, 4, n).astype(np.float32) for _ in range(3)] + [np.array([rng.uniform(-2, 2)], np.float32)]
    ints = [rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(2)] + [np.array([rng.integers(1, 100)], np.uint32)]
    table = {}

    class EK:                                       # the slice of the module API the generator uses
        Float32 = _Factory(ek.Float32, table, 0x7f0000000000)
        UInt32 = _Factory(ek.UInt32, table, 0x7a0000000000)
        fmadd, max_, min_, sqrt, floor, sin, exp, select = ek.fmadd, ek.max_, ek.min_, ek.sqrt, ek.floor, ek.sin, ek.exp, ek.select

    fl, it = fuzz._build(EK, oracle, P, rng, n, int(rng.integers(8, 40)), floats, ints)
    keep_f = [fl[k] for k in rng.choice(len(fl), size=min(4, len(fl)), replace=False)]
    keep_i = [it[k] for k in rng.choice(len(it), size=min(3, len(it)), replace=False)]
    red_i = ek.hsum(keep_i[0].e); red_f = ek.hsum(keep_f[0].e)
    del fl, it
    prog = ek.debug_program()
    emu = Emulator(oracle, table)
    try:
        emu.run(prog)
    except Unsupported as e:
        return f"emulator: {e}"

    def value(m):
        if m.e.index in emu.vars:
            return emu.vars[m.e.index]
        raise AssertionError(f"variable {m.e.index} was not produced by any sweep")

    for m in keep_f:
        got = value(m); want = np.broadcast_to(m.n, got.shape).astype(np.float32)
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (seed, n, got[~same][:3], want[~same][:3])
    for m in keep_i:
        got = value(m).view(np.uint32); want = np.broadcast_to(m.n, got.shape)
        assert (got == want).all(), (seed, n)
    gi = emu.vars[red_i.index].view(np.uint32)[0]
    assert gi == np.uint32(int(np.broadcast_to(keep_i[0].n, (len(keep_i[0].n),)).astype(np.uint64).sum()) & 0xffffffff)
    wf = np.asarray(keep_f[0].n, np.float64)
    if np.isfinite(wf).all():
        assert abs(float(emu.vars[red_f.index][0]) - wf.sum()) <= 2e-6 * (np.abs(wf).sum() + 1e-30) + 1e-6 * abs(wf.sum())
    return None


@pytest.mark.parametrize("seed", range(150))
def test_random_expression_dags_on_the_emulator(ek, oracle, P, seed):
    import gc
    gc.collect()
    ek.lib().ek_debug_discard_side_effects()       # scatters recorded by earlier CPU tests can never run here
    gc.collect()
    assert ek.debug_plan() == "", "unevaluated variables of an earlier test are still alive"
    reason = _case(ek, oracle, P, seed)
    gc.collect()
    if reason:
        pytest.skip(reason)


# ------------------------------------------------------------------ 64-bit values (lo / hi planes)
def _build64(EK, oracle, P, rng, n, n_nodes, data):
    M = fuzz.Mirror
    D, U, I = EK.Float64, EK.UInt64, EK.Int64
    dl = [M(D.copy(a), a) for a in data["f64"]]
    ul = [M(U.copy(a), a) for a in data["u64"]]
    fl = [M(EK.Float32.copy(a), a) for a in data["f32"]]
    SZ = fuzz.SZ
    bc = lambda a: np.ascontiguousarray(np.broadcast_to(a, (n,)) if a.shape[0] == 1 else a)

    def orc_unary(which, x):
        x = np.ascontiguousarray(x); out = np.zeros_like(x)
        oracle.or_unary_f64(which, P(x), P(out), SZ(len(x))); return out
    for _ in range(n_nodes):
        kind = rng.integers(0, 15)
        a, b, c = (dl[rng.integers(len(dl))] for _ in range(3))
        u, v = (ul[rng.integers(len(ul))] for _ in range(2))
        with np.errstate(all="ignore"):
            if kind == 0: dl.append(M(a.e + b.e, a.n + b.n))
            elif kind == 1: dl.append(M(a.e * b.e - c.e, a.n * b.n - c.n))
            elif kind == 2:
                an, bn, cn = bc(a.n), bc(b.n), bc(c.n); out = np.zeros(n)
                oracle.or_fma_f64(P(an), P(bn), P(cn), P(out), SZ(n))
                dl.append(M(EK.fmadd(a.e, b.e, c.e), out if max(len(a.n), len(b.n), len(c.n)) > 1 else out[:1]))
            elif kind == 3: dl.append(M(EK.sqrt(abs(a.e)), np.sqrt(np.abs(a.n))))
            elif kind == 4:
                which, fn = ((0, EK.sin), (2, EK.exp))[rng.integers(2)]
                dl.append(M(fn(a.e), orc_unary(which, a.n)))
            elif kind == 5: dl.append(M(EK.select(a.e < b.e, a.e, -c.e), np.where(a.n < b.n, a.n, -c.n)))
            elif kind == 6: dl.append(M(EK.floor(a.e), np.floor(a.n)))
            elif kind == 7: ul.append(M(u.e + v.e * U(3), u.n + v.n * np.uint64(3)))
            elif kind == 8: ul.append(M((u.e ^ v.e) | (u.e & v.e), (u.n ^ v.n) | (u.n & v.n)))
            elif kind == 9:
                s = int(rng.integers(1, 63))
                ul.append(M((u.e << U(s)) | (v.e >> U(s)), (u.n << np.uint64(s)) | (v.n >> np.uint64(s))))
            elif kind == 10: ul.append(M(EK.select(u.e < v.e, u.e, v.e), np.where(u.n < v.n, u.n, v.n)))
            elif kind == 11:                                    # u64 -> f64 -> scaled -> i64 -> u64 (bounded)
                x = D(u.e >> U(20)) * D(0.5); xn = (u.n >> np.uint64(20)).astype(np.float64) * 0.5
                dl.append(M(x, xn))
                ul.append(M(U(I(x)), np.trunc(xn).astype(np.int64).view(np.uint64)))
            elif kind == 12:                                    # narrowing and widening
                ul.append(M(U(EK.UInt32(u.e)) + v.e, (u.n & np.uint64(0xffffffff)) + v.n))
            elif kind == 13:                                    # f32 <-> f64
                g = fl[rng.integers(len(fl))]
                dl.append(M(D(g.e) * a.e, g.n.astype(np.float64) * a.n))
                fl.append(M(EK.Float32(a.e), a.n.astype(np.float32)))
            elif kind == 14: ul.append(M(u.e - v.e, u.n - v.n))
    return dl, ul, fl


def _case64(ek, oracle, P, seed):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([1, 33, 1000, 4097]))
    data = {"f64": [rng.uniform(-4, 4, n) for _ in range(3)] + [np.array([rng.uniform(-2, 2)])],
            "u64": [rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + np.uint64(rng.integers(2)) for _ in range(2)] + [np.array([rng.integers(1, 100)], np.uint64)],
            "f32": [rng.uniform(-4, 4, n).astype(np.float32)]}
    table = {}

    class EK:
        Float32 = _Factory(ek.Float32, table, 0x7f0000000000)
        UInt32 = _Factory(ek.UInt32, table, 0x7a0000000000)
        Float64 = _Factory(ek.Float64, table, 0x790000000000)
        UInt64 = _Factory(ek.UInt64, table, 0x780000000000)
        Int64 = _Factory(ek.Int64, table, 0x770000000000)
        fmadd, max_, min_, sqrt, floor, sin, exp, select = ek.fmadd, ek.max_, ek.min_, ek.sqrt, ek.floor, ek.sin, ek.exp, ek.select

    dl, ul, fl = _build64(EK, oracle, P, rng, n, int(rng.integers(6, 30)), data)
    keep = [dl[k] for k in rng.choice(len(dl), size=min(3, len(dl)), replace=False)] + \
           [ul[k] for k in rng.choice(len(ul), size=min(3, len(ul)), replace=False)] + [fl[-1]]
    del dl, ul, fl
    emu = Emulator(oracle, table)
    try:
        emu.run(ek.debug_program())
    except Unsupported as e:
        return f"emulator: {e}"
    for m in keep:
        if m.e.index not in emu.vars:
            raise AssertionError(f"variable {m.e.index} was not produced by any sweep")
        got = emu.vars[m.e.index]; want = np.broadcast_to(m.n, got.shape).astype(got.dtype)
        bits = np.uint64 if got.dtype.itemsize == 8 else np.uint32
        same = got.view(bits) == np.ascontiguousarray(want).view(bits)
        if got.dtype.kind == "f":
            same |= np.isnan(got) & np.isnan(want)
        assert same.all(), (seed, n, got.dtype, got[~same][:3], want[~same][:3])
    return None


@pytest.mark.parametrize("seed", range(100))
def test_random_64bit_dags_on_the_emulator(ek, oracle, P, seed):
    import gc
    gc.collect()
    ek.lib().ek_debug_discard_side_effects()
    gc.collect()
    assert ek.debug_plan() == "", "unevaluated variables of an earlier test are still alive"
    reason = _case64(ek, oracle, P, seed)
    gc.collect()
    if reason:
        pytest.skip(reason)


def _case_many_outputs(ek, oracle, P):
    """More live results than one kernel has slots / arguments for: the planner splits the group and recomputes shared
    sub-expressions (as the reference does across kernels); every output must still be right."""
    rng = np.random.default_rng(77)
    n = 1000
    table = {}
    F = _Factory(ek.Float32, table, 0x7f0000000000)
    xs = [rng.uniform(-2, 2, n).astype(np.float32) for _ in range(6)]
    X = [F.copy(a) for a in xs]
    outs, want = [], []
    base_e = ek.fmadd(X[0], X[1], X[2]); base_n = None
    an, bn, cn = xs[0], xs[1], xs[2]
    base_n = np.zeros(n, np.float32); oracle.or_fma_f32(P(an), P(bn), P(cn), P(base_n), fuzz.SZ(n))
    for k in range(300):
        a = X[k % 6]; an = xs[k % 6]
        e = (base_e + a) * ek.Float32(float(k + 1)) - X[(k + 3) % 6]
        w = ((base_n + an) * np.float32(k + 1) - xs[(k + 3) % 6]).astype(np.float32)
        outs.append(e); want.append(w)
    del base_e
    prog = ek.debug_program()
    n_sweeps = len(prog["sweeps"])
    emu = Emulator(oracle, table)
    emu.run(prog)
    for e, w in zip(outs, want):
        got = emu.vars[e.index]
        assert (got.view(np.uint32) == w.view(np.uint32)).all()
    return n_sweeps


def test_group_splitting_on_the_emulator(ek, oracle, P):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert ek.debug_plan() == ""
    n_sweeps = _case_many_outputs(ek, oracle, P)
    gc.collect()
    assert n_sweeps >= 2            # 300 outputs do not fit one kernel's argument words: the group was split


def _case_histogram(ek, oracle, P, n_bins, n):
    """C3 shape (tests/histogram.cpp:41-57): idx = UInt32((y + 4) * n_bins / 8); mask = idx < n_bins;
    w = gather(table, idx, mask); scatter_add(bins_u32, 1, idx, mask); scatter_add(hist_f32, w, idx, mask)."""
    from enoki_b200 import Float32, UInt32, gather, scatter_add
    rng = np.random.default_rng(n_bins)
    table = {}
    F = _Factory(Float32, table, 0x7f0000000000); U = _Factory(UInt32, table, 0x7a0000000000)
    y_n = rng.normal(0, 1.3, n).astype(np.float32)
    tab_n = np.linspace(0.5, 1.5, n_bins, dtype=np.float32)
    y = F.copy(y_n); tab = F.copy(tab_n)
    bins = U.copy(np.zeros(n_bins, np.uint32)); hist = F.copy(np.zeros(n_bins, np.float32))
    idx = UInt32((y + 4.0) * float(n_bins) / 8.0)
    mask = idx < UInt32(n_bins)
    w = gather(Float32, tab, idx, mask)
    scatter_add(bins, UInt32(1), idx, mask)
    scatter_add(hist, w, idx, mask)
    del idx, mask, w
    plan = ek.debug_plan()
    emu = Emulator(oracle, table, _Factory.addresses)
    emu.run(ek.debug_program())
    with np.errstate(all="ignore"):
        t = ((y_n + np.float32(4.0)) * np.float32(n_bins)) / np.float32(8.0)
        ok = (t > -9.2e18) & (t < 9.2e18)
        idx_n = np.where(ok, np.trunc(np.where(ok, t, 0)).astype(np.int64) & 0xffffffff, 0).astype(np.uint32)
    m = idx_n < n_bins
    want_bins = np.bincount(idx_n[m], minlength=n_bins).astype(np.uint32)
    want_hist = np.bincount(idx_n[m], weights=tab_n[idx_n[m]].astype(np.float64), minlength=n_bins)
    got_bins = emu.vars[bins.index].view(np.uint32); got_hist = emu.vars[hist.index].view(np.float32)
    assert (got_bins == want_bins).all()
    assert np.allclose(got_hist, want_hist, rtol=1e-4, atol=1e-3)
    ek.lib().ek_debug_discard_side_effects()
    return plan


@pytest.mark.parametrize("n_bins,n", [(31, 5000), (700, 4000), (5000, 3000)])
def test_histogram_programs_on_the_emulator(ek, oracle, P, n_bins, n):
    """31 and 700 bins use shared-memory bins (+ a staged table for 31); 5000 bins go to global atomics."""
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert ek.debug_plan() == ""
    plan = _case_histogram(ek, oracle, P, n_bins, n)
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    if n_bins <= 1024:
        assert "SCATTER_ADD_I32_SMEM" in plan and "SMEM_FLUSH_ADD_I32" in plan
    else:
        assert "SCATTER_ADD_I32 " in plan and "SMEM_ZERO" not in plan


# ------------------------------------------------------------------ signed integers, masks, conversions
def _build_int(EK, rng, n, n_nodes, data):
    M = fuzz.Mirror
    I, U, F, B = EK.Int32, EK.UInt32, EK.Float32, EK.Mask
    il = [M(I.copy(a), a) for a in data["i32"]]
    ul = [M(U.copy(a), a) for a in data["u32"]]
    fl = [M(F.copy(a), a) for a in data["f32"]]
    ml = []
    for _ in range(n_nodes):
        kind = rng.integers(0, 14)
        a, b = (il[rng.integers(len(il))] for _ in range(2))
        u, v = (ul[rng.integers(len(ul))] for _ in range(2))
        f, g = (fl[rng.integers(len(fl))] for _ in range(2))
        with np.errstate(all="ignore"):
            if kind == 0: il.append(M(a.e + b.e * I(3), a.n + b.n * np.int32(3)))
            elif kind == 1: il.append(M(-a.e - b.e, -a.n - b.n))
            elif kind == 2: il.append(M(EK.min_(a.e, b.e) ^ EK.max_(a.e, b.e), np.minimum(a.n, b.n) ^ np.maximum(a.n, b.n)))
            elif kind == 3:
                s = int(rng.integers(1, 31))
                il.append(M(a.e >> I(s), a.n >> np.int32(s)))                       # arithmetic shift
            elif kind == 4: il.append(M(abs(a.e), np.abs(a.n)))
            elif kind == 5: ml.append(M(a.e < b.e, a.n < b.n))
            elif kind == 6: ml.append(M(u.e >= v.e, u.n >= v.n))
            elif kind == 7: ml.append(M(f.e <= g.e, f.n <= g.n))
            elif kind == 8 and len(ml) >= 2:
                p, q = (ml[rng.integers(len(ml))] for _ in range(2))
                ml.append(M((p.e & q.e) | ~p.e, (p.n & q.n) | ~p.n))
            elif kind == 9 and ml:
                p = ml[rng.integers(len(ml))]
                il.append(M(EK.select(p.e, a.e, b.e), np.where(p.n, a.n, b.n)))
                fl.append(M(EK.select(p.e, f.e, g.e), np.where(p.n, f.n, g.n)))
            elif kind == 10: fl.append(M(F(a.e), a.n.astype(np.float32)))
            elif kind == 11:
                x = EK.max_(EK.min_(f.e, F(3.0e4)), F(-3.0e4))
                xn = np.maximum(np.minimum(np.where(np.isnan(f.n), np.float32(3.0e4), f.n), np.float32(3.0e4)), np.float32(-3.0e4))
                il.append(M(I(x), np.trunc(xn).astype(np.int32)))
            elif kind == 12: ul.append(M(U(a.e) + u.e, a.n.view(np.uint32) + u.n))
            elif kind == 13: ul.append(M(EK.mulhi(u.e, v.e), ((u.n.astype(np.uint64) * v.n.astype(np.uint64)) >> np.uint64(32)).astype(np.uint32)))
    return il, ul, fl, ml


def _case_int(ek, oracle, P, seed):
    rng = np.random.default_rng(12000 + seed)
    n = int(rng.choice([1, 65, 1000, 4097]))
    data = {"i32": [rng.integers(-2 ** 31, 2 ** 31, n, dtype=np.int64).astype(np.int32) for _ in range(2)] + [np.array([rng.integers(-50, 50)], np.int32)],
            "u32": [rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(2)],
            "f32": [rng.uniform(-1e5, 1e5, n).astype(np.float32), np.array([rng.uniform(-2, 2)], np.float32)]}
    table = {}

    class EK:
        Float32 = _Factory(ek.Float32, table, 0x7f0000000000)
        UInt32 = _Factory(ek.UInt32, table, 0x7a0000000000)
        Int32 = _Factory(ek.Int32, table, 0x760000000000)
        Mask = ek.Mask
        max_, min_, select, mulhi = ek.max_, ek.min_, ek.select, ek.mulhi

    il, ul, fl, ml = _build_int(EK, rng, n, int(rng.integers(8, 36)), data)
    pick = lambda lst, k: [lst[j] for j in rng.choice(len(lst), size=min(k, len(lst)), replace=False)] if lst else []
    keep = pick(il, 3) + pick(ul, 2) + pick(fl, 2) + pick(ml, 2)
    cnt = ek.hsum(il[-1].e)
    del il, ul, fl, ml
    emu = Emulator(oracle, table)
    try:
        emu.run(ek.debug_program())
    except Unsupported as e:
        return f"emulator: {e}"
    for m in keep:
        if m.e.index not in emu.vars:
            raise AssertionError(f"variable {m.e.index} was not produced by any sweep")
        got = emu.vars[m.e.index]; want = np.broadcast_to(m.n, got.shape)
        if got.dtype == np.bool_:
            assert (got == want).all(), (seed, n, "mask")
        else:
            same = np.ascontiguousarray(got).view(np.uint32) == np.ascontiguousarray(want.astype(got.dtype)).view(np.uint32)
            if got.dtype.kind == "f":
                same |= np.isnan(got) & np.isnan(want)
            assert same.all(), (seed, n, got.dtype, got[~same][:3], want[~same][:3])
    return None


@pytest.mark.parametrize("seed", range(100))
def test_random_integer_and_mask_dags_on_the_emulator(ek, oracle, P, seed):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert ek.debug_plan() == "", "unevaluated variables of an earlier test are still alive"
    reason = _case_int(ek, oracle, P, seed)
    gc.collect()
    if reason:
        pytest.skip(reason)


def _case_phases(ek, oracle, P):
    """Reductions feed later phases of the same evaluation: y = x / hsum(x), z = (x - hmin(x)) * (hmax(x) - x)."""
    rng = np.random.default_rng(3)
    n = 3001
    table = {}
    F = _Factory(ek.Float32, table, 0x7f0000000000); U = _Factory(ek.UInt32, table, 0x7a0000000000)
    xn = rng.uniform(0.5, 2.0, n).astype(np.float32); un = rng.integers(0, 1000, n, dtype=np.uint64).astype(np.uint32)
    x = F.copy(xn); u = U.copy(un)
    y = x / ek.hsum(x)
    z = (x - ek.hmin(x)) * (ek.hmax(x) - x)
    c = u + ek.hmax(u) - ek.hmin(u)
    p = ek.hprod(F.copy(np.linspace(0.9, 1.1, 9, dtype=np.float32)))
    prog = ek.debug_program()
    phases = sorted({sw["phase"] for sw in prog["sweeps"]})
    emu = Emulator(oracle, table)
    emu.run(prog)
    s = np.float32(xn.astype(np.float64).sum())
    assert np.allclose(emu.vars[y.index], xn / s, rtol=1e-6)
    zn = (xn - xn.min()) * (xn.max() - xn)
    assert (emu.vars[z.index].view(np.uint32) == zn.view(np.uint32)).all()
    assert (emu.vars[c.index].view(np.uint32) == un + un.max() - un.min()).all()
    assert abs(float(emu.vars[p.index][0]) - float(np.prod(np.linspace(0.9, 1.1, 9, dtype=np.float32).astype(np.float64)))) < 1e-6
    return phases


def test_reduction_phases_on_the_emulator(ek, oracle, P):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert ek.debug_plan() == ""
    phases = _case_phases(ek, oracle, P)
    gc.collect()
    assert phases == [0, 1]


def _case_simplified_weights(ek, oracle, P):
    """simplify_graph() (autodiff.cpp:990-1074) on a chain with a diamond: the surviving leaf -> root edge carries
    w0*w1*w2*(w4 + w3*w5) as a traced mul_nz / fma_nz expression; executed here by the emulator."""
    import ctypes
    lib = ek.lib(); F32 = ek.EK_FLOAT32
    n = 257
    rng = np.random.default_rng(11)
    table = {}
    F = _Factory(ek.Float32, table, 0x7f0000000000)
    wn = [rng.uniform(0.5, 1.5, n).astype(np.float32) for _ in range(6)]
    wn[1][::7] = 0.0                             # exact zeros exercise the *_nz guards
    ws = [F.copy(a) for a in wn]

    def node(label, srcs, wl):
        idx = (ctypes.c_uint32 * len(srcs))(*srcs); wh = (ctypes.c_uint32 * len(wl))(*[w.index for w in wl])
        h = lib.ek_tape_append(F32, label, n, len(srcs), idx, wh); assert h
        return h
    leaf = lib.ek_tape_append_leaf(F32, n)
    a = node(b"a", [leaf], [ws[0]]); b = node(b"b", [a], [ws[1]]); c = node(b"c", [b], [ws[2]])
    c2 = node(b"c2", [c], [ws[3]]); d = node(b"d", [c, c2], [ws[4], ws[5]])
    for h in (a, b, c, c2):
        lib.ek_tape_dec_ref_ext(F32, h)
    assert lib.ek_tape_simplify(F32) == 0 and lib.ek_tape_node_count(F32) == 2
    w = lib.ek_debug_tape_edge_weight(F32, leaf, d)
    assert w != 0
    emu = Emulator(oracle, table)
    emu.run(ek.debug_program())
    got = emu.vars[w].astype(np.float64)
    want = (wn[0].astype(np.float64) * wn[1] * wn[2]) * (wn[4].astype(np.float64) + wn[3].astype(np.float64) * wn[5])
    assert np.allclose(got, want, rtol=2e-6, atol=0) and (got[::7] == 0).all()
    lib.ek_tape_dec_ref_ext(F32, d); lib.ek_tape_dec_ref_ext(F32, leaf)
    assert lib.ek_tape_node_count(F32) == 0


def test_simplified_edge_weights_on_the_emulator(ek, oracle, P):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert ek.debug_plan() == ""
    _case_simplified_weights(ek, oracle, P)
    gc.collect()


def test_side_effects_run_in_recording_order(ek, oracle, P):
    """roots were planned by descending handle, so of two scatters into the same target the one
    recorded FIRST won.  The reference walks monotonically increasing ids (jit.cu:1385-1416): the later one wins, and a
    scatter_add recorded after a scatter adds to the scattered value.  Handles are recycled here (LIFO free list), so the
    test also churns the handle table first to make handle order differ from creation order."""
    import gc
    from enoki_b200 import Float32, UInt32, scatter, scatter_add
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    n, m = 1000, 64
    rng = np.random.default_rng(7)
    for churn in range(2):
        junk = [Float32(float(k)) for k in range(5 + 3 * churn)]       # allocate, then free in an order that scrambles the free list
        del junk[::2]; del junk
        table = {}
        F = _Factory(Float32, table, 0x7f0000000000); U = _Factory(UInt32, table, 0x7a0000000000)
        a_n, b_n = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
        i_n = rng.integers(0, m, n).astype(np.uint32)
        t1_n = np.zeros(m, np.float32); t2_n = np.zeros(m, np.float32)
        a, b, idx = F.copy(a_n), F.copy(b_n), U.copy(i_n)
        t1, t2 = F.copy(t1_n), F.copy(t2_n)
        scatter(t1, a * 2.0, idx)                   # recorded first
        scatter(t1, b * 3.0, idx)                   # recorded second: must win wherever both write
        scatter(t2, a, idx)
        scatter_add(t2, b, idx)                     # must see the scattered values
        emu = Emulator(oracle, table, _Factory.addresses)
        emu.run(ek.debug_program())
        want1 = t1_n.copy(); want1[i_n] = (a_n * np.float32(2.0)); want1[i_n] = (b_n * np.float32(3.0))
        assert (emu.vars[t1.index].view(np.uint32) == want1.view(np.uint32)).all()
        want2 = t2_n.copy(); want2[i_n] = a_n
        acc = want2.astype(np.float64); np.add.at(acc, i_n, b_n.astype(np.float64))
        assert np.allclose(emu.vars[t2.index], acc, rtol=1e-5, atol=1e-6)
        ek.lib().ek_debug_discard_side_effects()
        del a, b, idx, t1, t2
        gc.collect()
