"""GPU: random expression DAGs through the whole evaluator (planner, slot allocation, superinstructions, group splitting,
fused reductions) against a numpy / oracle mirror.  f32 arithmetic, sqrt, floor, sin, exp and all integer results
must be bit-exact; float reductions are compared with the fp64 sum."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SZ = ctypes.c_size_t


class Mirror:
    """A value that exists twice: as an enoki_b200 array and as the numpy array it has to equal."""
    def __init__(self, e, n):
        self.e, self.n = e, n


def _unary_oracle(oracle, P, which, x):
    out = np.zeros_like(x)
    oracle.or_unary_f32(which, P(np.ascontiguousarray(x)), P(out), SZ(len(x)))
    return out


def _build(ek, oracle, P, rng, n, n_nodes, floats, ints):
    F, U = ek.Float32, ek.UInt32
    fl = [Mirror(F.copy(a), a) for a in floats]
    it = [Mirror(U.copy(a), a) for a in ints]
    bc = lambda a: np.broadcast_to(a, (n,)) if a.shape[0] == 1 else a
    for _ in range(n_nodes):
        kind = rng.integers(0, 17)
        a, b, c = (fl[rng.integers(len(fl))] for _ in range(3))
        i, j = (it[rng.integers(len(it))] for _ in range(2))
        with np.errstate(all="ignore"):
            if kind == 0: fl.append(Mirror(a.e + b.e, a.n + b.n))
            elif kind == 1: fl.append(Mirror(a.e - b.e, a.n - b.n))
            elif kind == 2: fl.append(Mirror(a.e * b.e, a.n * b.n))
            elif kind == 3:
                an, bn, cn = (np.ascontiguousarray(bc(x.n)) for x in (a, b, c))
                out = np.zeros(n, np.float32); oracle.or_fma_f32(P(an), P(bn), P(cn), P(out), SZ(n))
                fl.append(Mirror(ek.fmadd(a.e, b.e, c.e), out if max(len(a.n), len(b.n), len(c.n)) > 1 else out[:1]))
            elif kind == 4:
                an, bn = (np.ascontiguousarray(bc(x.n)) for x in (a, b))
                is_max = int(rng.integers(2)); out = np.zeros(n, np.float32)
                oracle.or_minmax_f32(is_max, P(an), P(bn), P(out), SZ(n))
                fl.append(Mirror(ek.max_(a.e, b.e) if is_max else ek.min_(a.e, b.e), out if max(len(a.n), len(b.n)) > 1 else out[:1]))
            elif kind == 5: fl.append(Mirror(abs(a.e), np.abs(a.n)))
            elif kind == 6: fl.append(Mirror(-a.e, -a.n))
            elif kind == 7: fl.append(Mirror(ek.sqrt(abs(a.e)), np.sqrt(np.abs(a.n))))
            elif kind == 8: fl.append(Mirror(ek.floor(a.e), np.floor(a.n)))
            elif kind == 9:
                which, fn = ((0, ek.sin), (2, ek.exp))[rng.integers(2)]
                fl.append(Mirror(fn(a.e), _unary_oracle(oracle, P, which, a.n)))
            elif kind == 10: fl.append(Mirror(ek.select(a.e < b.e, a.e, c.e), np.where(a.n < b.n, a.n, c.n).astype(np.float32)))
            elif kind == 11: it.append(Mirror(i.e + j.e, i.n + j.n))
            elif kind == 12: it.append(Mirror(i.e * j.e, i.n * j.n))
            elif kind == 13: it.append(Mirror((i.e ^ j.e) | (i.e & j.e), (i.n ^ j.n) | (i.n & j.n)))
            elif kind == 14:
                s = int(rng.integers(1, 31))
                it.append(Mirror((i.e << U(s)) | (j.e >> U(s)), (i.n << np.uint32(s)) | (j.n >> np.uint32(s))))
            elif kind == 15: it.append(Mirror(ek.select(i.e < j.e, i.e, j.e), np.where(i.n < j.n, i.n, j.n).astype(np.uint32)))
            elif kind == 16:
                # float -> uint32 of a bounded value, and back (exercises the conversion paths inside a fused program)
                v = ek.min_(abs(a.e), F(1.0e6)); vn = np.minimum(np.abs(a.n), np.float32(1.0e6))
                # (x86 min: NaN in the first operand yields the second)
                vn = np.where(np.isnan(np.abs(a.n)), np.float32(1.0e6), vn).astype(np.float32)
                it.append(Mirror(U(v), vn.astype(np.uint32)))
                fl.append(Mirror(F(i.e >> U(8)), (i.n >> np.uint32(8)).astype(np.float32)))
    return fl, it


@pytest.mark.parametrize("seed", range(40))
def test_random_expression_dags(gpu, oracle, P, seed):
    ek = gpu
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 33, 1000, 4097, 70_001, 300_000]))
    floats = [rng.uniform(-4, 4, n).astype(np.float32) for _ in range(3)] + [np.array([rng.uniform(-2, 2)], np.float32)]
    ints = [rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(2)] + [np.array([rng.integers(1, 100)], np.uint32)]
    fl, it = _build(ek, oracle, P, rng, n, int(rng.integers(8, 40)), floats, ints)
    # keep a random subset alive (outputs), drop the rest (temporaries that must not be stored), add reductions
    keep_f = [fl[k] for k in rng.choice(len(fl), size=min(4, len(fl)), replace=False)]
    keep_i = [it[k] for k in rng.choice(len(it), size=min(3, len(it)), replace=False)]
    red_i = ek.hsum(keep_i[0].e); red_f = ek.hsum(keep_f[0].e)
    del fl, it
    for m in keep_f:
        got = m.e.numpy(); want = np.broadcast_to(m.n, got.shape)
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (seed, n, got[~same][:3], want[~same][:3])
    for m in keep_i:
        got = m.e.numpy(); want = np.broadcast_to(m.n, got.shape)
        assert (got == want).all(), (seed, n, got[got != want][:3], want[got != want][:3])
    want_i = np.uint32(np.broadcast_to(keep_i[0].n, (max(len(keep_i[0].n), 1),)).sum(dtype=np.uint64) & 0xffffffff)
    if len(keep_i[0].n) > 1 or n == 1:
        assert red_i.numpy()[0] == want_i
    wf = np.broadcast_to(keep_f[0].n, (len(keep_f[0].n),)).astype(np.float64)
    if np.isfinite(wf).all():
        s = wf.sum(); scale = np.abs(wf).sum() + 1e-30
        assert abs(float(red_f.numpy()[0]) - s) <= 2e-6 * scale + 1e-6 * abs(s), (seed, n)
