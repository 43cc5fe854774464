"""CPU: the planner + assembler executed by a numpy interpreter of the sweep ISA (tests/ek_emulator.py) on random
expression DAGs -- the same generator as the GPU fuzz test, no GPU needed."""
import numpy as np
import pytest

from ek_emulator import Emulator, Unsupported
import test_gpu_fuzz as fuzz


class _Factory:
    """Stands in for Float32.copy / UInt32.copy: a fake device mapping plus the data the emulator reads."""
    def __init__(self, cls, table, base):
        self.cls, self.table, self.base = cls, table, base

    def copy(self, a):
        r = self.cls.map(self.base + 0x1000000 * (len(self.table) + 1), len(a))
        self.table[r.index] = np.ascontiguousarray(a)
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
