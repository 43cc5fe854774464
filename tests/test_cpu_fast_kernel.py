"""CPU: the 32-bit fast sweep kernel ITSELF (enoki_b200/csrc/ek_sweep_fast.cu compiled as host code, tests/cpu_kernel)
executes the lowered programs of the emulator test cases -- random expression DAGs, histograms (privatised bins), group
splitting, reductions, overlapping scatters -- and has to reproduce the expected values.  The numpy interpreter
(ek_emulator.py) still runs the sweeps the fast kernel does not take (size-1 groups, 64-bit programs)."""
import numpy as np
import pytest

import test_cpu_emulator as base
from fast_kernel_emu import FastKernelEmulator


@pytest.fixture(autouse=True)
def native_kernel(monkeypatch):
    monkeypatch.setattr(base, "Emulator", FastKernelEmulator)
    yield


@pytest.mark.parametrize("seed", range(0, 150, 5))
def test_random_dags_on_the_host_compiled_kernel(ek, oracle, P, seed):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    # (sizes 1 and 33 give size-1 / tiny groups: those seeds would not reach the kernel -> pick seeds whose n is wide)
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([1, 33, 1000, 4097]))
    if n < 1000:
        pytest.skip("size-1 / 33-element case: not a fast-kernel sweep")
    before = FastKernelEmulator.native_sweeps
    skip = base._case(ek, oracle, P, seed)
    gc.collect()
    if skip:
        pytest.skip(skip)
    if FastKernelEmulator.native_sweeps == before:
        pytest.skip("this DAG uses an operation outside the fast kernel's set: every sweep took the general form")


@pytest.mark.parametrize("n_bins,n", [(31, 5000), (31, 40_003), (300, 9000), (2000, 9000)])
def test_histogram_on_the_host_compiled_kernel(ek, oracle, P, n_bins, n):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    base._case_histogram(ek, oracle, P, n_bins, n)
    gc.collect()


def test_group_splitting_on_the_host_compiled_kernel(ek, oracle, P):
    import gc
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    assert base._case_many_outputs(ek, oracle, P) >= 2
    gc.collect()


def test_side_effect_order_on_the_host_compiled_kernel(ek, oracle, P):
    base.test_side_effects_run_in_recording_order(ek, oracle, P)


def _differential(ek, oracle, table, outputs, float_tol=()):
    """Runs the pending trace on the numpy interpreter (assembler's program form) and on the host-compiled fast kernel
    (lowered form) and compares every output bit for bit (`float_tol`: names compared with 1e-5 relative instead)."""
    from ek_emulator import Emulator
    prog = ek.debug_program()
    ref = Emulator(oracle, {k: np.array(v, copy=True) for k, v in table.items()}, base._Factory.addresses)
    ref.run(prog)
    nat = FastKernelEmulator(oracle, {k: np.array(v, copy=True) for k, v in table.items()}, base._Factory.addresses)
    before = FastKernelEmulator.native_sweeps
    nat.run(prog)
    assert FastKernelEmulator.native_sweeps > before
    for name, var in outputs.items():
        a, b = ref.vars[var.index], nat.vars[var.index]
        assert a.shape == b.shape and a.dtype == b.dtype, (name, a.shape, b.shape, a.dtype, b.dtype)
        if name in float_tol:
            scale = np.abs(a[np.isfinite(a)].astype(np.float64)).sum() if a.size == 1 or name.startswith("sum") else max(float(np.abs(a[np.isfinite(a)]).max()), 1e-30)
            assert np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=0, atol=1e-5 * max(scale, 1e-30), equal_nan=True), name
        else:
            av, bv = a.view(np.uint8), b.view(np.uint8)
            assert (av == bv).all(), (name, a[:8], b[:8], int((av != bv).sum()))
    ek.lib().ek_debug_discard_side_effects()


@pytest.mark.parametrize("n", [4097, 20_011])
def test_directed_programs_numpy_vs_host_compiled_kernel(ek, oracle, P, n):
    """The programs of the kernel-qualification battery (csrc/ek_qualify.cpp `directed`): mask outputs (8-bit stores)
    reused as staged 8-bit inputs, global gather / scatter / scatter_add with out-of-range indices masked off, signed
    integer operations, rounding conversions, literal operands in every position, min / max / product / count
    reductions, the guarded multiply-adds of the tape."""
    import gc
    from enoki_b200 import Float32, UInt32, Int32, Mask, gather, scatter, scatter_add, fmadd, select, hsum, hmin, hmax, hprod
    import enoki_b200 as E
    gc.collect(); ek.lib().ek_debug_discard_side_effects(); gc.collect()
    rng = np.random.default_rng(n)
    table = {}
    F = base._Factory(Float32, table, 0x7f0000000000); U = base._Factory(UInt32, table, 0x7a0000000000)
    xs_n = [rng.uniform(-4, 4, n).astype(np.float32) for _ in range(4)]
    xs_n[0][:5] = [0.0, -0.0, np.inf, np.nan, 1e-41]
    x0, x1, x2, x3 = (F.copy(a) for a in xs_n)

    # -- masks as outputs, then as staged inputs of a second evaluation
    m1 = x0 > x1; m2 = x2 <= 0.5
    m = (m1 & m2) | ~m1
    _differential(ek, oracle, table, {"mask": m})
    # (give the mask storage so that the next trace stages it as bytes)
    m_n = ((xs_n[0] > xs_n[1]) & (xs_n[2] <= 0.5)) | ~(xs_n[0] > xs_n[1])
    M = base._Factory(Mask, table, 0x7b0000000000).copy(m_n)
    sel = select(M, x0, x1 * 2.0)
    cnt = E._reduce("COUNT", M, UInt32)
    _differential(ek, oracle, table, {"select": sel, "count": cnt})
    del sel, cnt, m, m1, m2

    # -- global gather / scatter_add (6000-entry targets: beyond the shared-memory paths), masked out-of-range indices
    mm = 6000
    src_n = rng.uniform(-1, 1, mm).astype(np.float32)
    idx_n = rng.integers(0, mm + 50, n).astype(np.uint32)
    S = F.copy(src_n); IDX = U.copy(idx_n); TU = U.copy(np.zeros(mm, np.uint32)); TF = F.copy(np.zeros(mm, np.float32))
    ok = IDX < UInt32(mm)
    g = gather(Float32, S, IDX, ok) * x1
    scatter_add(TU, IDX & UInt32(7), IDX, ok)
    scatter_add(TF, x2, IDX, ok)
    _differential(ek, oracle, table, {"gather": g, "add_u32": TU, "add_f32": TF}, float_tol=("add_f32",))
    del g, ok

    # -- index arithmetic with literals, signed operations, conversions
    i = UInt32.arange(n)
    h = i * np.uint32(2654435761) + np.uint32(974711)
    h = (h ^ (h >> 15)) * np.uint32(2246822519)
    s = Int32(x0 * 1000.0)
    s2 = E.max_(s >> Int32(3), -abs(s))
    tof = Float32(s) + Float32(h)
    lt = s < Int32(0)
    _differential(ek, oracle, table, {"hash": h, "signed": s2, "tofloat": tof, "lt": lt})
    del h, s, s2, tof, lt, i

    # -- literal operands in every position
    a = fmadd(x0, 1.5, x1); b = fmadd(x0, x1, -0.75); c = 2.0 - x2; d = 3.0 / x3
    e_ = select(x0 >= 0.25, Float32(1.5), x1); f_ = E.min_(abs(x1), 2.5)
    _differential(ek, oracle, table, {"fma_ub": a, "fma_uc": b, "rsub": c, "rdiv": d, "sel_lit": e_, "min_lit": f_})
    del a, b, c, d, e_, f_

    # -- reductions (NaN-free data) and the guarded multiply-adds
    y0 = F.copy(rng.uniform(-4, 4, n).astype(np.float32)); y1 = F.copy(rng.uniform(-4, 4, n).astype(np.float32))
    r_min = hmin(y0 * y1); r_max = hmax(E.sin(y1)); r_sum = hsum(y0 * y0)
    _differential(ek, oracle, table, {"hmin": r_min, "hmax": r_max, "sum": r_sum}, float_tol=("sum",))
    gc.collect()


def test_block_size_128_instantiation():
    """The same cases once more with the fast kernel restricted to its 128-thread instantiation (EK_FAST_T=128 is read once
    per process, hence the child process): operand offsets, group strides and the reduction epilogue for T = 128."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EK_FAST_T="128")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_cpu_fast_kernel.py"), "-q", "-x",
                        "-k", "not block_size_128", "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=1200, env=env, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
