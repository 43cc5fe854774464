"""GPU parity tests for the DiffArray tape path (C ABI ek_tape_*) against the oracle / golden vectors."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SZ = ctypes.c_size_t
GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32 = 10


def build_tape(ek, g):
    """Materialise a graph description (make_golden.make_tape format) on the GPU tape; returns node ids."""
    L = ek.lib()
    n_nodes = len(g["node_size"])
    ids = [0] * (n_nodes + 1)
    src, dst = g["src"], g["dst"]
    e = 0
    keep = []
    for i in range(1, n_nodes + 1):
        if e < len(dst) and dst[e] == i:
            ids[i] = L.ek_tape_append_node(F32, int(g["node_size"][i - 1]), b"n")
            while e < len(dst) and dst[e] == i:
                off, ws = int(g["woff"][e]), int(g["wsize"][e])
                w = ek.Float32.copy(g["weights"][off:off + ws])
                keep.append(w)
                assert L.ek_tape_append_edge(F32, ids[int(src[e])], ids[i], w.index) == 0
                e += 1
        else:
            ids[i] = L.ek_tape_append_leaf(F32, int(g["node_size"][i - 1]))
    return ids, keep


def run_backward(ek, g):
    L = ek.lib()
    # the CPU reference tape never simplifies (the pre-eval callback exists for CUDA arrays only,
    # autodiff.cpp:214-221): bit-exact comparisons need the unsimplified graph
    L.ek_tape_set_graph_simplification(F32, 0)
    try:
        return _run_backward(ek, g)
    finally:
        L.ek_tape_set_graph_simplification(F32, 1)


def _run_backward(ek, g):
    L = ek.lib()
    ids, keep = build_tape(ek, g)
    assert L.ek_tape_backward(F32, ids[int(g["root"])], 1) == 0, L.ek_last_error()
    outs = []
    for wnt in g["want"]:
        h = L.ek_tape_gradient(F32, ids[int(wnt)])
        sz = int(g["node_size"][int(wnt) - 1])
        if h == 0:
            outs.append(np.zeros(sz, np.float32)); continue
        L.ek_inc_ref_ext(h)
        a = ek.Float32.from_index(h).numpy()
        outs.append(np.broadcast_to(a, (sz,)).copy() if a.size == 1 else a)
    for i in ids[1:]:
        L.ek_tape_dec_ref_ext(F32, i)
    return np.concatenate(outs)


@pytest.mark.parametrize("tag", ["small", "wide"])
def test_tape_golden_bit_exact(gpu, tag):
    """Layered graph (C4 shape) vs the reference tape's gradients (bit-exact: same per-source
    accumulation order, safe_mul / safe_fmadd)."""
    g = dict(np.load(os.path.join(GOLD, f"tape_{tag}.npz")))
    got = run_backward(gpu, g)
    want = g["grads"]
    assert got.shape == want.shape
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), np.abs(got - want).max()
    assert gpu.lib().ek_tape_node_count(F32) == 0          # free_graph released every node


@pytest.mark.parametrize("L_,K,w", [(3, 4, 1), (5, 16, 2048), (12, 32, 4099), (8, 64, 20000)])
def test_tape_vs_oracle(gpu, oracle, P, L_, K, w):
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import make_tape
    rng = np.random.default_rng(L_ * 1000 + K)
    g = make_tape(rng, L_, K, w)
    got = run_backward(gpu, g)
    want = np.zeros(len(g["want"]) * w, np.float32)
    rc = oracle.or_tape_backward(len(g["node_size"]), P(g["node_size"]), len(g["src"]), P(g["src"]), P(g["dst"]),
                                 P(g["weights"]), P(g["woff"]), P(g["wsize"]), int(g["root"]), len(g["want"]),
                                 P(g["want"]), P(want))
    assert rc == 0
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), np.abs(got - want).max()


def test_tape_linearity(gpu):
    """Size-independent property: the adjoint sweep is linear in the seed -- backward with all
    weights doubled on the last level doubles the leaf gradients exactly (power-of-two scaling)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import make_tape
    g = make_tape(np.random.default_rng(42), 6, 16, 3000, zero_frac=0.0)
    a = run_backward(gpu, g)
    g2 = dict(g); w2 = g["weights"].copy()
    w2[-16:] = 2.0            # the 16 weight-1 edges into the loss node
    g2["weights"] = w2
    b = run_backward(gpu, g2)
    assert (b == 2 * a).all()


def test_ad_expressions_golden(gpu, ulp):
    """tests/autodiff.cpp-style expressions through the public DiffArray mirror; values bit-exact,
    gradients within 4 ulp of the reference tape (weights use rcp, which is the only op that is not
    bit-identical between the CPU and this backend)."""
    ek = gpu
    from enoki_b200 import autodiff as ad
    g = np.load(os.path.join(GOLD, "ad_expr.npz"))
    x = g["x"]
    exprs = [
        lambda v: v * v,
        lambda v: ad.sin(v) * ad.exp(v),
        lambda v: ad.sqrt(ad.abs_(v) + 1.0) / (v * v + 2.0),
        lambda v: ad.log(v * v + 1.0) + ad.cos(v),
        lambda v: ad.fmadd(v, v, v) * ad.rcp(v * v + 1.0),
    ]
    for k, f in enumerate(exprs):
        xd = ad.FloatD(ek.Float32.copy(x))
        ad.set_requires_gradient(xd)
        y = f(xd)
        loss = ad.hsum(y)
        val = y.value.numpy()
        ad.backward(loss)
        grad = ad.gradient(xd).numpy()
        vt = 0 if k not in (2, 4) else 2
        assert ulp(val, g[f"val{k}"]).max() <= vt, (k, ulp(val, g[f"val{k}"]).max())
        assert np.allclose(grad, g[f"grad{k}"], rtol=2e-6, atol=1e-6), (k, np.abs(grad - g[f"grad{k}"]).max())


def test_backward_scalar_leaf_hsum(gpu):
    """size-1 leaf feeding a wide expression: gradient = hsum of the wide adjoint (autodiff.cpp:867-871)."""
    ek = gpu
    from enoki_b200 import autodiff as ad
    x = np.random.default_rng(1).uniform(-1, 1, 10_000).astype(np.float32)
    s = ad.FloatD(ek.Float32(1.5))
    ad.set_requires_gradient(s)
    y = ad.FloatD(ek.Float32.copy(x)) * s          # d/ds sum(x*s) = sum(x)
    loss = ad.hsum(y * y)                          # d/ds sum(x^2 s^2) = 2 s sum(x^2)
    ad.backward(loss)
    got = float(ad.gradient(s).numpy()[0])
    want = 2 * 1.5 * float((x.astype(np.float64) ** 2).sum())
    assert abs(got - want) <= 1e-4 * abs(want)


def test_gradient_descent(gpu):
    """tests/autodiff.cpp:550-562 -- 10 steps of gradient descent on a scalar."""
    ek = gpu
    from enoki_b200 import autodiff as ad
    x = ad.FloatD(ek.Float32(3.0))
    for _ in range(10):
        ad.set_requires_gradient(x, True)
        loss = (x - 1.0) * (x - 1.0)
        ad.backward(loss)
        gx = ad.gradient(x)
        nv = x.value - gx * 0.1
        x.set_requires_gradient(False)
        x = ad.FloatD(nv)
    final = float(x.value.numpy()[0])
    assert abs(final - (1 + 2 * 0.8 ** 10)) < 1e-5


def test_simplify_graph_preserves_gradients(gpu):
    """Greedy vertex elimination (autodiff.cpp:990-1074) collapses interior nodes into zero-guarded weight
    products; gradients must be unchanged (up to reassociation of the products)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import make_tape
    ek = gpu
    L = ek.lib()
    g = make_tape(np.random.default_rng(3), 5, 6, 257, zero_frac=0.02)
    L.ek_tape_set_graph_simplification(F32, 0)
    a = _run_backward(ek, g)
    ids, keep = build_tape(ek, g)
    n_before = L.ek_tape_node_count(F32)
    L.ek_tape_set_graph_simplification(F32, 1)
    for i in ids[1:]:
        if i not in (ids[int(g["root"])],) and i not in [ids[int(w)] for w in g["want"]]:
            L.ek_tape_dec_ref_ext(F32, i)            # interior nodes are only referenced by the graph
    assert L.ek_tape_simplify(F32) == 0
    assert L.ek_tape_node_count(F32) < n_before       # interior nodes with degree product <= 10 are gone
    assert L.ek_tape_backward(F32, ids[int(g["root"])], 1) == 0, L.ek_last_error()
    outs = []
    for wnt in g["want"]:
        h = L.ek_tape_gradient(F32, ids[int(wnt)])
        L.ek_inc_ref_ext(h)
        outs.append(ek.Float32.from_index(h).numpy())
    b = np.concatenate(outs)
    for wnt in g["want"]:
        L.ek_tape_dec_ref_ext(F32, ids[int(wnt)])
    L.ek_tape_dec_ref_ext(F32, ids[int(g["root"])])
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert L.ek_tape_node_count(F32) == 0
