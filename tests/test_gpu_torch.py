"""GPU: PyTorch interop after the reference's tests/python/test_pytorch.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_array_round_trip(gpu):
    torch = pytest.importorskip("torch")
    ek = gpu
    import enoki_b200.torch_interop  # noqa: F401  (installs .torch() / .from_torch())
    rng = np.random.default_rng(3)
    for cls, arr in ((ek.Float32, rng.normal(size=10_001).astype(np.float32)),
                     (ek.Int32, rng.integers(-1000, 1000, 777).astype(np.int32)),
                     (ek.Float64, rng.normal(size=513))):
        t = cls.copy(arr).torch()
        assert isinstance(t, torch.Tensor) and t.is_cuda
        assert (t.cpu().numpy() == arr).all()
        t2 = t * 2 + 1                                  # torch work on torch's stream ...
        back = cls.from_torch(t2)                       # ... is ordered before the copy into the backend
        assert (back.numpy() == (arr * 2 + 1).astype(arr.dtype)).all()
    # the tensor is a copy: modifying it leaves the Enoki array alone (test_pytorch.py:49-57)
    a = ek.Float32.full(42.0, 10)
    t = a.torch(); t += 8
    assert np.allclose(t.cpu().numpy(), 50) and np.allclose(a.numpy(), 42)


def test_autograd_function(gpu):
    """The documented torch.autograd.Function pattern (test_pytorch.py:6-29, 60-71) with f = x * sin(y) + exp(x)."""
    torch = pytest.importorskip("torch")
    ek = gpu
    import enoki_b200.torch_interop  # noqa: F401
    from enoki_b200 import autodiff as ad

    class EnokiFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, y):
            ctx.in1 = ad.FloatD.from_torch(x); ctx.in2 = ad.FloatD.from_torch(y)
            ad.set_requires_gradient(ctx.in1, x.requires_grad)
            ad.set_requires_gradient(ctx.in2, y.requires_grad)
            ctx.out = ctx.in1 * ad.sin(ctx.in2) + ad.exp(ctx.in1)
            out = ctx.out.torch()
            ek.lib().ek_malloc_trim()
            return out

        @staticmethod
        def backward(ctx, grad_out):
            ad.set_gradient(ctx.out, ek.Float32.from_torch(grad_out))
            ad.backward_static()
            res = (ad.gradient(ctx.in1).torch() if ctx.in1.requires_gradient() else None,
                   ad.gradient(ctx.in2).torch() if ctx.in2.requires_gradient() else None)
            del ctx.out, ctx.in1, ctx.in2
            ek.lib().ek_malloc_trim()
            return res

    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(4096, device="cuda", generator=g, requires_grad=True)
    y = torch.randn(4096, device="cuda", generator=g, requires_grad=True)
    w = torch.randn(4096, device="cuda", generator=g)
    (EnokiFn.apply(x, y) * w).sum().backward()
    gx, gy = x.grad.clone(), y.grad.clone()
    x.grad = None; y.grad = None
    ((x * torch.sin(y) + torch.exp(x)) * w).sum().backward()
    assert torch.allclose(gx, x.grad, rtol=2e-5, atol=1e-6)
    assert torch.allclose(gy, y.grad, rtol=2e-5, atol=1e-6)
