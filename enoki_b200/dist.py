"""Multi-GPU plumbing: one rank per GPU, element-range sharding of every wide array,
and ONE all-reduce of the size-1 results (loss / gradients of size-1 leaves / small scatter targets).

Every vertical op and the tape sweep are element-wise independent, so each rank records the
identical trace on its slice and no data-path collective is needed; the adjoint sweep is linear
in the adjoints, hence reducing the leaf scalars once at the end equals reducing at every hsum.
`torch.distributed` is only the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous element range [lo, hi) of rank `rank` ([r*N/W, (r+1)*N/W))."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def allreduce_scalars(values, op="sum"):
    """All-reduce a small vector of per-rank partial scalars (host numpy or torch tensor).
    Returns a numpy array.  With no process group initialised this is the identity."""
    import torch
    import torch.distributed as dist
    t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
        dist.all_reduce(t, op=ops[op])
    return t.detach().cpu().numpy()


def allreduce_device_scalar(ptr, dtype, stream_ptr, device):
    """In-place NCCL all-reduce of a scalar that lives in the backend's device memory, enqueued on
    the backend's own stream (no host synchronisation)."""
    import torch
    import torch.distributed as dist

    class _Dev:
        def __init__(self, p, typestr):
            self.__cuda_array_interface__ = {"shape": (1,), "typestr": typestr, "data": (p, False), "version": 3}
    typestr = {"f32": "<f4", "f64": "<f8", "u32": "<u4"}[dtype]
    ext = torch.cuda.ExternalStream(stream_ptr, device=device)
    with torch.cuda.stream(ext):
        ten = torch.as_tensor(_Dev(ptr, typestr), device=device)
        dist.all_reduce(ten)
    return ten
