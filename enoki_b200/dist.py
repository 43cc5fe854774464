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


_native = {"ok": False}


def init_native(rank, world, device=None):
    """Create the backend's OWN NCCL communicator (C ABI: ek_dist_unique_id / ek_dist_init, csrc/ek_dist.cpp) so that the
    per-step all-reduce is one ncclAllReduce enqueued by the library on its stream -- no torch call in the step loop.
    torch.distributed (already initialised by the launcher) only carries the 128-byte id from rank 0 to the other ranks,
    once.  The communicator is checked with one all-reduce of the rank numbers; on any failure this returns False and
    callers keep using allreduce_device_scalar() (torch.distributed) -- the collective path of round 1."""
    import ctypes
    import sys
    import torch
    import torch.distributed as dist
    from . import lib
    L = lib()
    _native["ok"] = False
    if world == 1:
        _native["ok"] = L.ek_dist_init(0, 1, None) == 0
        return _native["ok"]
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    on_gpu = dist.get_backend() == "nccl"

    def all_agree(ok):
        """every rank must take the same path: minimum of the per-rank flags (a collective every rank reaches)"""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
        if on_gpu:
            flag = flag.to(dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    # 1. can every rank load NCCL?  (ncclCommInitRank is collective: a rank that skips it would hang the others)
    buf = (ctypes.c_uint8 * 128)()
    ok = L.ek_dist_unique_id(buf) == 0
    if not ok:
        print(f"enoki_b200.dist: rank {rank}: {L.ek_last_error().decode()}", file=sys.stderr)
    if not all_agree(ok):
        return False
    # 2. rank 0's id to everybody
    ident = torch.tensor(list(buf), dtype=torch.uint8)
    if on_gpu:
        ident = ident.to(dev)
    dist.broadcast(ident, src=0)
    raw = bytes(ident.cpu().tolist())
    # 3. the communicator, then one all-reduce of the rank numbers as a self-check
    ok = L.ek_dist_init(rank, world, ctypes.c_char_p(raw)) == 0
    if not ok:
        print(f"enoki_b200.dist: rank {rank}: {L.ek_last_error().decode()}", file=sys.stderr)
    if not all_agree(ok):
        return False
    from . import Float32
    probe = Float32.copy(np.array([float(rank + 1)], np.float32))
    h = (ctypes.c_uint32 * 1)(probe.index)
    ok = L.ek_allreduce_scalars(h, 1) == 0 and abs(float(probe.numpy()[0]) - world * (world + 1) / 2) < 1e-6
    _native["ok"] = all_agree(ok)
    return _native["ok"]


def native_ready():
    return _native["ok"]


def allreduce_handles(arrays):
    """Sum the given size-1 (or small) backend arrays over all ranks through the library's own communicator."""
    import ctypes
    from . import lib
    h = (ctypes.c_uint32 * len(arrays))(*[a.index for a in arrays])
    if lib().ek_allreduce_scalars(h, len(arrays)) != 0:
        raise RuntimeError(lib().ek_last_error().decode())
