"""Worker of tests/test_gpu_zz_round2.py::test_two_rank_sharded_matches_single_gpu (launched by torch.distributed.run, one
rank per GPU).  SURVEY 8e: every rank records the identical trace on its element range; the only cross-GPU values are the
size-1 results -- here the loss and the gradients of two scalar leaves -- summed by ONE all-reduce.  Rank 0 then evaluates
the whole range alone and compares (reassociation bound: the sums are folded in a different order)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem(ek, ad, lo, hi):
    """loss = sum over [lo, hi) of (sin(a x + b) x)^2 ;  x_i = (i mod 4093) / 4093 - 0.5 ;  a, b differentiable scalars"""
    from enoki_b200 import Float32, UInt32
    i = UInt32.arange(hi - lo) + UInt32(lo)
    x = Float32(i % UInt32(4093)) * Float32(1.0 / 4093.0) - 0.5
    a = ad.FloatD(Float32(1.5)); b = ad.FloatD(Float32(-0.25))
    ad.set_requires_gradient(a); ad.set_requires_gradient(b)
    xd = ad.FloatD(x)
    y = ad.sin(a * xd + b) * xd
    loss = ad.hsum(y * y)
    ad.backward(loss)
    return loss.value, ad.gradient(a), ad.gradient(b)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import enoki_b200 as ek
    from enoki_b200 import autodiff as ad
    from enoki_b200.dist import shard_range, allreduce_device_scalar, init_native, allreduce_handles
    L = ek.lib()
    L.ek_set_device(local)
    assert L.ek_init() == 0
    n = (1 << 22) + 12345
    lo, hi = shard_range(n, rank, world)
    out = {}
    # (a) the transport measured in round 1: torch.distributed on the backend's stream
    vals = problem(ek, ad, lo, hi)
    for v in vals:
        v.eval()
        allreduce_device_scalar(L.ek_var_ptr(v.index), "f32", L.ek_stream(), torch.device("cuda", local))
    ek.cuda_sync(); torch.cuda.synchronize()
    out["torch"] = [float(v.numpy()[0]) for v in vals]
    # (b) the library's own communicator behind the C ABI (csrc/ek_dist.cpp)
    if init_native(rank, world, torch.device("cuda", local)):
        vals = problem(ek, ad, lo, hi)
        allreduce_handles(list(vals))
        ek.cuda_sync()
        out["native"] = [float(v.numpy()[0]) for v in vals]
    else:
        out["native"] = None
    dist.barrier()
    if rank == 0:
        ref = problem(ek, ad, 0, n)
        out["single"] = [float(v.numpy()[0]) for v in ref]
        print("TWO_RANK_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
