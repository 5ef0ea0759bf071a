#!/usr/bin/env python
"""Diagnostic: time the forward operator and the transposed (backward) operator of DistSpMM
separately, per phase, on N GPUs.

    torchrun --nproc-per-node 2 tools/dist_probe.py --shape reddit --k 128 --balance nnz
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="reddit")
ap.add_argument("--k", type=int, default=128)
ap.add_argument("--balance", default="rows")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--sort-degree", action="store_true")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ["NCCL_DEBUG"] = "WARN"
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from isplib_b200 import synth  # noqa: E402
from isplib_b200.dist import DistSpMM  # noqa: E402

g = synth.make_graph(a.shape, values="uniform", seed=0, device=dev)
rowptr, col, value = g.rowptr, g.col, g.value
if a.sort_degree:
    rowptr, col, value = synth.relabel_by_degree(rowptr, col, value, g.n)
op = DistSpMM(rowptr, col, value, g.n, device=dev, balance=a.balance)
f = op.fwd
t = op.bwd_op(False)
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(f.Rc, a.k, device=dev, generator=gen)
go = torch.randn(t.Rc, a.k, device=dev, generator=gen)


def timed(fn):
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    v = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return round(float(v.item()), 3)


def block_only(o, blk, xin):
    out = torch.empty((o.R, a.k), device=dev)
    return lambda: o.block_spmm(0, blk, xin, out, None, 0, None, o.nnz, o.variant)


res = {"balance": a.balance, "n_gpus": world, "sorted_by_degree": bool(a.sort_degree), "row_bounds": f.row_bounds}
for name, o, xin in (("fwd", f, x), ("bwd", t, go)):
    res[name + "_full_ms"] = timed(lambda: o.forward(xin, "sum"))
    res[name + "_local_ms"] = timed(block_only(o, o.local, xin))
    if world > 1:
        gathered = o._all_gather(xin)
        res[name + "_remote_ms"] = timed(block_only(o, o.remote, gathered))
        res[name + "_allgather_ms"] = timed(lambda: o._all_gather(xin))
    res[name + "_nnz"] = [o.local.nnz, o.remote.nnz]
    res[name + "_R_Rc"] = [o.R, o.Rc]
xr = x.clone().requires_grad_(True)
gof = torch.randn(f.R, a.k, device=dev, generator=gen)


def fb():
    xr.grad = None
    op(xr, "sum").backward(gof)


res["autograd_fwd_bwd_ms"] = timed(fb)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
