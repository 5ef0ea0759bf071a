#!/usr/bin/env python
"""Row-partitioned SpMM forward + backward on N GPUs (BASELINE.json config 5: SpMM-max with
argmax backward on the AmazonProducts-shaped graph, K=200, at 2/4/8 B200).

    torchrun --nproc-per-node N tools/dist_bench.py --shape amazon --k 200 --reduce max
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="amazon")
ap.add_argument("--k", type=int, default=200)
ap.add_argument("--reduce", default="max")
ap.add_argument("--steps", type=int, default=10)
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
op = DistSpMM(g.rowptr, g.col, g.value, g.n, device=dev)
f = op.fwd
gen = torch.Generator(device=dev).manual_seed(0)
c0, c1 = rank * f.Rc, min((rank + 1) * f.Rc, g.n)
x = f.pad_x(torch.randn(g.n, a.k, device=dev, generator=gen)[c0:c1]).requires_grad_(True)
go = torch.randn(f.R, a.k, device=dev, generator=gen)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, n):
    for _ in range(3):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def fwd():
    with torch.no_grad():
        op(x, a.reduce)


def fwd_bwd():
    x.grad = None
    op(x, a.reduce).backward(go)


t_f = timed(fwd, a.steps)
t_fb = timed(fwd_bwd, a.steps)
if rank == 0:
    b = synth.algorithmic_bytes(g.m, g.nnz, a.k, True, a.reduce)
    print(json.dumps({"shape": a.shape, "nodes": g.m, "nnz": g.nnz, "K": a.k, "reduce": a.reduce, "n_gpus": world,
                      "fwd_ms": round(t_f, 3), "fwd_bwd_ms": round(t_fb, 3), "bwd_ms": round(t_fb - t_f, 3),
                      "fwd_total_effective_gbs": round(b / t_f / 1e6, 1)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
