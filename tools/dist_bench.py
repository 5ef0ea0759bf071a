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
ap.add_argument("--mode", default=None, choices=["auto", "fused", "nccl"], help="default: auto (measured per width)")
ap.add_argument("--balance", default="nnz", choices=["nnz", "rows"], help="row ranges balanced by stored entries or by rows")
ap.add_argument("--prebuild", action="store_true", help="build the transposed (backward) operator before timing")
ap.add_argument("--sort-degree", action="store_true",
                help="relabel the nodes by descending degree first (ids sorted by degree: the skewed case)")
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
op = DistSpMM(rowptr, col, value, g.n, device=dev, balance=a.balance, mode=a.mode)
f = op.fwd
if a.prebuild and a.reduce in ("sum", "mean"):
    op.bwd_op(a.reduce == "mean")
nnz_rank = torch.tensor([f.local.nnz + f.remote.nnz], device=dev, dtype=torch.int64)
nnz_all = [torch.zeros_like(nnz_rank) for _ in range(world)]
if world > 1:
    dist.all_gather(nnz_all, nnz_rank)
else:
    nnz_all = [nnz_rank]
nnz_all = [int(t.item()) for t in nnz_all]
gen = torch.Generator(device=dev).manual_seed(0)
c0, c1 = f.col_range()
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
                      "fwd_total_effective_gbs": round(b / t_f / 1e6, 1), "balance": a.balance,
                      "mode": f.mode, "mode_chosen": f.mode_for(a.k, a.reduce) if world > 1 else "single",
                      "mode_times_ms": f._mode_choice.get((a.k, a.reduce in ("max", "min"))),
                      "sorted_by_degree": bool(a.sort_degree), "row_bounds": f.row_bounds,
                      "nnz_per_rank": nnz_all, "nnz_imbalance": round(max(nnz_all) * world / max(1, sum(nnz_all)), 3)}),
          flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
