#!/usr/bin/env python
"""Diagnostic for the row-partitioned multi-GPU SpMM: per-phase timing of the forward operator, the
transposed (backward) operator and the autograd round trip, for the fused gather kernel (sweeping
its two knobs: copy CTAs and arrival groups) and for the NCCL all-gather + two-block path.

    torchrun --nproc-per-node 2 tools/dist_probe.py --shape reddit --k 128
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
ap.add_argument("--reduce", default="sum")
ap.add_argument("--balance", default="nnz")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--sort-degree", action="store_true")
ap.add_argument("--copy-ctas", type=int, nargs="+", default=[32, 64, 128])
ap.add_argument("--groups", type=int, nargs="+", default=[1, 2, 3])
ap.add_argument("--gather", nargs="+", default=["tiles", "owners"], help="arrival groups: K tiles and/or column owners")
ap.add_argument("--no-nccl", action="store_true")
ap.add_argument("--no-bwd", action="store_true")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ["NCCL_DEBUG"] = "WARN"
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from isplib_b200 import capi, synth  # noqa: E402
from isplib_b200.dist import DistSpMM, RowPartitionedSpMM  # noqa: E402

g = synth.make_graph(a.shape, values="uniform", seed=0, device=dev)
rowptr, col, value = g.rowptr, g.col, g.value
if a.sort_degree:
    rowptr, col, value = synth.relabel_by_degree(rowptr, col, value, g.n)
gen = torch.Generator(device=dev).manual_seed(0)


def timed(fn, steps=None):
    steps = steps or a.steps
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    v = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return round(float(v.item()), 4)


res = {"shape": a.shape, "K": a.k, "reduce": a.reduce, "balance": a.balance, "n_gpus": world,
       "sorted_by_degree": bool(a.sort_degree), "fused": [], "nccl": None}

# ---- fused gather + SpMM: one kernel per forward --------------------------------------------------
combos = []
for mode in a.gather:
    for groups in (a.groups if mode == "owners" else [0]):
        combos += [(mode, groups, c) for c in a.copy_ctas]
for mode, groups, ctas in combos:
    if True:
        os.environ["ISPLIB_B200_DIST_GATHER"] = "owners" if mode == "owners" else "auto"
        os.environ["ISPLIB_B200_DIST_GROUPS"] = str(max(groups, 1))
        os.environ["ISPLIB_B200_DIST_COPY_CTAS"] = str(ctas)
        op = RowPartitionedSpMM(rowptr, col, value, g.n, device=dev, balance=a.balance, mode="fused" if world > 1 else None)
        x = torch.randn(op.Rc, a.k, device=dev, generator=gen)
        row = {"gather": mode, "groups": groups, "copy_ctas": ctas}
        row["forward_ms"] = timed(lambda: op.forward(x, a.reduce))
        if world > 1:
            # the multiply alone: same CSR, same grouped plan, X already gathered (no pulls, no waits)
            pb = op.peer_buffers(a.k)
            xg = pb.bufs[op._epoch[a.k] & 1][:, :a.k]
            full = op.full
            plan, variant, row["arrival_groups"] = op.fused_plan_and_variant(a.k, a.reduce)
            out = torch.empty((op.R, a.k), device=dev)
            arg = torch.empty((op.R, a.k), dtype=torch.int64, device=dev) if a.reduce in ("max", "min") else None
            row["multiply_only_ms"] = timed(lambda: capi.spmm_csr(a.reduce, full.rowptr, full.col, full.val, xg, plan, variant,
                                                                  out=out, arg_out=arg, edge_ids=full.edge_ids,
                                                                  arg_sentinel=op.nnz))
            row["items"] = int(plan.info.num_items)
            op.check_status()
        res["fused"].append(row)
        del op
        torch.cuda.empty_cache()
        if world == 1:
            break

# ---- reference: NCCL all-gather on a side stream + local / remote block kernels --------------------
if world > 1 and not a.no_nccl:
    op = RowPartitionedSpMM(rowptr, col, value, g.n, device=dev, balance=a.balance, mode="nccl")
    x = torch.randn(op.Rc, a.k, device=dev, generator=gen)
    r = {"forward_ms": timed(lambda: op.forward(x, a.reduce))}
    r["allgather_only_ms"] = timed(lambda: op._all_gather(x, persistent=True))
    out = torch.empty((op.R, a.k), device=dev)
    gathered = op._all_gather(x)
    code = capi.REDUCE_CODE[a.reduce]
    inner = 0 if code == 3 else code
    arg = torch.empty((op.R, a.k), dtype=torch.int64, device=dev) if inner in (1, 2) else None
    r["local_block_ms"] = timed(lambda: op.block_spmm(inner, op.local, x, out, arg, 0, None, op.nnz, op.variant))
    r["remote_block_ms"] = timed(lambda: op.block_spmm(inner, op.remote, gathered, out, arg, 1, None, op.nnz, op.variant))
    res["nccl"] = r
    del op, gathered
    torch.cuda.empty_cache()

# ---- autograd round trip (forward + transposed operator) with the defaults --------------------------
if not a.no_bwd and a.reduce in ("sum", "mean"):
    for k in ("ISPLIB_B200_DIST_GROUPS", "ISPLIB_B200_DIST_COPY_CTAS", "ISPLIB_B200_DIST_GATHER"):
        os.environ.pop(k, None)
    dop = DistSpMM(rowptr, col, value, g.n, device=dev, balance=a.balance)
    f = dop.fwd
    t = dop.bwd_op(a.reduce == "mean")
    xr = torch.randn(f.Rc, a.k, device=dev, generator=gen).requires_grad_(True)
    gof = torch.randn(f.R, a.k, device=dev, generator=gen)
    go_t = torch.randn(t.Rc, a.k, device=dev, generator=gen)

    def fb():
        xr.grad = None
        dop(xr, a.reduce).backward(gof)

    res["default_mode"] = f.mode
    res["transposed_forward_ms"] = timed(lambda: t.forward(go_t, "sum"))
    res["autograd_fwd_bwd_ms"] = timed(fb)
    # the intermittent slow backward of round 1: look at the spread, not only the mean
    ts = []
    for _ in range(30):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fb()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    v = torch.tensor([max(ts), sorted(ts)[len(ts) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    res["autograd_fwd_bwd_single_step_ms"] = {"max_of_30": round(float(v[0]), 3), "median": round(float(v[1]), 3)}
    f.check_status()
    t.check_status()

if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
