#!/usr/bin/env python
"""How fast do the copy CTAs of the fused gather kernel push a slice of X to the peers?  Times push-only
launches (phase 1 of isplib_b200_spmm_csr_gather: no multiply) for a sweep of copy CTA counts and the
NCCL all-gather of the same slices next to it.

    torchrun --nproc-per-node 2 tools/push_probe.py --rows 1200000 --k 256
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_200_000, help="rows of X per rank")
ap.add_argument("--k", type=int, default=256)
ap.add_argument("--copy-ctas", type=int, nargs="+", default=[16, 32, 64, 128])
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
world = int(os.environ["WORLD_SIZE"])
rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
from isplib_b200.dist import RowPartitionedSpMM  # noqa: E402

# a tiny graph (one entry per row) over world * rows nodes: the multiply is nothing, the slices are real
n = world * a.rows
rowptr = torch.arange(n + 1, device=dev, dtype=torch.int64)
col = torch.arange(n, device=dev, dtype=torch.int64)
res = {"world": world, "rows_per_rank": a.rows, "K": a.k, "slice_MB": round(a.rows * a.k * 4 / 1e6, 1), "push": []}


def timed(fn):
    for _ in range(2):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for ctas in a.copy_ctas:
    os.environ["ISPLIB_B200_DIST_COPY_CTAS"] = str(ctas)
    op = RowPartitionedSpMM(rowptr, col, None, n, device=dev, balance="rows", mode="fused")
    x = torch.randn(op.Rc, a.k, device=dev)
    out = torch.empty(op.R, a.k, device=dev)
    xin = op.next_input_slice(a.k)
    ms = timed(lambda: op._forward_fused(op.next_input_slice(a.k), 0, out, None, phase=1))
    egress = (world - 1) * op.Rc * a.k * 4
    res["push"].append({"copy_ctas": ctas, "ms": round(ms, 4), "egress_GBs_per_gpu": round(egress / ms / 1e6, 1)})
    op.check_status()
    del op
    torch.cuda.empty_cache()

xs = torch.randn(a.rows, a.k, device=dev)
gathered = torch.empty(world * a.rows, a.k, device=dev)
ms = timed(lambda: dist.all_gather_into_tensor(gathered, xs))
res["nccl_all_gather"] = {"ms": round(ms, 4), "ingress_GBs_per_gpu": round((world - 1) * a.rows * a.k * 4 / ms / 1e6, 1)}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
