#!/usr/bin/env python
"""Effect of node order on the SpMM (isplib_b200.reorder): a community-structured graph whose
node ids were shuffled (what a raw dataset often looks like) vs the same graph after a
locality-recovering reordering.  Prints ms / effective GB/s for each order.

    python tools/reorder_demo.py [--nodes 500000] [--deg 100] [--community 2000] [--k 128]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from isplib_b200 import capi, io, reorder, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=500_000)
ap.add_argument("--deg", type=int, default=100)
ap.add_argument("--community", type=int, default=2000)
ap.add_argument("--inside", type=float, default=0.9)
ap.add_argument("--k", type=int, default=128)
ap.add_argument("--rcm", action="store_true", help="also run scipy's RCM (host side, slow on big graphs)")
a = ap.parse_args()
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(0)
N, D = a.nodes, a.deg
row = torch.arange(N, device=dev).repeat_interleave(D)
comm = (row // a.community) * a.community
inside = torch.rand(N * D, device=dev, generator=gen) < a.inside
col = torch.where(inside, comm + torch.randint(0, a.community, (N * D,), device=dev, generator=gen),
                  torch.randint(0, N, (N * D,), device=dev, generator=gen)).clamp_(max=N - 1)
val = torch.rand(N * D, device=dev, generator=gen)
x = torch.randn(N, a.k, device=dev, generator=gen)


def measure(name, adj):
    rowptr, c, v = adj.csr()
    rp, co = capi.narrow_i64_to_i32(rowptr), capi.narrow_i64_to_i32(c)
    plan = capi.Plan(rp, co.numel())
    best, times = capi.spmm_autotune("sum", rp, co, v, x, plan, iters=5)
    b = synth.algorithmic_bytes(N, co.numel(), a.k, True)
    print(f"{name:34s} {times[best]:8.3f} ms  {b / times[best] / 1e6:9.1f} GB/s  ({capi.variant_names()[best]})", flush=True)
    return times[best]


original = io.from_edge_index(torch.stack([row, col]), val, N, N, dev)
t0 = measure("community order (ground truth)", original)
shuffle = torch.randperm(N, device=dev, generator=gen)
shuffled = reorder.permute(original, shuffle)
t1 = measure("shuffled node ids", shuffled)
t2 = measure("shuffled -> degree order", reorder.permute(shuffled, reorder.degree_order(shuffled)))
if a.rcm:
    t = time.time()
    perm = reorder.reverse_cuthill_mckee(shuffled)
    print(f"  (RCM on the host: {time.time() - t:.1f} s)")
    t3 = measure("shuffled -> reverse Cuthill-McKee", reorder.permute(shuffled, perm))
    print(f"speed-up of RCM over shuffled: {t1 / t3:.2f}x; ground-truth order: {t1 / t0:.2f}x")
