#!/usr/bin/env python
"""Minimal driver for ncu: builds one synthetic graph and launches one kernel a few times.

    python tools/prof_driver.py --shape reddit --k 128 --reduce max [--variant 7] [--bwd]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from isplib_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="reddit")
ap.add_argument("--k", type=int, default=128)
ap.add_argument("--reduce", default="sum")
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--novalue", action="store_true")
ap.add_argument("--bwd", action="store_true", help="profile the arg-scatter backward (max/min)")
ap.add_argument("--aux", action="store_true", help="with --bwd: the streamed scatter fed by the forward's col[arg] / val[arg] outputs")
ap.add_argument("--binned", action="store_true", help="with --bwd --aux: the partition-then-apply scatter")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = "cuda:0"
g = synth.make_graph(a.shape, values=None if a.novalue else "uniform", seed=0, device=dev)
rp, co = capi.narrow_i64_to_i32(g.rowptr), capi.narrow_i64_to_i32(g.col)
plan = capi.Plan(rp, g.nnz)
x = torch.randn(g.n, a.k, device=dev)
v = a.variant
if v < 0:
    v = capi.lib().isplib_b200_variant_default(capi.REDUCE_CODE[a.reduce], g.n, a.k, a.k, a.k, x.data_ptr(), x.data_ptr(), 0.0)
print("variant", v, capi.variant_names()[v])
out, arg = capi.spmm_csr(a.reduce, rp, co, g.value, x, plan, v)
if a.bwd and a.aux:
    go = torch.randn(g.m, a.k, device=dev)
    acol = torch.empty(g.m, a.k, dtype=torch.int32, device=dev)
    aval = torch.empty(g.m, a.k, device=dev) if g.value is not None else None
    capi.spmm_csr(a.reduce, rp, co, g.value, x, plan, v, out=out, arg_out=arg, arg_col=acol, arg_val=aval)
    for _ in range(a.reps):
        capi.spmm_arg_backward_aux(acol, aval, go, g.n, binned=a.binned)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        capi.spmm_arg_backward_aux(acol, aval, go, g.n, binned=a.binned)
    e1.record()
    torch.cuda.synchronize()
    print(f"arg scatter ({'binned' if a.binned else 'direct'}) {e0.elapsed_time(e1) / a.reps:.3f} ms per call")
elif a.bwd:
    go = torch.randn(g.m, a.k, device=dev)
    for _ in range(a.reps):
        capi.spmm_arg_backward(co, g.value, None, arg, go, g.n, True, False)
else:
    for _ in range(a.reps):
        capi.spmm_csr(a.reduce, rp, co, g.value, x, plan, v, out=out, arg_out=arg)
torch.cuda.synchronize()
print("done")
