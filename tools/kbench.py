#!/usr/bin/env python
"""Kernel iteration tool (GPU box): times every supported variant of the forward SpMM
through the C ABI on one synthetic shape and prints ms / effective GB/s per variant.

    python tools/kbench.py --shape reddit --k 128 --reduce sum [--novalue] [--iters 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from isplib_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--k", type=int, nargs="+", default=[128])
    ap.add_argument("--reduce", nargs="+", default=["sum"])
    ap.add_argument("--novalue", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--seg-len", type=int, nargs="+", default=[0])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--bwd", action="store_true", help="also time the backward (A^T SpMM / arg scatter)")
    ap.add_argument("--all", action="store_true", help="time every variant, not only the default candidate set")
    ap.add_argument("--pad8", action="store_true", help="rows padded to a multiple of 8 floats (32-byte aligned), as the "
                                                       "op layer / pad_features() lay out widths like 47 or 100")
    a = ap.parse_args()
    if a.all:
        os.environ["ISPLIB_B200_TUNE_ALL"] = "1"
        os.environ["ISPLIB_B200_TUNE_BULK"] = "1"
    dev = "cuda:0"
    g = synth.make_graph(a.shape, values=None if a.novalue else "uniform", seed=0, device=dev, scale=a.scale)
    rp, co = capi.narrow_i64_to_i32(g.rowptr), capi.narrow_i64_to_i32(g.col)
    names = capi.variant_names()
    results = []
    for seg in a.seg_len:
        plan = capi.Plan(rp, g.nnz, seg)
        i = plan.info
        print(f"# {a.shape} m={g.m} nnz={g.nnz} maxdeg={g.max_degree} seg_len={i.seg_len} items={i.num_items} "
              f"split_rows={i.num_split_rows} split_items={i.num_split_items}")
        for k in a.k:
            kp = (k + 7) // 8 * 8 if a.pad8 else (k + 3) // 4 * 4      # rows padded like the op layer does for odd K
            x = torch.randn(g.n, kp, device=dev)[:, :k]
            for red in a.reduce:
                best, times = capi.spmm_autotune(red, rp, co, g.value, x, plan, iters=a.iters)
                b = synth.algorithmic_bytes(g.m, g.nnz, k, g.value is not None, red)
                for v, t in enumerate(times):
                    if t >= 0:
                        mark = " <== best" if v == best else ""
                        print(f"K={k:4d} {red:4s} seg={i.seg_len:4d} {names[v]:24s} {t:8.3f} ms  {b / t / 1e6:9.1f} GB/s{mark}")
                        results.append(dict(k=k, reduce=red, seg_len=i.seg_len, variant=names[v], ms=t, gbs=b / t / 1e6))
    if a.bwd:
        def ev_time(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_tr = ev_time(lambda: capi.csr_transpose(rp, co, g.n), 2)
        colptr, row_t, csr2csc = capi.csr_transpose(rp, co, g.n)
        plan_t = capi.Plan(colptr, g.nnz)
        vt = capi.permute_values(g.value, csr2csc, row_t, rp, False) if g.value is not None else None
        print(f"# CSC view build (one-time): {t_tr:.2f} ms")
        plan = capi.Plan(rp, g.nnz)
        for k in a.k:
            go = torch.randn(g.m, k, device=dev)
            x = torch.randn(g.n, (k + 7) // 8 * 8, device=dev)[:, :k] if a.pad8 else torch.randn(g.n, k, device=dev)
            best, times = capi.spmm_autotune("sum", colptr, row_t, vt, go, plan_t, iters=a.iters)
            b = synth.algorithmic_bytes(g.n, g.nnz, k, vt is not None, "sum")
            print(f"K={k:4d} bwd(sum/mean) = A^T SpMM  {names[best]:24s} {times[best]:8.3f} ms  {b / times[best] / 1e6:9.1f} GB/s")
            results.append(dict(k=k, reduce="bwd_sum", variant=names[best], ms=times[best], gbs=b / times[best] / 1e6))
            t = ev_time(lambda: capi.sddmm_csr(rp, co, go, x, plan, False))
            b = 4 * (g.m + 1) + 4 * g.nnz + 4 * k * g.nnz + 4 * k * g.m + 4 * g.nnz
            print(f"K={k:4d} bwd(sum) grad_value = SDDMM      {'sddmm_lean256 / sddmm_seg':24s} {t:8.3f} ms  {b / t / 1e6:9.1f} GB/s")
            results.append(dict(k=k, reduce="sddmm", variant="sddmm_lean256 / sddmm_seg", ms=t, gbs=b / t / 1e6))
            for red in ("max",):
                _, arg = capi.spmm_csr(red, rp, co, g.value, x, plan)
                t = ev_time(lambda: capi.spmm_arg_backward(co, g.value, None, arg, go, g.n, True, False))
                hv = g.value is not None
                b = 8 * k * g.m + 4 * k * g.m + 4 * k * g.m + (4 * k * g.m if hv else 0) + 8 * k * g.m + 4 * k * g.n
                print(f"K={k:4d} bwd({red}) arg scatter            {'':24s} {t:8.3f} ms  {b / t / 1e6:9.1f} GB/s")
                results.append(dict(k=k, reduce="bwd_" + red, variant="arg_backward", ms=t, gbs=b / t / 1e6))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
