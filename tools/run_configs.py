#!/usr/bin/env python
"""Measures every BASELINE.json config on one B200 and writes a markdown table
(profiles/rNN_configs.md): forward (best variant, on-device autotune) and backward of each
reduction, as ms / GFLOP/s / effective GB/s (B_alg) / fraction of the measured HBM peak.

    python tools/run_configs.py --out gpurun_out/configs.md [--quick]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from isplib_b200 import capi, synth  # noqa: E402

DEV = "cuda:0"


def ev_time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run_shape(shape, ks, reduces, values, rows, iters):
    g = synth.make_graph(shape, values=values, seed=0, device=DEV)
    rp, co = capi.narrow_i64_to_i32(g.rowptr), capi.narrow_i64_to_i32(g.col)
    plan = capi.Plan(rp, g.nnz)
    colptr, row_t, csr2csc = capi.csr_transpose(rp, co, g.n)
    plan_t = capi.Plan(colptr, g.nnz)
    names = capi.variant_names()
    hv = g.value is not None
    pk = peak()
    for k in ks:
        kp = (k + 7) // 8 * 8                         # the op layer's / pad_features() layout: rows padded to 32 bytes
        x = torch.randn(g.n, kp, device=DEV)[:, :k]
        go = torch.randn(g.m, kp, device=DEV)[:, :k]
        for red in reduces:
            best, times = capi.spmm_autotune(red, rp, co, g.value, x, plan, iters=iters)
            t = times[best]
            b = synth.algorithmic_bytes(g.m, g.nnz, k, hv, red)
            rows.append((shape, g.m, g.nnz, k, red, "fwd", names[best], t, 2 * g.nnz * k / t / 1e6, b / t / 1e6, b / t / 1e6 / pk))
            if red in ("sum", "mean"):
                w = capi.permute_values(g.value, csr2csc, row_t, rp, red == "mean") if (hv or red == "mean") else None
                bb, tt = capi.spmm_autotune("sum", colptr, row_t, w, go, plan_t, iters=iters)
                t = tt[bb]
                b = synth.algorithmic_bytes(g.n, g.nnz, k, w is not None, "sum")
                rows.append((shape, g.m, g.nnz, k, red, "bwd (A^T SpMM)", names[bb], t, 2 * g.nnz * k / t / 1e6, b / t / 1e6, b / t / 1e6 / pk))
                if hv and red == "sum":
                    t = ev_time(lambda: capi.sddmm_csr(rp, co, go, x, plan, False), iters)
                    b = 4 * (g.m + 1) + 4 * g.nnz + 4 * k * g.nnz + 4 * k * g.m + 4 * g.nnz
                    rows.append((shape, g.m, g.nnz, k, red, "bwd grad_value (SDDMM)", "sddmm_lean256 / sddmm_seg", t, 2 * g.nnz * k / t / 1e6, b / t / 1e6, b / t / 1e6 / pk))
            else:
                # the op layer's path: the forward also writes col[arg] / val[arg]; the backward streams them
                # (direct REDs while grad_x can be L2-resident, partition-then-apply beyond 256 MB)
                acol = torch.empty(g.m, k, dtype=torch.int32, device=DEV)
                aval = torch.empty(g.m, k, device=DEV) if hv else None
                capi.spmm_csr(red, rp, co, g.value, x, plan, best, arg_col=acol, arg_val=aval)
                binned = g.n * k * 4 > 256 * 2**20
                goc = go.contiguous()
                t = ev_time(lambda: capi.spmm_arg_backward_aux(acol, aval, goc, g.n, binned=binned), iters)
                b = (4 + 4 + (4 if hv else 0) + 8) * k * g.m + 4 * k * g.n      # SURVEY 8d with the 4-byte aux stream instead of the 8-byte arg + col gather
                rows.append((shape, g.m, g.nnz, k, red, "bwd (arg scatter)", "arg_backward_binned" if binned else "arg_backward_aux_kernel", t, 2 * k * g.m / t / 1e6, b / t / 1e6, b / t / 1e6 / pk))
            print(rows[-2], flush=True)
            print(rows[-1], flush=True)
    del g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/configs.md")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    rows = []
    it = 3 if a.quick else 5
    run_shape("cora", [64], ["sum"], "gcn", rows, 20)
    run_shape("reddit", [128] if a.quick else [32, 64, 128, 256], ["sum", "mean", "max", "min"], "uniform", rows, it)
    if not a.quick:
        run_shape("products", [100, 256, 47], ["sum"], "gcn", rows, it)
        run_shape("proteins", [128], ["mean", "sum"], None, rows, it)
        run_shape("amazon", [200], ["max", "sum"], "uniform", rows, it)
    with open(a.out, "w") as f:
        f.write(f"HBM peak used for the fraction: {peak():.1f} GB/s (MEASURED_PEAKS.json). Times: CUDA events, best variant of\n"
                f"the on-device autotune, 1 B200. GB/s = B_alg / t (SURVEY 8d); > 1.0 of HBM peak means X is L2-resident.\n\n")
        f.write("| shape | nodes | nnz | K | reduce | pass | kernel variant | ms | GFLOP/s | eff. GB/s | x HBM peak |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]:,} | {r[2]:,} | {r[3]} | {r[4]} | {r[5]} | {r[6]} | {r[7]:.3f} | {r[8]:,.0f} | {r[9]:,.0f} | {r[10]:.2f} |\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
