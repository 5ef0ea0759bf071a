"""The literal drop-in: the reference's UNMODIFIED operator layer (csrc/fusedmm.cpp) linked
against isplib_b200's `fusedMM_csr` symbol (isplib_b200/libfusedmm_b200_compat.so) instead of
the CPU kernel library -- built by `make -C oracle dropin` into
oracle/_ref/_fusedmm_cpu_on_b200.so.  The reference's CPU-tensor ops then run their forward
AND their own autograd backward on the B200 kernels; results must equal the golden vectors
(which the same operator layer produced over the CPU oracle kernel).

Runs in a subprocess: the reference registers the same op names as isplib_b200's own ops.
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN_CASES, ROOT

pytestmark = pytest.mark.gpu
SO = os.path.join(ROOT, "oracle", "_ref", "_fusedmm_cpu_on_b200.so")

CHILD = r'''
import json, sys, numpy as np, torch
so, golden_dir, names = sys.argv[1], sys.argv[2], sys.argv[3:]
torch.ops.load_library(so)
ops = torch.ops.isplib
res = {}
for name in names:
    z = np.load(f"{golden_dir}/{name}.npz")
    rowptr, col = torch.from_numpy(z["rowptr"]), torch.from_numpy(z["col"])
    M, N, nnz = rowptr.numel() - 1, int(z["N"]), col.numel()
    value = torch.from_numpy(z["value"]) if "value" in z.files else torch.ones(nnz)
    mat, go = torch.from_numpy(z["mat"]), torch.from_numpy(z["grad_out"])
    row = torch.repeat_interleave(torch.arange(M), rowptr[1:] - rowptr[:-1])
    rowcount = rowptr[1:] - rowptr[:-1]
    csr2csc = torch.argsort(col * M + row, stable=True)
    colptr = torch.zeros(N + 1, dtype=torch.long); colptr[1:] = torch.cumsum(torch.bincount(col, minlength=N), 0)
    ok = {}
    x = mat.clone().requires_grad_(True)
    o = ops.fusedmm_spmm(row, rowptr, col, value, colptr, csr2csc, x, value[csr2csc], row[csr2csc])
    o.backward(go)
    ok["sum_out"] = bool(np.allclose(o.detach().numpy(), z["sum_out"], rtol=1e-4, atol=1e-5))
    ok["sum_grad"] = bool(np.allclose(x.grad.numpy(), z["sum_grad_mat"], rtol=1e-4, atol=1e-5))
    x = mat.clone().requires_grad_(True)
    w = value[csr2csc] / rowcount[row][csr2csc].float().clamp(min=1)
    o = ops.fusedmm_spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, x, row[csr2csc], w)
    o.backward(go)
    ok["mean_out"] = bool(np.allclose(o.detach().numpy(), z["mean_out"], rtol=1e-4, atol=1e-5))
    ok["mean_grad"] = bool(np.allclose(x.grad.numpy(), z["mean_grad_mat"], rtol=1e-4, atol=1e-5))
    for red, fn in (("max", ops.fusedmm_spmm_max), ("min", ops.fusedmm_spmm_min)):
        x = mat.clone().requires_grad_(True)
        o, arg = fn(rowptr, col, value, x)
        o.backward(go)
        ok[red + "_out"] = bool(np.array_equal(o.detach().numpy(), z[red + "_out"]))
        ok[red + "_arg"] = bool(np.array_equal(arg.numpy(), z[red + "_arg"]))
        ok[red + "_grad"] = bool(np.allclose(x.grad.numpy(), z[red + "_grad_mat"], rtol=1e-4, atol=1e-5))
    res[name] = ok
print("RESULT " + json.dumps(res))
'''


def test_reference_operator_layer_runs_on_b200_kernels():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/_fusedmm_cpu_on_b200.so not built (needs /root/reference at build time)")
    env = dict(os.environ, ISPLIB_B200_SKIP_EXTENSION="1")
    p = subprocess.run([sys.executable, "-c", CHILD, SO, os.path.join(ROOT, "tests", "golden")] + GOLDEN_CASES,
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[7:])
    bad = {n: [k for k, v in ok.items() if not v] for n, ok in res.items() if not all(ok.values())}
    assert not bad, f"mismatches: {bad}"
    assert sorted(res) == sorted(GOLDEN_CASES)
