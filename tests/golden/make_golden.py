"""Generates tests/golden/*.npz by running the REFERENCE's own operator layer.

Run in the build container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

What runs: oracle/_ref/_fusedmm_cpu.so = the reference's UNMODIFIED csrc/fusedmm.cpp
(op registration, fusedmm_spmm_fw, the four autograd Functions incl. their backward)
linked against oracle/fusedmm_oracle.c, because the kernel library the reference links
(OnixHoque/FusedMM_Extended@spmm_variant, configure:2-7) is not in the tree.  So these
vectors pin the wrapper + autograd contract exactly and the inner kernel up to the
restatement ("parity unpinned" for the latter, see oracle/fusedmm_oracle.c).

The tensors handed to the ops are built the way the reference's plugin builds them
(isplib/__init__.py:58-106): row/rowcount/colptr/csr2csc with torch_sparse's
definitions, value[csr2csc] & row[csr2csc] for sum, new_row/new_rowcount for mean.
This script does not import isplib_b200.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_fusedmm_cpu.so")


def csr_from_coo(row, col, val, M, N):
    key = row * N + col
    perm = torch.argsort(key, stable=True)
    row, col = row[perm], col[perm]
    val = None if val is None else val[perm]
    rowptr = torch.zeros(M + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
    return row, rowptr, col, val


def cases():
    g = torch.Generator().manual_seed(0)
    out = {}
    # 1. README.md:105-116 -- 3x3 with a duplicate (0,0) entry, +0/-0 tie in min
    out["readme_3x3"] = dict(
        row=torch.tensor([2, 0, 1, 0, 0]), col=torch.tensor([1, 0, 0, 2, 0]),
        val=torch.tensor([3., 3., 4., 2., -2.]), M=3, N=3,
        mat=torch.tensor([[1., 0., 2.], [4., 0., 0.], [0., 3., 0.]]))
    # 2. gpu/fusedmm.cu:60-118 -- 16 rows x 1 entry on the diagonal, val 2.0
    mat = torch.full((16, 16), 10.0)
    mat[torch.arange(16), torch.arange(16)] = 10.0 * (torch.arange(16) + 1)
    out["gpu_diag16"] = dict(row=torch.arange(16), col=torch.arange(16), val=torch.full((16,), 2.0),
                             M=16, N=16, mat=mat)
    # 3. random heavy-tailed, rectangular, odd K, signed values
    M, N, K = 50, 40, 7
    deg = torch.clamp((torch.exp(1.0 * torch.randn(M, generator=g)) * 6).long(), 1, 60)
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, N, (row.numel(),), generator=g)
    out["powerlaw_50x40_k7"] = dict(row=row, col=col, val=torch.rand(row.numel(), generator=g) * 2 - 1,
                                    M=M, N=N, mat=torch.randn(N, K, generator=g))
    # 4. no values (SAGE/GIN path: adj_t.set_value(None))
    M, N, K = 64, 64, 16
    deg = torch.randint(1, 12, (M,), generator=g)
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, N, (row.numel(),), generator=g)
    out["novalue_64_k16"] = dict(row=row, col=col, val=None, M=M, N=N, mat=torch.randn(N, K, generator=g))
    # 5. empty rows (first, middle, last) -- pins the lowest()/max() + sentinel convention
    M, N, K = 12, 9, 5
    deg = torch.tensor([0, 3, 0, 0, 5, 1, 0, 2, 7, 0, 4, 0])
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, N, (row.numel(),), generator=g)
    out["emptyrows_12x9_k5"] = dict(row=row, col=col, val=torch.rand(row.numel(), generator=g) * 2 - 1,
                                    M=M, N=N, mat=torch.randn(N, K, generator=g))
    # 6. one very long row (700 entries > default segment length 256) among short ones
    M, N, K = 9, 300, 33
    deg = torch.tensor([2, 700, 1, 0, 257, 256, 3, 513, 31])
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, N, (row.numel(),), generator=g)
    out["longrow_9x300_k33"] = dict(row=row, col=col, val=torch.rand(row.numel(), generator=g) * 2 - 1,
                                    M=M, N=N, mat=torch.randn(N, K, generator=g))
    # 7. many exact ties: small-integer values and features -> tie-break rule is exercised
    M, N, K = 20, 10, 12
    deg = torch.randint(1, 40, (M,), generator=g)
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, N, (row.numel(),), generator=g)
    out["ties_20x10_k12"] = dict(row=row, col=col,
                                 val=torch.randint(-2, 3, (row.numel(),), generator=g).float(),
                                 M=M, N=N, mat=torch.randint(-2, 3, (N, K), generator=g).float())
    return out


def run_case(c):
    M, N = c["M"], c["N"]
    row, rowptr, col, val = csr_from_coo(c["row"].long(), c["col"].long(), c["val"], M, N)
    mat = c["mat"].float().contiguous()
    K = mat.shape[1]
    nnz = col.numel()
    # torch_sparse storage definitions
    rowcount = rowptr[1:] - rowptr[:-1]
    csr2csc = torch.argsort(col * M + row, stable=True)
    colptr = torch.zeros(N + 1, dtype=torch.int64)
    colptr[1:] = torch.cumsum(torch.bincount(col, minlength=N), 0)
    # isplib/__init__.py:51-57: value None -> fp32 ones
    value = val if val is not None else torch.ones(nnz, dtype=torch.float32)
    gen = torch.Generator().manual_seed(1234)
    grad_out = torch.randn(M, K, generator=gen)

    res = dict(rowptr=rowptr.numpy(), col=col.numpy(), mat=mat.numpy(), grad_out=grad_out.numpy(),
               has_value=np.array(val is not None), N=np.array(N))
    if val is not None:
        res["value"] = val.numpy()
    ops = torch.ops.isplib

    # ---- sum (isplib/__init__.py:76-80,141)
    x = mat.clone().requires_grad_(True)
    vis = value.view(-1, 1).index_select(0, csr2csc).view(-1)
    ris = row.index_select(0, csr2csc)
    o = ops.fusedmm_spmm(row, rowptr, col, value, colptr, csr2csc, x, vis, ris)
    o.backward(grad_out)
    res["sum_out"], res["sum_grad_mat"] = o.detach().numpy(), x.grad.numpy().copy()

    # ---- mean (isplib/__init__.py:83-99,151)
    x = mat.clone().requires_grad_(True)
    new_row = row.index_select(0, csr2csc)
    new_rowcount = rowcount.index_select(0, row).type(x.type())
    new_rowcount.masked_fill_(new_rowcount < 1, 1)
    new_rowcount = value.view(-1, 1).index_select(0, csr2csc).view(-1).div(new_rowcount)
    # NB the reference divides value[csr2csc] by rowcount[row] WITHOUT permuting the
    # divisor (isplib/__init__.py:86-90) -- a reference bug whenever csr2csc is not the
    # identity; record what the documented math gives instead (divisor permuted too).
    deg_perm = rowcount.index_select(0, row).index_select(0, csr2csc).type(x.type()).clamp_(min=1)
    new_rowcount_fixed = value.view(-1, 1).index_select(0, csr2csc).view(-1).div(deg_perm)
    o = ops.fusedmm_spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, x, new_row, new_rowcount_fixed)
    o.backward(grad_out)
    res["mean_out"], res["mean_grad_mat"] = o.detach().numpy(), x.grad.numpy().copy()
    x2 = mat.clone().requires_grad_(True)
    o2 = ops.fusedmm_spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, x2, new_row, new_rowcount)
    o2.backward(grad_out)
    res["mean_grad_mat_asref"] = x2.grad.numpy().copy()

    # ---- max / min (isplib/__init__.py:143,145), grads wrt mat and value
    for name, fn in (("max", ops.fusedmm_spmm_max), ("min", ops.fusedmm_spmm_min)):
        x = mat.clone().requires_grad_(True)
        v = value.clone().requires_grad_(True)
        o, arg = fn(rowptr, col, v, x)
        o.backward(grad_out)
        res[f"{name}_out"], res[f"{name}_arg"] = o.detach().numpy(), arg.numpy()
        res[f"{name}_grad_mat"], res[f"{name}_grad_value"] = x.grad.numpy().copy(), v.grad.numpy().copy()
    return res


def main():
    if not os.path.exists(REF_SO):
        sys.exit(f"{REF_SO} missing: run `make -C oracle ref` first")
    torch.ops.load_library(REF_SO)
    torch.set_num_threads(1)
    for name, c in cases().items():
        res = run_case(c)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: nnz={res['col'].shape[0]} K={res['mat'].shape[1]} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
