"""Generates tests/golden/callers_*.npz: what the reference's CALLERS compute around its SpMM, with the
reference's own operator layer underneath (oracle/_ref/_fusedmm_cpu.so = unmodified csrc/fusedmm.cpp over
the restated kernel) and plain torch ops for the rest -- the unfused math the fused epilogue of
isplib_b200 (isplib_b200_spmm_csr_fused, isplib::fusedmm_spmm_fused, isplib_b200.nn) has to reproduce:

  gcn   relu( A (x W^T) + b )        GCNConv(normalize=False) + F.relu   tests/cpu/gcn-sparse.py:61-68
  gin   (1 + eps) x + A x            GINConv's aggregation input          tests/cpu/gin-sparse.py:73-78
  sage  W_l mean_j(x_j) + b + W_r x  SAGEConv(aggr=mean)                  tests/cpu/graphSAGE-sparse.py:71-78

with the gradients w.r.t. every dense input for a fixed upstream gradient.  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_callers.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF_SO, csr_from_coo  # noqa: E402


def graph(seed, M, max_deg, with_value):
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(0, max_deg, (M,), generator=g)
    deg[3] = 3 * max_deg
    row = torch.repeat_interleave(torch.arange(M), deg)
    col = torch.randint(0, M, (row.numel(),), generator=g)
    val = (torch.rand(row.numel(), generator=g) * 2 - 1) if with_value else None
    return g, csr_from_coo(row, col, val, M, M)


def ref_sum(row, rowptr, col, value, x, M):
    """torch.ops.isplib.fusedmm_spmm called the way the reference's plugin calls it (isplib/__init__.py:69-80,141)."""
    nnz = col.numel()
    v = value if value is not None else torch.ones(nnz)
    csr2csc = torch.argsort(col * M + row, stable=True)
    colptr = torch.zeros(M + 1, dtype=torch.int64)
    colptr[1:] = torch.cumsum(torch.bincount(col, minlength=M), 0)
    return torch.ops.isplib.fusedmm_spmm(row, rowptr, col, v, colptr, csr2csc, x,
                                         v.view(-1, 1).index_select(0, csr2csc).view(-1), row.index_select(0, csr2csc))


def ref_mean(row, rowptr, col, value, x, M):
    nnz = col.numel()
    v = value if value is not None else torch.ones(nnz)
    rowcount = rowptr[1:] - rowptr[:-1]
    csr2csc = torch.argsort(col * M + row, stable=True)
    colptr = torch.zeros(M + 1, dtype=torch.int64)
    colptr[1:] = torch.cumsum(torch.bincount(col, minlength=M), 0)
    new_row = row.index_select(0, csr2csc)
    deg_perm = rowcount.index_select(0, row).index_select(0, csr2csc).float().clamp_(min=1)   # the true adjoint
    w = v.view(-1, 1).index_select(0, csr2csc).view(-1).div(deg_perm)
    return torch.ops.isplib.fusedmm_spmm_mean(row, rowptr, col, v, rowcount, colptr, csr2csc, x, new_row, w)


def main():
    if not os.path.exists(REF_SO):
        sys.exit(f"{REF_SO} missing: run `make -C oracle ref` first")
    torch.ops.load_library(REF_SO)
    torch.set_num_threads(1)
    for name, M, K_in, K_out, with_value in (("a", 60, 20, 47, True), ("b", 90, 33, 16, False)):
        g, (row, rowptr, col, val) = graph(7 + M, M, 9, with_value)
        x0 = torch.randn(M, K_in, generator=g)
        W0 = torch.randn(K_out, K_in, generator=g) * 0.3
        b0 = torch.randn(K_out, generator=g)
        Wr0 = torch.randn(K_out, K_in, generator=g) * 0.3
        go_out = torch.randn(M, K_out, generator=g)
        go_in = torch.randn(M, K_in, generator=g)
        res = dict(rowptr=rowptr.numpy(), col=col.numpy(), x=x0.numpy(), W=W0.numpy(), b=b0.numpy(), Wr=Wr0.numpy(),
                   grad_out=go_out.numpy(), grad_in=go_in.numpy(), has_value=np.array(with_value), eps=np.array(0.25))
        if with_value:
            res["value"] = val.numpy()
        # ---- GCN layer + ReLU
        x, W, b = (t.clone().requires_grad_(True) for t in (x0, W0, b0))
        out = torch.relu(ref_sum(row, rowptr, col, val, x @ W.t(), M) + b)
        out.backward(go_out)
        res.update(gcn_out=out.detach().numpy(), gcn_grad_x=x.grad.numpy().copy(), gcn_grad_W=W.grad.numpy().copy(),
                   gcn_grad_b=b.grad.numpy().copy())
        # ---- GIN aggregation input (values dropped, like adj_t.set_value(None))
        x = x0.clone().requires_grad_(True)
        out = (1 + 0.25) * x + ref_sum(row, rowptr, col, None, x, M)
        out.backward(go_in)
        res.update(gin_out=out.detach().numpy(), gin_grad_x=x.grad.numpy().copy())
        # ---- SAGE-mean layer (values dropped)
        x, W, b, Wr = (t.clone().requires_grad_(True) for t in (x0, W0, b0, Wr0))
        out = ref_mean(row, rowptr, col, None, x, M) @ W.t() + b + x @ Wr.t()
        out.backward(go_out)
        res.update(sage_out=out.detach().numpy(), sage_grad_x=x.grad.numpy().copy(), sage_grad_Wl=W.grad.numpy().copy(),
                   sage_grad_b=b.grad.numpy().copy(), sage_grad_Wr=Wr.grad.numpy().copy())
        path = os.path.join(HERE, f"callers_{name}_{M}x{K_in}x{K_out}.npz")
        np.savez_compressed(path, **res)
        print(f"{os.path.basename(path)}: nnz={col.numel()} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
