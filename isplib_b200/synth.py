"""Synthetic graphs with the shapes the reference's benchmarks use.

There is no network, so the datasets the reference loads
(/root/reference/tests/cpu/dataset_loader.py:8-142) are replaced by seeded generators
that reproduce their shapes (node/edge counts pinned by
tests/cpu/dataset_tester.ipynb and tests/cpu/tmp/error.log:56):

    cora      2,708 nodes      10,556 nnz
    reddit    232,965 nodes    114,615,892 nnz
    products  2,449,029 nodes  123,718,280 nnz   (100 features, 47 classes)
    proteins  132,534 nodes    79,122,504 nnz
    amazon    1,569,960 nodes  264,339,468 nnz   (200 features)

Degree sequence: heavy-tailed (log-normal, or Zipf-like for the power-law shapes),
rescaled and integer-corrected so that sum(deg) == nnz exactly, every row >= 1 entry
unless ``empty_frac`` > 0.  Columns: uniform in [0, N), sorted within each row,
duplicates allowed (SparseTensor semantics; the README fixture has one).  Values:
None, U(-1, 1), or GCN's symmetric normalisation 1/sqrt(d_i d_j).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

SHAPES = {
    #            nodes      nnz          law        sigma/alpha
    "cora":     (2_708,     10_556,      "lognormal", 0.8),
    "reddit":   (232_965,   114_615_892, "lognormal", 1.3),
    "products": (2_449_029, 123_718_280, "zipf",      2.1),
    "proteins": (132_534,   79_122_504,  "lognormal", 1.2),
    "amazon":   (1_569_960, 264_339_468, "zipf",      2.1),
}


@dataclass
class SynthGraph:
    name: str
    m: int
    n: int
    rowptr: torch.Tensor          # int64 [m+1]
    col: torch.Tensor             # int64 [nnz]
    value: Optional[torch.Tensor]  # fp32 [nnz] or None
    max_degree: int
    gini: float

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def to(self, device):
        return SynthGraph(self.name, self.m, self.n, self.rowptr.to(device), self.col.to(device),
                          None if self.value is None else self.value.to(device), self.max_degree, self.gini)

    def sparse_tensor(self):
        import sys
        import isplib_b200  # noqa: F401  (installs the torch_sparse shim if needed)
        ST = sys.modules["torch_sparse"].SparseTensor
        return ST(rowptr=self.rowptr, col=self.col, value=self.value, sparse_sizes=(self.m, self.n),
                  is_sorted=True)


def degree_sequence(m: int, nnz: int, law: str, param: float, gen: torch.Generator, device,
                    empty_frac: float = 0.0) -> torch.Tensor:
    """int64 [m], sum == nnz exactly, >= 1 except for an `empty_frac` share of rows."""
    if m == 0:
        return torch.zeros(0, dtype=torch.int64, device=device)
    if law == "lognormal":
        w = torch.exp(param * torch.randn(m, generator=gen, device=device, dtype=torch.float64))
    elif law == "zipf":
        # Pareto tail with exponent `param` on the degree pdf  (P[d > t] ~ t^-(param-1))
        u = torch.rand(m, generator=gen, device=device, dtype=torch.float64).clamp_(min=1e-12)
        w = u.pow(-1.0 / (param - 1.0))
        w = torch.minimum(w, w.sum() * 1e-3)   # no single row owns more than ~0.1% of the edges
    elif law == "uniform":
        w = torch.ones(m, device=device, dtype=torch.float64)
    else:
        raise ValueError(law)
    n_empty = int(round(empty_frac * m))
    if n_empty > 0:
        idx = torch.randperm(m, generator=gen, device=device)[:n_empty]
        w[idx] = 0.0
    live = w > 0
    n_live = int(live.sum())
    if nnz < n_live:
        raise ValueError("nnz smaller than the number of non-empty rows")
    if n_live == 0:
        return torch.zeros(m, dtype=torch.int64, device=device)
    # every live row gets 1, the rest is shared proportionally to w
    extra = nnz - n_live
    share = w / w.sum() * extra
    deg = torch.floor(share).to(torch.int64)
    deg[live] += 1
    rem = nnz - int(deg.sum())
    if rem > 0:  # hand the remainder to the rows with the largest fractional parts
        frac = (share - torch.floor(share))
        frac[~live] = -1.0
        top = torch.topk(frac, rem).indices
        deg[top] += 1
    assert int(deg.sum()) == nnz
    return deg


def gini(deg: torch.Tensor) -> float:
    if deg.numel() == 0:
        return 0.0
    d, _ = torch.sort(deg.to(torch.float64))
    n = d.numel()
    tot = float(d.sum())
    if tot == 0:
        return 0.0
    idx = torch.arange(1, n + 1, device=d.device, dtype=torch.float64)
    return float((2.0 * (idx * d).sum() / (n * tot)) - (n + 1.0) / n)


def make_graph(name_or_m, nnz: Optional[int] = None, *, n: Optional[int] = None, law: Optional[str] = None,
               param: Optional[float] = None, values: Optional[str] = "uniform", seed: int = 0,
               device="cpu", empty_frac: float = 0.0, scale: float = 1.0) -> SynthGraph:
    """``make_graph('reddit')`` or ``make_graph(1000, 20000, law='lognormal', param=1.0)``.

    ``values``: None | 'uniform' (U(-1,1)) | 'gcn' (1/sqrt(d_i d_j)) | 'ones'.
    ``scale`` shrinks a named shape (nodes and nnz) by that factor, for tests.
    """
    if isinstance(name_or_m, str):
        name = name_or_m
        m0, nnz0, law0, param0 = SHAPES[name]
        m = max(1, int(round(m0 * scale)))
        nnz = max(m, int(round(nnz0 * scale))) if nnz is None else nnz
        law = law or law0
        param = param0 if param is None else param
    else:
        name = "custom"
        m = int(name_or_m)
        assert nnz is not None
        law = law or "lognormal"
        param = 1.0 if param is None else param
    n = m if n is None else int(n)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)   # the reference scripts seed 0 (tests/cpu/gcn-sparse.py:10-12)

    deg = degree_sequence(m, nnz, law, param, gen, dev, empty_frac)
    rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr[1:])
    row = torch.repeat_interleave(torch.arange(m, device=dev, dtype=torch.int64), deg)
    col = torch.randint(0, max(n, 1), (nnz,), generator=gen, device=dev, dtype=torch.int64)
    # sort columns inside each row: one global sort of the (row, col) key
    key = row * n + col
    key, _ = torch.sort(key)
    col = key - row * n
    del key

    value = None
    if values == "uniform":
        value = torch.rand(nnz, generator=gen, device=dev, dtype=torch.float32) * 2.0 - 1.0
    elif values == "ones":
        value = torch.ones(nnz, device=dev, dtype=torch.float32)
    elif values == "gcn":
        d = deg.clamp(min=1).to(torch.float32)
        dc = torch.bincount(col, minlength=n).clamp(min=1).to(torch.float32)
        value = (d[row] * dc[col]).rsqrt()
    elif values is not None:
        raise ValueError(values)
    del row
    return SynthGraph(name, m, n, rowptr, col, value, int(deg.max()) if m else 0, gini(deg))


def relabel_by_degree(rowptr: torch.Tensor, col: torch.Tensor, value: Optional[torch.Tensor], n: int):
    """P A P^T with the nodes renumbered by descending out-degree (square graphs): the layout in
    which an even row split is badly skewed.  Returns (rowptr, col, value), rows sorted by column."""
    m = rowptr.numel() - 1
    assert m == n, "relabel_by_degree needs a square adjacency"
    dev = col.device
    deg = rowptr[1:] - rowptr[:-1]
    perm = torch.argsort(deg, descending=True, stable=True)            # new node i = old node perm[i]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(m, device=dev)
    new_row = inv[torch.repeat_interleave(torch.arange(m, device=dev), deg)]
    new_col = inv[col]
    order = torch.argsort(new_row * n + new_col)
    out_rowptr = torch.zeros_like(rowptr)
    out_rowptr[1:] = torch.cumsum(deg[perm], 0)
    return out_rowptr, new_col[order].contiguous(), None if value is None else value[order].contiguous()


def algorithmic_bytes(m: int, nnz: int, k: int, has_value: bool, reduce: str = "sum") -> int:
    """B_alg of SURVEY.md section 8d / BASELINE.md section 3 (int32 CSR, fp32)."""
    b = 4 * (m + 1) + 4 * nnz + (4 * nnz if has_value else 0) + 4 * k * nnz + 4 * k * m
    if reduce in ("max", "min"):
        b += 8 * k * m
    return b


def min_bytes(m: int, n: int, nnz: int, k: int, has_value: bool, reduce: str = "sum") -> int:
    """B_min: as above with X counted once (perfect cache)."""
    return algorithmic_bytes(m, nnz, k, has_value, reduce) - 4 * k * nnz + 4 * k * n


def spmm_flops(nnz: int, k: int) -> int:
    return 2 * nnz * k


def sqrt_int(x: int) -> int:
    return int(math.isqrt(x))
