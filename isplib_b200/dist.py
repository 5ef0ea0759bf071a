"""Row-partitioned multi-GPU SpMM (new functionality; the reference is single-process,
SURVEY.md section 8e).

One process per GPU (``torchrun``), ``torch.distributed`` over NCCL/NVLink.

* A (M x N) is 1-D partitioned by rows: rank p owns the contiguous rows
  ``[row_bounds[p], row_bounds[p+1])``.  ``balance="rows"`` (default) is the even split
  ``R = ceil(M / P)``; ``balance="nnz"`` cuts the ranges so that every rank holds about nnz/P
  stored entries (power-law graphs whose ids are sorted by locality or degree are badly skewed
  under an even row split: degree-sorted Reddit-shape graph on 2 GPUs, forward 2.76 -> 1.86 ms).
  The backward operator (A^T) has to reuse the same bounds, because grad_x is owned like x; it
  is balanced too when A is symmetric (undirected graphs), not in general.  Rows of X / out are owned the same way (square A: the same
  bounds; otherwise the columns are split evenly) and zero-padded to the widest range
  (``R`` / ``Rc``) so the all-gather is regular.
* Each rank's row block is split by COLUMN OWNER into a *local* CSR block (columns the
  rank already holds) and a *remote* CSR block (everything else).
* forward:   all-gather X slices on a side stream  ||  local-block SpMM
             -> wait -> remote-block SpMM with ISPLIB_FLAG_ACCUMULATE into the same rows.
  sum: plain accumulate.  mean: both blocks as sums, the second call divides by the full
  row degree (``row_divisor``).  max/min: both blocks carry their GLOBAL edge ids
  (``edge_ids``), the merge is (value, edge id)-lexicographic, so out/arg are bit-identical
  to the single-GPU result.
* backward (sum/mean): the same operator built from A^T (each rank owns the rows of A^T
  = columns of A in its range) applied to the all-gathered grad_out.
  backward (max/min): local arg-scatter into a zeroed [P*R, K] partial, then
  reduce-scatter(sum).

* ``mode="fused"``: NO collective call at all.  The rank's whole row block is ONE CSR over the
  owner-major gathered columns, X lives in a double-buffered symmetric-memory allocation, and ONE
  kernel (``isplib_b200_spmm_csr_gather``) pushes the rank's own slice into every peer's buffer over
  NVLink with its first CTAs while the others multiply -- each work item as soon as the rows of its
  arrival group (a K tile of the launch, or a group of column owners with a grouped plan) have
  landed.  Flow control by credit words ("I have started step e"), arrivals by release-adds on the
  peer's counters; nobody waits for a barrier kernel.  max/min/mean need no cross-block merge any
  more (a row is one CSR row again), so results equal the 1-GPU kernel's.
* ``mode="nccl"`` keeps the all-gather + two-block path below; ``mode="auto"`` (the default on CUDA
  with more than one rank) builds both and times them per feature width on first use.

The block SpMM is injectable (``block_spmm``) so the partitioning / merge logic can be
tested on CPU with gloo; the default is the CUDA C ABI (isplib_b200.capi) and there is no
CPU fallback in the product path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

SUM, MAX, MIN, MEAN = 0, 1, 2, 3
REDUCE_CODE = {"sum": SUM, "add": SUM, "max": MAX, "min": MIN, "mean": MEAN}
FLAG_ACCUMULATE = 0x1


@dataclass
class CsrBlock:
    rowptr: torch.Tensor            # int32 [R+1]
    col: torch.Tensor               # int32 [nnz_b]  (block-local column index)
    val: Optional[torch.Tensor]     # fp32  [nnz_b] or None
    edge_ids: torch.Tensor          # int32 [nnz_b]  global edge id of each entry
    plan: object = None             # capi.Plan (CUDA path only)

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


def rows_per_rank(m: int, world: int) -> int:
    return (m + world - 1) // world


def even_bounds(n: int, world: int):
    """[0, R, 2R, ..., n] with R = ceil(n / world) (clipped): the even split."""
    R = rows_per_rank(n, world)
    return [min(p * R, n) for p in range(world + 1)]


def nnz_balanced_bounds(rowptr: torch.Tensor, world: int):
    """Contiguous row ranges with about nnz/world stored entries each:
    bounds[p] = the row whose prefix count is closest to p * nnz / world."""
    m = rowptr.numel() - 1
    nnz = int(rowptr[-1])
    if world == 1 or m == 0 or nnz == 0:
        return even_bounds(m, world)
    rp = rowptr.to(torch.int64).cpu()
    bounds = [0]
    for p in range(1, world):
        target = (p * nnz) // world
        hi = int(torch.searchsorted(rp, torch.tensor(target, dtype=torch.int64)))   # first row with prefix >= target
        hi = min(max(hi, 0), m)
        lo = max(hi - 1, 0)
        b = hi if abs(int(rp[hi]) - target) <= abs(int(rp[lo]) - target) else lo
        bounds.append(min(max(b, bounds[-1]), m))
    bounds.append(m)
    return bounds


def bounds_width(bounds) -> int:
    return max(1, max(bounds[p + 1] - bounds[p] for p in range(len(bounds) - 1)))


def slice_position(idx: torch.Tensor, bounds, width: int) -> torch.Tensor:
    """Global index -> position in the all-gathered, width-padded layout: owner * width + offset."""
    b = torch.as_tensor(bounds, dtype=torch.int64, device=idx.device)
    owner = torch.bucketize(idx, b[1:-1], right=True)
    return owner * width + (idx - b[owner])


def split_row_block(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor],
                    rank: int, world: int, n_cols: int, row_bounds=None, col_bounds=None
                    ) -> Tuple[CsrBlock, CsrBlock, torch.Tensor, int]:
    """From the GLOBAL CSR (int64, any device) build this rank's (local, remote) blocks.

    Returns (local, remote, row_degree[R] fp32 (clamped >= 1), R).  Local block columns are
    rebased to the rank's own X slice; remote block columns index the gathered [P*Rc, K]
    matrix (Rc = widest column range), where own-slice columns never appear.
    """
    m = rowptr.numel() - 1
    row_bounds = even_bounds(m, world) if row_bounds is None else row_bounds
    col_bounds = even_bounds(n_cols, world) if col_bounds is None else col_bounds
    R = bounds_width(row_bounds)
    Rc = bounds_width(col_bounds)
    r0, r1 = row_bounds[rank], row_bounds[rank + 1]
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    dev = col.device
    sub_col = col[e0:e1].to(torch.int64)
    sub_val = None if val is None else val[e0:e1]
    eid = torch.arange(e0, e1, device=dev, dtype=torch.int64)
    deg = (rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1])
    row = torch.repeat_interleave(torch.arange(r1 - r0, device=dev, dtype=torch.int64), deg)
    c0, c1 = col_bounds[rank], col_bounds[rank + 1]
    is_local = (sub_col >= c0) & (sub_col < c1)

    def make(mask, cols):
        cnt = torch.bincount(row[mask], minlength=R) if mask.numel() else torch.zeros(R, dtype=torch.int64, device=dev)
        rp = torch.zeros(R + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(cnt, 0)
        return CsrBlock(rp.to(torch.int32), cols.to(torch.int32),
                        None if sub_val is None else sub_val[mask].contiguous(), eid[mask].to(torch.int32))

    local = make(is_local, sub_col[is_local] - c0)
    remote = make(~is_local, slice_position(sub_col[~is_local], col_bounds, Rc))
    full_deg = torch.zeros(R, dtype=torch.float32, device=dev)
    full_deg[: r1 - r0] = deg.to(torch.float32)
    full_deg.clamp_(min=1.0)
    return local, remote, full_deg, R


def split_row_block_by_owner(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor],
                             rank: int, world: int, n_cols: int, row_bounds=None, col_bounds=None):
    """Like split_row_block but one CSR block PER COLUMN OWNER q (columns rebased to q's slice),
    for the per-source pipelined gather.  Returns (blocks[world], row_degree, R)."""
    m = rowptr.numel() - 1
    row_bounds = even_bounds(m, world) if row_bounds is None else row_bounds
    col_bounds = even_bounds(n_cols, world) if col_bounds is None else col_bounds
    R = bounds_width(row_bounds)
    r0, r1 = row_bounds[rank], row_bounds[rank + 1]
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    dev = col.device
    sub_col = col[e0:e1].to(torch.int64)
    sub_val = None if val is None else val[e0:e1]
    eid = torch.arange(e0, e1, device=dev, dtype=torch.int64)
    deg = (rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1])
    row = torch.repeat_interleave(torch.arange(r1 - r0, device=dev, dtype=torch.int64), deg)
    cb = torch.as_tensor(col_bounds, dtype=torch.int64, device=dev)
    owner = torch.bucketize(sub_col, cb[1:-1], right=True)
    blocks = []
    for q in range(world):
        mask = owner == q
        cnt = torch.bincount(row[mask], minlength=R) if mask.numel() else torch.zeros(R, dtype=torch.int64, device=dev)
        rp = torch.zeros(R + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(cnt, 0)
        blocks.append(CsrBlock(rp.to(torch.int32), (sub_col[mask] - col_bounds[q]).to(torch.int32),
                               None if sub_val is None else sub_val[mask].contiguous(), eid[mask].to(torch.int32)))
    full_deg = torch.zeros(R, dtype=torch.float32, device=dev)
    full_deg[: r1 - r0] = deg.to(torch.float32)
    full_deg.clamp_(min=1.0)
    return blocks, full_deg, R


def owner_groups(world: int, rank: int, n_remote_groups: int):
    """Arrival group of every owner's slice for `rank`: 0 for its own, 1..G for the peers by ring
    distance (rank+1 first), split as evenly as possible.  Ring order makes the ranks read from
    different peers at any moment."""
    n_remote = world - 1
    G = max(1, min(n_remote_groups, n_remote)) if n_remote > 0 else 0
    groups = [0] * world
    for d in range(1, world):
        groups[(rank + d) % world] = 1 + ((d - 1) * G) // n_remote
    return groups, G + 1


def group_runs(groups, slice_rows: int):
    """Contiguous owner runs of equal group in absolute (owner-major) column order:
    (run_start[], run_group[]) for isplib_b200_plan_build_grouped."""
    starts, grp = [], []
    for o, g in enumerate(groups):
        if not grp or grp[-1] != g:
            starts.append(o * slice_rows)
            grp.append(g)
    return starts, grp


class _PeerBuffers:
    """The double-buffered gathered-X allocation of one feature width, visible to every rank.

    Layout (fp32 words): [buf 0: world*Rc*Kp][buf 1: world*Rc*Kp][credit: 64 uint32]
    [arrive parity 0: 16 uint32][arrive parity 1: 16 uint32].
    Real ranks: torch symmetric memory (CUDA VMM peer mappings over NVLink), rendezvoused once per
    width.  ``emulated``: a dict shared by the emulated ranks of ONE process (tests on one GPU): the
    "peers" are ordinary local tensors, the kernel's pushes are local copies."""

    CREDIT_WORDS, ARRIVE_WORDS = 64, 16

    def __init__(self, world, rank, Rc, K, device, group, emulated=None):
        self.world, self.rank, self.Rc, self.K = world, rank, Rc, K
        self.Kp = (K + 7) // 8 * 8
        self.buf_words = world * Rc * self.Kp
        n_words = 2 * self.buf_words + self.CREDIT_WORDS + 2 * self.ARRIVE_WORDS
        if emulated is None:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            self.t = symm.empty(n_words, dtype=torch.float32, device=device)
            self.t.zero_()
            torch.cuda.synchronize(device)
            self.hdl = symm.rendezvous(self.t, g)
            dist.barrier(group=g)                  # nobody pushes into words that are not zeroed yet
            self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        else:
            ts = emulated.setdefault(("bufs", K), [torch.zeros(n_words, dtype=torch.float32, device=device)
                                                   for _ in range(world)])
            self.t = ts[rank]
            self.ptrs = [int(t.data_ptr()) for t in ts]
        self.bufs = [self.t[b * self.buf_words:(b + 1) * self.buf_words].view(world * Rc, self.Kp) for b in (0, 1)]

    def peer_x(self, b):
        return [p + 4 * b * self.buf_words for p in self.ptrs]

    def peer_credit(self):
        return [p + 4 * 2 * self.buf_words for p in self.ptrs]

    def peer_arrive(self, b):
        return [p + 4 * (2 * self.buf_words + self.CREDIT_WORDS + b * self.ARRIVE_WORDS) for p in self.ptrs]

    def own_slice(self, b):
        return self.bufs[b][self.rank * self.Rc:(self.rank + 1) * self.Rc, :self.K]


def emulated_step(ops, xs, reduce: str = "sum", epilogues=None):
    """One forward of several ranks emulated in ONE process / on ONE GPU (tests): every rank pushes
    its slice first (phase 1), then every rank multiplies (phase 2, still waiting on the arrival
    counters the pushes bumped).  Real ranks do both in one launch (phase 0).  Returns [(out, arg)]."""
    code = REDUCE_CODE[reduce]
    is_arg = code in (MAX, MIN)
    outs = []
    for op, x in zip(ops, xs):
        K = x.size(1)
        out = torch.empty((op.R, K), dtype=torch.float32, device=x.device)
        arg = torch.empty((op.R, K), dtype=torch.int64, device=x.device) if is_arg else None
        op._forward_fused(x, code, out, arg, phase=1)
        outs.append((out, arg))
    for i, (op, x, (out, arg)) in enumerate(zip(ops, xs, outs)):
        op._forward_fused(x, code, out, arg, phase=2, epilogue=None if epilogues is None else epilogues[i])
    return outs


def _cuda_block_spmm(reduce_code, block: CsrBlock, x, out, arg_out, flags, row_divisor, arg_sentinel, variant=-1,
                     **epilogue):
    """One launch of isplib_b200_spmm_csr_fused on a column block.  ``epilogue`` (bias / addend / addend_scale /
    relu) is only ever given to the LAST launch that writes a row block: the kernel's final store applies it
    after the ACCUMULATE merge and the mean division."""
    from . import capi
    if block.plan is None:
        block.plan = capi.Plan(block.rowptr, block.nnz)
    if variant < 0 and block.nnz >= (1 << 16) and os.environ.get("ISPLIB_B200_AUTOTUNE", "1") != "0" \
            and not torch.cuda.is_current_stream_capturing():
        # on-device variant selection per column block, once per (reduction, width, alignment class):
        # every rank times its own block (no collective inside), like the single-GPU op layer does
        key = (int(reduce_code), x.size(1), x.stride(0), x.data_ptr() % 32)
        tuned = block.__dict__.setdefault("tuned", {})
        if key not in tuned:
            tuned[key], _ = capi.spmm_autotune(reduce_code, block.rowptr, block.col, block.val, x, block.plan, iters=2)
        variant = tuned[key]
    return capi.spmm_csr(reduce_code, block.rowptr, block.col, block.val, x, block.plan, variant,
                         out=out, arg_out=arg_out, flags=flags, row_divisor=row_divisor,
                         edge_ids=block.edge_ids, arg_sentinel=arg_sentinel, **epilogue)


def make_epilogue(bias=None, addend=None, addend_scale: float = 1.0, relu: bool = False):
    """{} when there is nothing to fuse, else the keyword arguments of the last kernel launch."""
    if bias is None and addend is None and not relu:
        return {}
    return dict(bias=bias, addend=addend, addend_scale=float(addend_scale), relu=bool(relu))


def _epilogue_columns(epi: dict, c0: int, c1: int) -> dict:
    """The epilogue of feature columns [c0, c1) (K-chunked launches)."""
    if not epi:
        return epi
    e = dict(epi)
    if e.get("bias") is not None:
        e["bias"] = e["bias"][c0:c1]
    if e.get("addend") is not None:
        e["addend"] = e["addend"][:, c0:c1]
    return e


def _epilogue_torch(out: torch.Tensor, epi: dict) -> torch.Tensor:
    """The same epilogue as separate passes, for the one path whose launches are not ordered last-writer
    (the opt-in copy-engine pipeline)."""
    if epi.get("addend") is not None:
        out.add_(epi["addend"], alpha=epi.get("addend_scale", 1.0))
    if epi.get("bias") is not None:
        out.add_(epi["bias"])
    if epi.get("relu"):
        out.relu_()
    return out


class RowPartitionedSpMM:
    """out_local = (A @ X)[own rows] with X given as this rank's row slice."""

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, value: Optional[torch.Tensor], n_cols: int,
                 group=None, device=None, block_spmm: Optional[Callable] = None, overlap: bool = True,
                 pipelined: Optional[bool] = None, balance: str = "nnz", row_bounds=None, col_bounds=None,
                 mode: Optional[str] = None, emulate=None):
        self.group = group
        if emulate is not None:       # (world, rank, shared dict): several ranks emulated in one process
            self.world, self.rank, self._emulated = int(emulate[0]), int(emulate[1]), emulate[2]
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            self._emulated = None
        self.m = rowptr.numel() - 1
        self.n = int(n_cols)
        self.nnz = int(col.numel())
        self.device = torch.device(device) if device is not None else col.device
        self.block_spmm = block_spmm or _cuda_block_spmm
        self.overlap = overlap
        if balance not in ("nnz", "rows"):
            raise ValueError(f"balance must be 'nnz' or 'rows', got {balance!r}")
        if row_bounds is None:
            row_bounds = nnz_balanced_bounds(rowptr, self.world) if balance == "nnz" else even_bounds(self.m, self.world)
        if col_bounds is None:
            # square A (a graph): X rows are owned like A rows, so a layer's output slice is the
            # next layer's input slice; otherwise split the columns evenly
            col_bounds = list(row_bounds) if self.n == self.m else even_bounds(self.n, self.world)
        self.row_bounds, self.col_bounds = list(row_bounds), list(col_bounds)
        if (len(self.row_bounds) != self.world + 1 or len(self.col_bounds) != self.world + 1 or
                self.row_bounds[0] != 0 or self.row_bounds[-1] != self.m or
                self.col_bounds[0] != 0 or self.col_bounds[-1] != self.n):
            raise ValueError("row_bounds / col_bounds must be world+1 non-decreasing offsets from 0 to m / n")
        local, remote, deg, R = split_row_block(rowptr, col, value, self.rank, self.world, self.n,
                                                self.row_bounds, self.col_bounds)
        self.R = R
        self.Rc = bounds_width(self.col_bounds)
        mv = lambda b: CsrBlock(b.rowptr.to(self.device), b.col.to(self.device),
                                None if b.val is None else b.val.to(self.device), b.edge_ids.to(self.device))
        self.local, self.remote = mv(local), mv(remote)
        self.row_degree = deg.to(self.device)
        self.comm_stream = torch.cuda.Stream(self.device) if (self.device.type == "cuda" and overlap) else None
        self.variant = -1
        # per-source pipelined gather (copy-engine P2P over NVLink through symmetric memory, one
        # SpMM per column owner as its slice lands): opt-in, or ISPLIB_B200_DIST_PIPELINED=1
        if pipelined is None:
            pipelined = os.environ.get("ISPLIB_B200_DIST_PIPELINED", "0") == "1"
        self.pipelined = bool(pipelined) and self.world > 1 and self.device.type == "cuda"
        self.owner_blocks = None
        if self.pipelined:
            blocks, _, _ = split_row_block_by_owner(rowptr, col, value, self.rank, self.world, self.n,
                                                    self.row_bounds, self.col_bounds)
            self.owner_blocks = [mv(b) for b in blocks]
        self.k_chunk = None        # feature-chunk width of the all-gather/SpMM pipeline (None = auto)
        self._gather_bufs = {}     # persistent all-gather receive buffers, keyed by (K, dtype, device)
        # "fused": gather + SpMM in one kernel, no collective.  "nccl": all-gather on a side stream +
        # local / remote block kernels.  "auto" (the default wherever the fused kernel can run): both
        # are built and the faster one is MEASURED per feature width on first use (on-device selection,
        # like the single-GPU variants): the all-gather hides completely behind the local block while
        # that block is a large share of the work (2-4 ranks), the fused kernel wins beyond
        if mode is None:
            mode = os.environ.get("ISPLIB_B200_DIST_MODE") or (
                "auto" if (self.world > 1 and self.device.type == "cuda" and block_spmm is None and not self.pipelined
                           and self._emulated is None)
                else ("fused" if self._emulated is not None else "nccl"))
        if mode not in ("auto", "fused", "nccl"):
            raise ValueError(f"mode must be 'auto', 'fused' or 'nccl', got {mode!r}")
        self.mode = mode
        self._mode_choice = {}      # (K, is_arg) -> "fused" | "nccl" (+ the two measured times)
        if mode in ("auto", "fused"):
            r0, r1 = self.row_bounds[self.rank], self.row_bounds[self.rank + 1]
            e0, e1 = int(rowptr[r0]), int(rowptr[r1])
            rp = torch.zeros(self.R + 1, dtype=torch.int64, device=rowptr.device)
            rp[1:r1 - r0 + 1] = rowptr[r0 + 1:r1 + 1] - e0
            rp[r1 - r0 + 1:] = e1 - e0
            pos = slice_position(col[e0:e1].to(torch.int64), self.col_bounds, self.Rc)   # monotone per row: owner-major keeps the order
            self.full = CsrBlock(rp.to(torch.int32).to(self.device), pos.to(torch.int32).to(self.device),
                                 None if value is None else value[e0:e1].contiguous().to(self.device),
                                 torch.arange(e0, e1, dtype=torch.int32, device=self.device))
            n_remote_groups = int(os.environ.get("ISPLIB_B200_DIST_GROUPS", "3"))
            self.owner_group, self.n_groups = owner_groups(self.world, self.rank, n_remote_groups)
            # the arrival group THIS rank's slice belongs to at every peer (the pusher needs it)
            self.my_group_at_peer = [owner_groups(self.world, q, n_remote_groups)[0][self.rank] for q in range(self.world)]
            # copy CTAs: one CTA pushes ~8 GB/s of LOADED bytes (4 x 16 B in flight per thread), every loaded
            # vector is stored to world-1 peers, and the link saturates near 590 GB/s of egress
            # (profiles/r2_push_probe_n2.json: 16/32/64/128 CTAs -> 129/254/492/587 GB/s to one peer; NCCL's
            # all-gather of the same slices: 469 GB/s) -> 128 CTAs for one peer, 32 from 7 peers up, where
            # 32, 64 and 128 measure the same (profiles/r2_dist_probe_n8_push.json)
            default_ctas = max(32, min(128, (256 // max(1, self.world - 1)) // 32 * 32))
            self.copy_ctas = int(os.environ.get("ISPLIB_B200_DIST_COPY_CTAS", str(default_ctas)))
            # what the arrival groups of the fused kernel are: "tiles" = the K tiles of the launch
            # (rows stay whole, plain plan; needs >= 2 tiles, i.e. K >= 128 in 64-wide tiles) or
            # "owners" = column owners (grouped plan, rows split per group); "auto" = tiles when possible
            self.gather_groups = os.environ.get("ISPLIB_B200_DIST_GATHER", "auto")
            self._plain_plan = None
            self._fpv_cache = {}
            self._peer = {}            # K -> _PeerBuffers
            self._epoch = {}           # K -> launches so far on that buffer set
            self._gflags = {}          # K -> (flags uint32[8], status uint32[1])

    # rows this rank owns (without padding)
    @property
    def own_rows(self) -> int:
        return self.row_bounds[self.rank + 1] - self.row_bounds[self.rank]

    def row_range(self, rank: Optional[int] = None) -> Tuple[int, int]:
        """Global rows of A / out owned by `rank` (default: this rank)."""
        r = self.rank if rank is None else rank
        return self.row_bounds[r], self.row_bounds[r + 1]

    def col_range(self, rank: Optional[int] = None) -> Tuple[int, int]:
        """Global rows of X (columns of A) owned by `rank` (default: this rank)."""
        r = self.rank if rank is None else rank
        return self.col_bounds[r], self.col_bounds[r + 1]

    def pad_x(self, x_own: torch.Tensor) -> torch.Tensor:
        """[own_cols, K] -> [Rc, K] zero padded (the regular slice the all-gather needs)."""
        if x_own.size(0) == self.Rc:
            return x_own.contiguous()
        out = torch.zeros((self.Rc, x_own.size(1)), dtype=x_own.dtype, device=x_own.device)
        out[: x_own.size(0)] = x_own
        return out

    def pad_out(self, y_own: torch.Tensor) -> torch.Tensor:
        """[rows <= R, K] -> contiguous [R, K] zero padded (an epilogue addend shaped like the output slice)."""
        if y_own.size(0) == self.R and y_own.stride(1) == 1:
            return y_own
        out = torch.zeros((self.R, y_own.size(1)), dtype=y_own.dtype, device=y_own.device)
        n = min(self.R, y_own.size(0))
        out[:n] = y_own[:n]
        return out

    def _all_gather(self, x_slice: torch.Tensor, persistent: bool = False) -> torch.Tensor:
        if self.world == 1:
            return x_slice
        shape = (self.world * self.Rc, x_slice.size(1))
        if persistent and x_slice.is_cuda:
            # one receive buffer per feature width, kept for the life of the operator: a fresh
            # multi-GB torch.empty per call on the side stream (plus record_stream, which delays
            # its reuse) makes the caching allocator grow and cudaMalloc in the hot path.  Safe
            # because every forward starts with comm_stream.wait_stream(current), i.e. after the
            # previous call's remote-block SpMM has read the buffer, and nothing is saved from it.
            key = (shape[1], x_slice.dtype, x_slice.device)
            gathered = self._gather_bufs.get(key)
            if gathered is None:
                gathered = torch.empty(shape, dtype=x_slice.dtype, device=x_slice.device)
                self._gather_bufs[key] = gathered
        else:
            gathered = torch.empty(shape, dtype=x_slice.dtype, device=x_slice.device)
        dist.all_gather_into_tensor(gathered, x_slice, group=self.group)
        return gathered

    def forward(self, x_slice: torch.Tensor, reduce: str = "sum", aux=None, epilogue: Optional[dict] = None):
        """x_slice: [Rc, K] (use pad_x).  Returns (out [R, K], arg_out [R, K] int64 or None);
        rows beyond own_rows are padding (zeros / init values).  ``aux``: max/min in fused mode only -- a
        dict that receives 'arg_col' / 'arg_val' ([R, K]: the winner's column, as a position in the
        gathered layout, and its value) for the streamed backward scatter.  ``epilogue`` (make_epilogue):
        relu?(result + addend_scale * addend[R, K] + bias[K]) in the final store of the last launch that
        writes the rows -- the fused gather kernel, or the remote-block kernel of the NCCL path."""
        epi = epilogue or {}
        if reduce not in REDUCE_CODE:
            raise ValueError(f"isplib_b200.dist: reduce must be one of {sorted(REDUCE_CODE)}, got {reduce!r}")
        if x_slice.dim() != 2 or x_slice.size(0) != self.Rc or x_slice.dtype != torch.float32:
            raise ValueError(f"isplib_b200.dist: x_slice must be this rank's padded fp32 slice [{self.Rc}, K] (pad_x), "
                             f"got {tuple(x_slice.shape)} {x_slice.dtype}")
        if x_slice.device.type != self.device.type or (
                self.device.index is not None and x_slice.device.index not in (None, self.device.index)):
            raise ValueError(f"isplib_b200.dist: x_slice is on {x_slice.device}, the partition on {self.device}")
        if epi.get("bias") is not None and tuple(epi["bias"].shape) != (x_slice.size(1),):
            raise ValueError(f"isplib_b200.dist: bias must have shape [{x_slice.size(1)}], got {tuple(epi['bias'].shape)}")
        if epi.get("addend") is not None and tuple(epi["addend"].shape) != (self.R, x_slice.size(1)):
            raise ValueError(f"isplib_b200.dist: addend must have shape [{self.R}, {x_slice.size(1)}] (pad_out), "
                             f"got {tuple(epi['addend'].shape)}")
        code = REDUCE_CODE[reduce]
        is_arg = code in (MAX, MIN)
        K = x_slice.size(1)
        out = torch.empty((self.R, K), dtype=torch.float32, device=x_slice.device)
        arg = torch.empty((self.R, K), dtype=torch.int64, device=x_slice.device) if is_arg else None
        inner = SUM if code == MEAN else code
        div = self.row_degree if code == MEAN else None

        if self.world == 1:
            self.block_spmm(inner, self.local, x_slice, out, arg, 0, div, self.nnz, self.variant, **epi)
            return out, arg

        if self.mode_for(K, reduce, x_slice) == "fused":
            return self._forward_fused(x_slice, code, out, arg, aux=aux if is_arg else None, epilogue=epi)

        if self.pipelined:
            out, arg = self._forward_pipelined(x_slice, inner, div, out, arg)
            return (_epilogue_torch(out, epi) if epi else out), arg

        # K-chunk pipeline: the all-gather of feature chunk c+1 runs on the comm stream while the
        # SpMM of chunk c runs on the compute stream, so only the first chunk's transfer is
        # exposed; a 64-wide chunk is also the K tile that keeps an [N, 64] slab L2-resident.
        chunks = self._k_chunks(K)
        if self.comm_stream is not None:
            cur = torch.cuda.current_stream(x_slice.device)
            self.comm_stream.wait_stream(cur)
            gathered, events = [], []
            with torch.cuda.stream(self.comm_stream):
                for (c0, c1) in chunks:
                    xc = x_slice if len(chunks) == 1 else x_slice[:, c0:c1].contiguous()
                    # chunks of one call are in flight together, so only the unchunked default
                    # can share one persistent buffer
                    gathered.append(self._all_gather(xc, persistent=(len(chunks) == 1)))
                    ev = torch.cuda.Event()
                    ev.record(self.comm_stream)
                    events.append(ev)
            for ci, (c0, c1) in enumerate(chunks):
                xo, oo = x_slice[:, c0:c1], out[:, c0:c1]
                ao = None if arg is None else arg[:, c0:c1]
                self.block_spmm(inner, self.local, xo, oo, ao, 0, None, self.nnz, self.variant)   # overlaps the gather
                cur.wait_event(events[ci])
                if len(chunks) > 1:
                    gathered[ci].record_stream(cur)
                self.block_spmm(inner, self.remote, gathered[ci], oo, ao, FLAG_ACCUMULATE, div, self.nnz, self.variant,
                                **(epi if len(chunks) == 1 else _epilogue_columns(epi, c0, c1)))
            return out, arg
        gathered = self._all_gather(x_slice)
        self.block_spmm(inner, self.local, x_slice, out, arg, 0, None, self.nnz, self.variant)
        self.block_spmm(inner, self.remote, gathered, out, arg, FLAG_ACCUMULATE, div, self.nnz, self.variant, **epi)
        return out, arg

    def mode_for(self, K: int, reduce: str = "sum", x_slice: Optional[torch.Tensor] = None) -> str:
        """'fused' or 'nccl' for this feature width.  mode='auto' times both once per (K, arg/additive)
        -- 3 warm-up + 10 timed forwards each, max over ranks, every rank takes the same decision."""
        if self.mode != "auto":
            return self.mode
        code = REDUCE_CODE[reduce]
        key = (K, code in (MAX, MIN))
        hit = self._mode_choice.get(key)
        if hit is not None:
            return hit[0]
        if x_slice is None:
            return "fused"
        times = {}
        dev = x_slice.device
        WARM, ITERS = 3, 10       # the first launches of either path run cold (plans, symmetric buffers, L2): 1 + 3 was too noisy
        for m in ("fused", "nccl"):
            self._mode_choice[key] = (m,)
            try:
                for _ in range(WARM):
                    self.forward(x_slice, reduce)
            except Exception as ex:      # e.g. no symmetric memory on this box: the NCCL path still works
                if m != "fused":
                    raise
                import warnings
                warnings.warn(f"isplib_b200: fused gather kernel unavailable ({ex!r}); using the NCCL path")
                times[m] = float("inf")
                continue
            dist.barrier(group=self.group)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ITERS):
                self.forward(x_slice, reduce)
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / ITERS], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            times[m] = float(t.item())
        best = min(times, key=times.get)
        self._mode_choice[key] = (best, round(times["fused"], 4) if times["fused"] != float("inf") else None,
                                  round(times["nccl"], 4))
        return best

    def _fused_plan(self):
        from . import capi
        if self.full.plan is None:
            starts, grp = group_runs(self.owner_group, self.Rc)
            self.full.plan = capi.GroupedPlan(self.full.rowptr, self.full.col, starts, grp, self.n_groups)
        return self.full.plan

    def peer_buffers(self, K: int) -> "_PeerBuffers":
        """The symmetric gathered-X allocation for feature width K (collective on first use per K)."""
        pb = self._peer.get(K)
        if pb is None:
            pb = _PeerBuffers(self.world, self.rank, self.Rc, K, self.device, self.group, self._emulated)
            self._peer[K] = pb
            self._epoch[K] = 0
            self._gflags[K] = torch.zeros(1, dtype=torch.int32, device=self.device)     # timeout status
            self._tiles_used = getattr(self, "_tiles_used", {})
        return pb

    def next_input_slice(self, K: int) -> torch.Tensor:
        """The [Rc, K] view the NEXT forward of width K reads this rank's slice from: a producer that
        writes there (instead of handing forward() a separate tensor) saves the staging copy."""
        pb = self.peer_buffers(K)
        return pb.own_slice((self._epoch[K] + 1) & 1)

    def _tile_variant(self, K: int, code, x) -> int:
        """Variant id of the 64-wide-K-tile lean kernel if the fused forward can run in tile mode at
        this width (>= 2 whole tiles), else -1."""
        from . import capi
        if self.gather_groups == "owners" or K < 128 or K % 64 != 0 or K // 64 > 8:
            return -1
        v = capi.variant_names().index("lean256/w4/kt64")
        ok = capi.lib().isplib_b200_variant_supported(v, code, K, x.stride(0), K, x.data_ptr(), x.data_ptr())
        return v if ok else -1

    def _forward_fused(self, x_slice, code, out, arg, phase=0, aux=None, epilogue=None):
        """phase 0: the product path (push + multiply in one launch).  phases 1 / 2 split a step into
        its push and its multiply half: only for several ranks emulated on ONE GPU (emulated_step)."""
        from . import capi
        if self._emulated is not None and phase == 0:
            raise RuntimeError("emulated ranks cannot run the fused step in one launch: use dist.emulated_step()")
        K = x_slice.size(1)
        pb = self.peer_buffers(K)
        if phase != 2:
            self._epoch[K] += 1
        epoch = self._epoch[K]
        b = epoch & 1
        own = pb.own_slice(b)
        if phase != 2 and x_slice.data_ptr() != own.data_ptr():
            own.copy_(x_slice)                     # staging copy into the peer-visible buffer (Rc x K)
        full = self.full
        xg = pb.bufs[b][:, :K]
        plan, variant, groups = self.fused_plan_and_variant(K, code)
        tile_mode = groups == "K tiles"
        self._tiles_used[K] = tile_mode
        arg_col = arg_val = None
        if aux is not None:
            arg_col = torch.empty((self.R, K), dtype=torch.int32, device=x_slice.device)
            arg_val = torch.empty((self.R, K), dtype=torch.float32, device=x_slice.device) if full.val is not None else None
            aux["arg_col"], aux["arg_val"] = arg_col, arg_val
        capi.spmm_csr_gather(code, full.rowptr, full.col, full.val, xg, plan,
                             world=self.world, rank=self.rank, peer_x=pb.peer_x(b), peer_arrive=pb.peer_arrive(b),
                             peer_credit=pb.peer_credit(), owner_group=self.owner_group,
                             my_group_at_peer=self.my_group_at_peer, slice_rows=self.Rc, status=self._gflags[K],
                             epoch=epoch, parity_launch=(epoch + 1) // 2, tile_mode=tile_mode, phase=phase,
                             copy_ctas=self.copy_ctas, variant=variant, out=out,
                             arg_out=arg, edge_ids=full.edge_ids, arg_sentinel=self.nnz,
                             arg_col=arg_col, arg_val=arg_val, **(epilogue or {}))
        return out, arg

    def phase_split(self, x_slice, reduce: str = "sum", steps: int = 10):
        """{'forward', 'multiply_only'} ms (max over ranks): the fused forward vs the same kernel and
        plan on an already gathered X (no pushes, no waits); the difference is what the gather costs."""
        from . import capi
        assert self.mode_for(x_slice.size(1), reduce) == "fused"
        K = x_slice.size(1)
        dev = x_slice.device

        def timed(fn):
            fn()
            if dist.is_initialized() and self._emulated is None:
                dist.barrier(group=self.group)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            if dist.is_initialized() and self._emulated is None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return round(float(t.item()), 4)

        fwd = timed(lambda: self.forward(x_slice, reduce))
        pb = self.peer_buffers(K)
        xg = pb.bufs[self._epoch[K] & 1][:, :K]
        full = self.full
        plan, variant, groups = self.fused_plan_and_variant(K, reduce)
        out = torch.empty((self.R, K), device=dev)
        arg = torch.empty((self.R, K), dtype=torch.int64, device=dev) if reduce in ("max", "min") else None
        mul = timed(lambda: capi.spmm_csr(reduce, full.rowptr, full.col, full.val, xg, plan, variant, out=out,
                                          arg_out=arg, edge_ids=full.edge_ids, arg_sentinel=self.nnz))
        return {"forward": fwd, "multiply_only": mul, "gather_exposed": round(max(0.0, fwd - mul), 4),
                "arrival_groups": groups}

    def fused_plan_and_variant(self, K: int, reduce="sum"):
        """(plan, variant id, 'K tiles' | 'column owners') the fused forward uses at width K."""
        from . import capi
        code = capi.REDUCE_CODE[reduce] if isinstance(reduce, str) else int(reduce)
        key = (K, code, self.variant)
        hit = self._fpv_cache.get(key)
        if hit is not None:
            return hit
        pb = self.peer_buffers(K)
        tv = self._tile_variant(K, code, pb.bufs[0][:, :K]) if self.variant < 0 else -1
        if tv >= 0:
            if self._plain_plan is None:
                self._plain_plan = capi.Plan(self.full.rowptr, self.full.nnz)
            res = (self._plain_plan, tv, "K tiles")
        else:
            res = (self._fused_plan(), self.variant, "column owners")
        self._fpv_cache[key] = res
        return res

    def check_status(self):
        """Raises if a fused-gather kernel gave up waiting for a peer (4 s timeout inside the kernel).
        Synchronises; call it outside hot loops."""
        if self.mode in ("fused", "auto"):
            for K, status in self._gflags.items():
                if int(status.item()) != 0:
                    raise RuntimeError(f"isplib_b200: fused gather (K={K}) timed out waiting for a peer's slice")

    def _forward_pipelined(self, x_slice, inner, div, out, arg):
        """X slices travel peer-to-peer over NVLink by the copy engines (torch symmetric memory,
        no SMs taken from the SpMM); the column block of source rank q is multiplied as soon as
        q's slice has landed, while the next slice is in flight.  The block kernels accumulate
        into the same rows, so they are chained by events: the merge order is fixed (own block
        first, then ring order) -> deterministic; max/min stay bit-identical to one GPU."""
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        gname = group.group_name
        if not symm.is_symm_mem_enabled_for_group(gname):
            symm.enable_symm_mem_for_group(gname)
        K = x_slice.size(1)
        ag_out = torch.empty((self.world * self.Rc, K), dtype=x_slice.dtype, device=x_slice.device)
        last_src = (self.rank + self.world - 1) % self.world
        state = {"prev": None}

        def consumer(shard, src):
            stream = torch.cuda.current_stream(x_slice.device)
            if state["prev"] is not None:
                stream.wait_event(state["prev"])
            flags = 0 if src == self.rank else FLAG_ACCUMULATE
            self.block_spmm(inner, self.owner_blocks[src], shard, out, arg, flags,
                            div if src == last_src else None, self.nnz, self.variant)
            ev = torch.cuda.Event()
            ev.record(stream)
            state["prev"] = ev

        symm._pipelined_all_gather_and_consume(x_slice.contiguous(), consumer, ag_out, gname, ag_out_needed=False)
        torch.cuda.current_stream(x_slice.device).wait_event(state["prev"])
        return out, arg

    def _k_chunks(self, K: int):
        kc = self.k_chunk
        if kc is None and os.environ.get("ISPLIB_B200_DIST_KCHUNK"):
            kc = int(os.environ["ISPLIB_B200_DIST_KCHUNK"])
        if kc is None:
            # default: ONE all-gather of the full width, overlapped with the local block only.
            # Measured on 4xB200 (Reddit-shape K=128): no chunking 1.02 ms, 32-wide chunks
            # 1.73 ms, 64-wide 9.96 ms -- NCCL's all-gather kernels and the SpMM fight for the
            # same SMs/L2 when they really run concurrently, so the pipeline is opt-in
            # (k_chunk attribute or ISPLIB_B200_DIST_KCHUNK).
            kc = K
        kc = max(4, (int(kc) + 3) // 4 * 4)
        if kc >= K:
            return [(0, K)]
        return [(c0, min(K, c0 + kc)) for c0 in range(0, K, kc)]

    def launches_per_forward(self, K: Optional[int] = None, reduce: str = "sum") -> int:
        """how many of OUR kernels one forward launches (for bench.py's gpu_launches)."""
        if self.world > 1 and (self.mode == "fused" or (K is not None and self.mode_for(K, reduce) == "fused")):
            return 1                      # gather + SpMM are one kernel
        n = 1 if self.world == 1 else 2   # one spmm_seg_kernel per column block (x feature chunks)
        return n


# ----------------------------------------------------------------------------------------
# autograd over the partitioned operator
# ----------------------------------------------------------------------------------------
def transpose_csr(rowptr: torch.Tensor, col: torch.Tensor, value: Optional[torch.Tensor], n_cols: int,
                  mean_weights: bool = False):
    """A^T in CSR form (colptr, row[csr2csc], value[csr2csc]) -- isplib/__init__.py:76-99 of
    the reference -- with plain torch ops (one-time, any device).  mean_weights divides the
    permuted values by max(deg(row),1): the weights of the mean backward."""
    m = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(m, device=col.device, dtype=torch.int64), deg)
    csr2csc = torch.argsort(col * m + row, stable=True)
    colptr = torch.zeros(n_cols + 1, dtype=torch.int64, device=col.device)
    colptr[1:] = torch.cumsum(torch.bincount(col, minlength=n_cols), 0)
    row_t = row[csr2csc]
    val_t = None if value is None else value[csr2csc]
    if mean_weights:
        w = deg.clamp(min=1).to(torch.float32)[row_t]
        val_t = (1.0 / w) if val_t is None else val_t / w
    return colptr, row_t, val_t


class _OffsetArray:
    """A rank's slice [e0, e0 + local.numel()) of a conceptually GLOBAL 1-D array of `total` elements:
    exactly what the partitioning code touches of `col` / `value` (one slice inside the rank's own edge
    range), so a rank that only ever loaded its own rows can be partitioned like one that holds the whole
    graph (rank-local ingest, `from_local_rows`)."""

    def __init__(self, local: torch.Tensor, e0: int, total: int):
        self.local, self.e0, self.total = local, int(e0), int(total)

    def numel(self) -> int:
        return self.total

    @property
    def device(self):
        return self.local.device

    @property
    def dtype(self):
        return self.local.dtype

    def __getitem__(self, sl):
        assert isinstance(sl, slice) and sl.step in (None, 1)
        a = 0 if sl.start is None else sl.start
        b = self.total if sl.stop is None else sl.stop
        assert self.e0 <= a <= b <= self.e0 + self.local.numel(), "only the rank's own edge range is resident"
        return self.local[a - self.e0:b - self.e0]


def exchange_by_owner(arrays, owner: torch.Tensor, group=None):
    """Every element i of the parallel 1-D `arrays` goes to rank owner[i]; returns the received arrays,
    concatenated in source-rank order.  Point-to-point (batch_isend_irecv): works on gloo and NCCL."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = owner.device
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world).to(torch.int64)
    all_counts = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)                    # all_counts[s][q] = elements s sends to q
    send_off = torch.zeros(world + 1, dtype=torch.int64)
    send_off[1:] = torch.cumsum(counts.cpu(), 0)
    recv_counts = [int(all_counts[s][rank]) for s in range(world)]
    out = []
    for a in arrays:
        a_sorted = a[order].contiguous()
        recv = [torch.empty(recv_counts[s], dtype=a.dtype, device=dev) for s in range(world)]
        ops = []
        for d in range(1, world):
            dst, src = (rank + d) % world, (rank - d) % world
            chunk = a_sorted[int(send_off[dst]):int(send_off[dst + 1])]
            if chunk.numel():
                ops.append(dist.P2POp(dist.isend, chunk, dist.get_global_rank(group, dst) if group is not None else dst, group))
            if recv_counts[src]:
                ops.append(dist.P2POp(dist.irecv, recv[src], dist.get_global_rank(group, src) if group is not None else src, group))
        recv[rank] = a_sorted[int(send_off[rank]):int(send_off[rank + 1])]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        out.append(torch.cat(recv) if world > 1 else recv[0])
    return out


def distributed_transpose(rowptr_global, col_local, val_local, r0, r1, n_cols, col_bounds, mean_weights, group=None):
    """The rank's rows of A^T from every rank's rows of A, without any rank holding the whole graph:
    each stored entry (i, j, w) travels to the owner of column j, which sorts what it received by
    (j, i).  Returns (colptr GLOBAL [n_cols + 1] int64, row_t local, val_t local, first local position)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = col_local.device
    m = rowptr_global.numel() - 1
    deg = (rowptr_global[r0 + 1:r1 + 1] - rowptr_global[r0:r1]).to(dev)
    assert int(deg.sum()) == col_local.numel()
    row = torch.repeat_interleave(torch.arange(r0, r1, device=dev, dtype=torch.int64), deg)
    w = None
    if val_local is not None or mean_weights:
        w = torch.ones(col_local.numel(), dtype=torch.float32, device=dev) if val_local is None else val_local.to(torch.float32)
        if mean_weights:
            w = w / deg.clamp(min=1).to(torch.float32)[row - r0]
    cb = torch.as_tensor(col_bounds, dtype=torch.int64, device=dev)
    owner = torch.bucketize(col_local.to(torch.int64), cb[1:-1], right=True)
    arrays = [col_local.to(torch.int64), row] + ([w] if w is not None else [])
    got = exchange_by_owner(arrays, owner, group)
    j, i = got[0], got[1]
    order = torch.argsort(j * m + i, stable=True)
    row_t = i[order].contiguous()
    val_t = got[2][order].contiguous() if w is not None else None
    c0, c1 = col_bounds[rank], col_bounds[rank + 1]
    cnt_local = torch.bincount(j - c0, minlength=c1 - c0) if j.numel() else torch.zeros(c1 - c0, dtype=torch.int64, device=dev)
    # global colptr: every rank contributes the counts of its own column range
    width = bounds_width(col_bounds)
    padded = torch.zeros(width, dtype=torch.int64, device=dev)
    padded[: c1 - c0] = cnt_local
    allc = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, padded, group=group)
    counts = torch.cat([allc[q][: col_bounds[q + 1] - col_bounds[q]] for q in range(world)])
    colptr = torch.zeros(n_cols + 1, dtype=torch.int64, device=dev)
    colptr[1:] = torch.cumsum(counts, 0)
    return colptr, row_t, val_t, int(colptr[c0])


def _cuda_arg_backward(col32, val, arg, grad_out, n_rows_out, arg_sentinel):
    from . import capi
    gx, _ = capi.spmm_arg_backward(col32, val, None, arg, grad_out, n_rows_out, True, False, arg_sentinel)
    return gx


class DistSpMM:
    """Autograd-aware row-partitioned ``matmul``: ``out_slice = dist_spmm(x_slice, reduce)``.

    x_slice / out_slice are this rank's padded row slices ([Rc, K] / [R, K]).  Gradients flow
    to x_slice: sum/mean through the partitioned A^T (all-gather of grad_out + SpMM, the same
    machinery as the forward), max/min through a local arg-scatter + reduce-scatter."""

    def __init__(self, rowptr, col, value, n_cols, group=None, device=None, block_spmm=None,
                 arg_backward=None, overlap=True, pipelined=None, balance="nnz", mode=None, emulate=None,
                 row_bounds=None, _local=None):
        self.rowptr, self.col, self.value = rowptr, col, value
        self.m, self.n = rowptr.numel() - 1, int(n_cols)
        self.group, self.device = group, device
        self._kw = dict(group=group, device=device, block_spmm=block_spmm, overlap=overlap, pipelined=pipelined,
                        mode=mode, emulate=emulate)
        self._local = _local         # (r0, r1, e0) when built by from_local_rows: col / value hold the rank's own rows only
        self.fwd = RowPartitionedSpMM(rowptr, col, value, n_cols, balance=balance, row_bounds=row_bounds, **self._kw)
        self._bwd = {}
        self._arg_backward = arg_backward or _cuda_arg_backward
        self._col32 = None

    @classmethod
    def from_local_rows(cls, rowptr_local, col_local, value_local, n_cols, group=None, device=None, **kw):
        """Rank-local ingest: every rank passes ONLY its own contiguous block of rows (`rowptr_local` starting
        at 0, global column ids), in rank order -- no rank ever holds the whole graph.  Row ownership = the
        blocks as given (balance them by stored entries when you cut the file); the small global row-pointer
        array is assembled with one all-gather, the transposed partition for the backward with one exchange
        of the entries by column owner (`distributed_transpose`)."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = col_local.device
        rows = rowptr_local.numel() - 1
        info = torch.tensor([rows, int(col_local.numel())], dtype=torch.int64, device=dev)
        infos = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(infos, info, group=group)
        rows_all = [int(t[0]) for t in infos]
        nnz_all = [int(t[1]) for t in infos]
        row_bounds = [0]
        for r in rows_all:
            row_bounds.append(row_bounds[-1] + r)
        e0 = sum(nnz_all[:rank])
        m, nnz = row_bounds[-1], sum(nnz_all)
        # global rowptr (m + 1 integers -- the only global array anybody holds)
        width = max(1, max(rows_all))
        deg = torch.zeros(width, dtype=torch.int64, device=dev)
        deg[:rows] = (rowptr_local[1:] - rowptr_local[:-1]).to(torch.int64)
        degs = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(degs, deg, group=group)
        deg_all = torch.cat([degs[q][: rows_all[q]] for q in range(world)])
        rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(deg_all, 0)
        col = _OffsetArray(col_local, e0, nnz)
        value = None if value_local is None else _OffsetArray(value_local, e0, nnz)
        kw.pop("balance", None)
        kw.pop("emulate", None)
        return cls(rowptr, col, value, n_cols, group=group, device=device, row_bounds=row_bounds,
                   _local=(row_bounds[rank], row_bounds[rank + 1], e0), **kw)

    def bwd_op(self, mean: bool) -> RowPartitionedSpMM:
        if mean not in self._bwd:
            if self._local is not None:
                r0, r1, _ = self._local
                colptr, row_t, val_t, p0 = distributed_transpose(
                    self.rowptr, self.col.local, None if self.value is None else self.value.local, r0, r1, self.n,
                    self.fwd.col_bounds, mean, self.group)
                nnz = self.col.numel()
                row_t = _OffsetArray(row_t, p0, nnz)
                val_t = None if val_t is None else _OffsetArray(val_t, p0, nnz)
            else:
                colptr, row_t, val_t = transpose_csr(self.rowptr, self.col, self.value, self.n, mean_weights=mean)
            # rows of A^T are the columns of A and vice versa: swap the forward's bounds
            self._bwd[mean] = RowPartitionedSpMM(colptr, row_t, val_t, self.m, row_bounds=self.fwd.col_bounds,
                                                 col_bounds=self.fwd.row_bounds, **self._kw)
        return self._bwd[mean]

    def __call__(self, x_slice: torch.Tensor, reduce: str = "sum", bias=None, addend=None, addend_scale: float = 1.0,
                 relu: bool = False) -> torch.Tensor:
        """``relu?(A @ x + addend_scale * addend + bias)`` on this rank's rows; the epilogue rides in the final
        store of the step's last kernel (single-GPU counterpart: ``isplib_b200.fused_matmul``)."""
        return _DistSpMMFn.apply(x_slice, self, reduce, bias, addend, float(addend_scale), bool(relu))


class _DistSpMMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_slice, op: DistSpMM, reduce: str, bias=None, addend=None, addend_scale=1.0, relu=False):
        # max/min through the fused kernel: let its final store also write col[arg] / val[arg] (positions in
        # the gathered layout = the scatter targets of the backward), so the backward streams them instead of
        # gathering through a rank-local copy of the GLOBAL column array
        aux = {} if (REDUCE_CODE[reduce] in (MAX, MIN) and x_slice.requires_grad) else None
        epi = make_epilogue(None if bias is None else bias.detach().contiguous(),
                            None if addend is None else addend.detach(), addend_scale, relu)
        if epi.get("addend") is not None:
            # the kernel reads the first R rows with unit inner stride: pad a shorter / strided addend, cut a
            # longer one (an x slice of Rc > R rows handed in as GIN's self term on a non-square operator)
            if addend.stride(1) != 1 or addend.size(0) < op.fwd.R:
                epi["addend"] = op.fwd.pad_out(addend.detach())
            elif addend.size(0) > op.fwd.R:
                epi["addend"] = addend.detach()[: op.fwd.R]
        out, arg = op.fwd.forward(x_slice.contiguous(), reduce, aux=aux, epilogue=epi)
        ctx.op, ctx.reduce = op, reduce
        ctx.aux = aux if aux else None            # stays empty when the NCCL path ran
        ctx.epi = (bias is not None, addend is not None, float(addend_scale), bool(relu),
                   None if addend is None else addend.size(0))
        ctx.save_for_backward(arg if arg is not None else None, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        op, reduce = ctx.op, ctx.reduce
        code = REDUCE_CODE[reduce]
        grad_out = grad_out.contiguous()
        arg, out_saved = ctx.saved_tensors
        has_bias, has_addend, addend_scale, relu, addend_rows = ctx.epi
        if relu:
            grad_out = grad_out * (out_saved > 0).to(grad_out.dtype)
        # the epilogue's own inputs: bias sums over the rank's REAL rows (the ranks' shares are summed by
        # whoever all-reduces the parameter gradients, like every other replicated weight)
        grad_bias = grad_out[: op.fwd.own_rows].sum(0) if (has_bias and ctx.needs_input_grad[3]) else None
        grad_addend = None
        if has_addend and ctx.needs_input_grad[4]:
            grad_addend = grad_out[:addend_rows] * addend_scale
            if addend_rows > grad_addend.size(0):          # rows of the addend the kernel never read
                grad_addend = torch.nn.functional.pad(grad_addend, (0, 0, 0, addend_rows - grad_addend.size(0)))
        tail = (None, None, grad_bias, grad_addend, None, None)
        if code in (SUM, MEAN):
            t = op.bwd_op(code == MEAN)
            g = grad_out
            if g.size(0) != t.Rc:                      # pad the row slice of grad_out like X
                g = t.pad_x(g[: t.Rc])
            gx, _ = t.forward(g, "sum")                # csrc/fusedmm.cpp:285 / :375 of the reference
            return (gx,) + tail
        # max / min: scatter through the global edge ids, then sum the partials of all ranks
        f = op.fwd
        dev = grad_out.device
        if ctx.aux is not None:
            from . import capi
            n_rows = f.world * f.Rc
            binned = n_rows * grad_out.size(1) * 4 > 256 * 2**20
            partial = capi.spmm_arg_backward_aux(ctx.aux["arg_col"], ctx.aux["arg_val"], grad_out, n_rows, binned=binned)
            out = torch.empty((f.Rc, partial.size(1)), dtype=partial.dtype, device=dev)
            dist.reduce_scatter_tensor(out, partial, group=f.group)
            return (out,) + tail
        if op._col32 is None:
            # scatter target = position of the column in the width-padded gathered layout
            col_res = op.col.local if op._local is not None else op.col
            val_res = None if op.value is None else (op.value.local if op._local is not None else op.value)
            op._col32 = slice_position(col_res.to(dev).to(torch.int64), f.col_bounds, f.Rc).to(torch.int32)
            op._val_dev = None if val_res is None else val_res.to(dev)
        if op._local is not None:
            # rank-local ingest: only the rank's own edge range is resident; its rows' winners all lie inside it
            e0, n_loc = op._local[2], op._col32.numel()
            arg = torch.where(arg == f.nnz, torch.full_like(arg, n_loc), arg - e0)
            partial = op._arg_backward(op._col32, op._val_dev, arg, grad_out, f.world * f.Rc, n_loc)
        else:
            partial = op._arg_backward(op._col32, op._val_dev, arg, grad_out, f.world * f.Rc, f.nnz)
        if f.world == 1:
            return (partial,) + tail
        out = torch.empty((f.Rc, partial.size(1)), dtype=partial.dtype, device=dev)
        if dist.get_backend(f.group) == "nccl":
            dist.reduce_scatter_tensor(out, partial, group=f.group)
        else:                                          # gloo (CPU tests) has no reduce-scatter
            dist.all_reduce(partial, group=f.group)
            out.copy_(partial[f.rank * f.Rc:(f.rank + 1) * f.Rc])
        return (out,) + tail


# ----------------------------------------------------------------------------------------
# the drop-in surface for the row-partitioned mode
# ----------------------------------------------------------------------------------------
class PartitionedAdj:
    """What ``torch_sparse.matmul(adj_t, x, reduce)`` receives in place of the SparseTensor when the
    graph is row-partitioned over the ranks of a process group: with ``iSpLibPlugin.patch_pyg()``
    active the patched matmul recognises it and runs ``DistSpMM`` (fused gather + SpMM kernel or
    NCCL path, autograd included), so a model written against ``matmul(adj_t, x, reduce)`` -- the
    reference's GCN / SAGE / GIN scripts, /root/reference/tests/cpu/gcn-sparse.py:55-68 -- trains on
    N GPUs unchanged: it is handed this object and ITS rank's padded row slice of x.

    Only what those callers touch is provided: ``sparse_sizes()``, ``has_value()``,
    ``set_value(None)`` (SAGE / GIN drop the values), plus the slice helpers."""

    is_partitioned = True

    def __init__(self, rowptr, col, value, n_cols, local_rows=False, **dist_kw):
        self._args = (rowptr, col, n_cols)
        self._value = value
        self._kw = dist_kw
        self._local_rows = bool(local_rows)
        # local_rows: (rowptr, col, value) are THIS rank's contiguous block of rows only (rank-local ingest)
        self.op = (DistSpMM.from_local_rows(rowptr, col, value, n_cols, **dist_kw) if local_rows
                   else DistSpMM(rowptr, col, value, n_cols, **dist_kw))
        self._novalue_twin = None

    # --- what the callers of matmul use ---
    def sparse_sizes(self):
        return (self.op.fwd.R, self.op.fwd.Rc)

    def has_value(self) -> bool:
        return self._value is not None

    def set_value(self, value, layout=None):
        if value is not None:
            raise NotImplementedError("PartitionedAdj.set_value: only set_value(None) (what SAGEConv / GINConv do)")
        if self._value is None:
            return self
        if self._novalue_twin is None:
            rowptr, col, n_cols = self._args
            self._novalue_twin = PartitionedAdj(rowptr, col, None, n_cols, local_rows=self._local_rows, **self._kw)
        return self._novalue_twin

    def matmul(self, x_slice: torch.Tensor, reduce: str = "sum", bias=None, addend=None, addend_scale: float = 1.0,
               relu: bool = False) -> torch.Tensor:
        return self.op(x_slice, reduce, bias=bias, addend=addend, addend_scale=addend_scale, relu=relu)

    # --- slice helpers for the training script ---
    def row_range(self):
        return self.op.fwd.row_range()

    def col_range(self):
        return self.op.fwd.col_range()

    def local_slice(self, x_global: torch.Tensor) -> torch.Tensor:
        """This rank's padded [Rc, K] slice of a replicated / host-side [N, K] feature matrix."""
        c0, c1 = self.col_range()
        return self.op.fwd.pad_x(x_global[c0:c1].to(self.op.fwd.device))


def partition(adj_t, group=None, device=None, local_rows=False, **dist_kw) -> PartitionedAdj:
    """Row-partition a ``torch_sparse.SparseTensor`` over the ranks of `group`.  By default `adj_t` is the
    whole (replicated) graph and the rows are cut by stored entries; with ``local_rows=True`` it holds only
    THIS rank's contiguous block of rows ([rows_local, N], global column ids, blocks in rank order), so no
    rank ever materialises the whole graph."""
    rowptr, col, value = adj_t.csr()
    n_cols = adj_t.sparse_sizes()[1]
    return PartitionedAdj(rowptr, col, value, n_cols, local_rows=local_rows, group=group, device=device, **dist_kw)
