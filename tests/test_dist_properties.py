"""Property tests (hypothesis) of the host-side partitioning logic in isplib_b200/dist.py: whatever the graph
and the number of ranks, every row / column / stored entry is owned exactly once and the layouts invert."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from isplib_b200.dist import (bounds_width, even_bounds, group_runs, nnz_balanced_bounds, owner_groups,
                              slice_position, split_row_block, split_row_block_by_owner)

FAST = settings(max_examples=60, deadline=None)


def is_partition(bounds, n, world):
    return (len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == n and
            all(bounds[i] <= bounds[i + 1] for i in range(world)))


@FAST
@given(n=st.integers(0, 5000), world=st.integers(1, 16))
def test_even_bounds_partition_the_range(n, world):
    b = even_bounds(n, world)
    assert is_partition(b, n, world)
    assert bounds_width(b) == max(1, (n + world - 1) // world)


@FAST
@given(deg=st.lists(st.integers(0, 400), min_size=0, max_size=300), world=st.integers(1, 9))
def test_nnz_balanced_bounds_partition_and_balance(deg, world):
    rowptr = torch.zeros(len(deg) + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.tensor(deg, dtype=torch.int64), 0) if deg else rowptr[1:]
    b = nnz_balanced_bounds(rowptr, world)
    m, nnz = len(deg), int(rowptr[-1])
    assert is_partition(b, m, world)
    if nnz and world > 1:
        # no cut is further than one (heaviest) row away from its ideal prefix count
        for p in range(1, world):
            assert abs(int(rowptr[b[p]]) - (p * nnz) // world) <= max(deg)


@FAST
@given(n=st.integers(1, 3000), world=st.integers(1, 8), data=st.data())
def test_slice_position_is_the_gathered_layout(n, world, data):
    bounds = sorted(data.draw(st.lists(st.integers(0, n), min_size=world - 1, max_size=world - 1)))
    bounds = [0] + bounds + [n]
    width = bounds_width(bounds)
    idx = torch.arange(n, dtype=torch.int64)
    pos = slice_position(idx, bounds, width)
    # owner-major, offset inside the owner's slice; injective; strictly increasing (row order survives)
    for p in range(world):
        seg = pos[bounds[p]:bounds[p + 1]]
        assert torch.equal(seg, p * width + torch.arange(bounds[p + 1] - bounds[p]))
    assert n < 2 or bool((pos[1:] > pos[:-1]).all())


@FAST
@given(world=st.integers(1, 16), g=st.integers(1, 7), rank_seed=st.integers(0, 1000))
def test_owner_groups_cover_the_peers_in_ring_order(world, g, rank_seed):
    rank = rank_seed % world
    groups, n_groups = owner_groups(world, rank, g)
    assert groups[rank] == 0 and len(groups) == world
    if world == 1:
        assert n_groups == 1
        return
    assert n_groups == 1 + min(g, world - 1)
    ring = [groups[(rank + d) % world] for d in range(1, world)]
    assert ring == sorted(ring) and ring[0] == 1 and ring[-1] == n_groups - 1      # nearer peers arrive first
    sizes = np.bincount(ring)[1:]
    assert sizes.max() - sizes.min() <= 1                                          # as even as possible
    starts, grp = group_runs(groups, 10)
    assert starts[0] == 0 and all(a < b for a, b in zip(starts, starts[1:]))
    assert all(x != y for x, y in zip(grp, grp[1:]))                               # maximal runs
    expanded = []
    for i, s in enumerate(starts):
        e = starts[i + 1] if i + 1 < len(starts) else world * 10
        expanded += [grp[i]] * ((e - s) // 10)
    assert expanded == groups


@settings(max_examples=25, deadline=None)
@given(m=st.integers(1, 60), n=st.integers(1, 60), world=st.integers(1, 5), seed=st.integers(0, 10_000),
       with_value=st.booleans())
def test_column_owner_split_keeps_every_entry_once(m, n, world, seed, with_value):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, 9, size=m)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    col = np.concatenate([np.sort(rng.integers(0, n, size=d)) for d in deg]) if nnz else np.zeros(0, dtype=np.int64)
    val = rng.standard_normal(nnz).astype(np.float32) if with_value else None
    rp_t, co_t = torch.from_numpy(rowptr), torch.from_numpy(col.astype(np.int64))
    va_t = None if val is None else torch.from_numpy(val)
    row_bounds = nnz_balanced_bounds(rp_t, world)
    col_bounds = even_bounds(n, world)
    Rc = bounds_width(col_bounds)
    seen = np.zeros(nnz, dtype=np.int64)
    for rank in range(world):
        local, remote, deg_f, R = split_row_block(rp_t, co_t, va_t, rank, world, n, row_bounds, col_bounds)
        blocks, _, _ = split_row_block_by_owner(rp_t, co_t, va_t, rank, world, n, row_bounds, col_bounds)
        r0, r1 = row_bounds[rank], row_bounds[rank + 1]
        assert torch.equal(deg_f[: r1 - r0], torch.from_numpy(np.maximum(deg[r0:r1], 1).astype(np.float32)))
        for blk, kind in [(local, "local"), (remote, "remote")] + [(b, q) for q, b in enumerate(blocks)]:
            assert blk.rowptr.numel() == R + 1 and int(blk.rowptr[-1]) == blk.nnz
            eid = blk.edge_ids.numpy().astype(np.int64)
            rows = np.repeat(np.arange(R), np.diff(blk.rowptr.numpy()))
            # the entry really is entry `eid` of the global CSR: same row, same column, same value
            assert np.array_equal(np.searchsorted(rowptr, eid, side="right") - 1, rows + r0)
            c = blk.col.numpy().astype(np.int64)
            if kind == "local":
                gcol = c + col_bounds[rank]
            elif kind == "remote":
                gcol = np.array([col_bounds[p // Rc] + p % Rc for p in c], dtype=np.int64)
                assert not np.any((gcol >= col_bounds[rank]) & (gcol < col_bounds[rank + 1]))
            else:
                gcol = c + col_bounds[kind]
            assert np.array_equal(gcol, col[eid])
            if val is not None:
                assert np.array_equal(blk.val.numpy(), val[eid])
            if kind in ("local", "remote"):
                np.add.at(seen, eid, 1)
            # edge ids ascend inside every row: the CSR order (= the tie-break order) survives the split
            for r in range(R):
                seg = eid[int(blk.rowptr[r]):int(blk.rowptr[r + 1])]
                assert np.all(np.diff(seg) > 0)
    assert np.all(seen == 1)
