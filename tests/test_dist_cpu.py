"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in isplib_b200/dist.py: row
partition, column-owner split, all-gather layout, ACCUMULATE/edge-id merge protocol,
transposed operator for the backward, reduce-scatter of the arg-scatter partials.

The per-block SpMM is injected (the oracle stands in for the CUDA C ABI, honouring the same
flags), so what is tested is everything AROUND the kernel.  Checked against the single-process
oracle on the whole graph."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def oracle_block_spmm(reduce_code, block, x, out, arg_out, flags, row_divisor, arg_sentinel, variant=-1, **epi):
    """numpy stand-in for isplib_b200_spmm_csr_fused with the same contract (edge_ids,
    ISPLIB_FLAG_ACCUMULATE, row_divisor, arg_sentinel; the epilogue after merge and division)."""
    from oracle import oracle
    if epi:
        epi = dict(bias=None if epi.get("bias") is None else epi["bias"].numpy(),
                   addend=None if epi.get("addend") is None else epi["addend"].numpy(),
                   addend_scale=epi.get("addend_scale", 1.0), relu=epi.get("relu", False))
    rp = block.rowptr.numpy().astype(np.int64)
    co = block.col.numpy().astype(np.int64)
    va = None if block.val is None else block.val.numpy()
    o, a = oracle.spmm_c(rp, co, va, x.numpy(), reduce_code)
    nnz_b = co.shape[0]
    if reduce_code in (1, 2):
        gid = np.where(a == nnz_b, arg_sentinel, block.edge_ids.numpy().astype(np.int64)[np.minimum(a, max(nnz_b - 1, 0))]
                       if nnz_b else arg_sentinel)
        if flags & 1:
            po, pa = out.numpy(), arg_out.numpy()
            better = (o > po) if reduce_code == 1 else (o < po)
            take = better | ((o == po) & (gid < pa))
            o, gid = np.where(take, o, po), np.where(take, gid, pa)
        if epi:
            o = oracle.apply_epilogue(o, **epi)
        out.copy_(torch.from_numpy(o))
        arg_out.copy_(torch.from_numpy(gid))
    else:
        if flags & 1:
            o = out.numpy() + o
        if row_divisor is not None:
            o = o / row_divisor.numpy()[:, None]
        o = o.astype(np.float32)
        if epi:
            o = oracle.apply_epilogue(o, **epi)
        out.copy_(torch.from_numpy(o))
    return out, arg_out


def oracle_arg_backward(col32, val, arg, grad_out, n_rows_out, arg_sentinel):
    from oracle import oracle
    gx, _ = oracle.arg_backward(col32.numpy().astype(np.int64), None if val is None else val.numpy(), None,
                                arg.numpy(), grad_out.numpy(), n_rows_out)
    return torch.from_numpy(gx)


def make_graph(seed, M, N, with_value):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, 30, size=M)
    deg[3] = 400
    rowptr = np.zeros(M + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    row = np.repeat(np.arange(M), deg)
    col = rng.integers(0, N, size=int(rowptr[-1]))
    order = np.lexsort((col, row))
    col = col[order]
    val = np.round(rng.random(col.shape[0]) * 4 - 2).astype(np.float32) if with_value else None   # ties on purpose
    return rowptr, col, val


def worker(rank, world, port, with_value, balance, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isplib_b200.dist import DistSpMM
        from oracle import oracle
        M = N = 101            # not divisible by world: exercises the padding
        K = 6
        rowptr, col, val = make_graph(5, M, N, with_value)
        x = np.random.default_rng(1).integers(-3, 4, size=(N, K)).astype(np.float32)
        go = np.random.default_rng(2).standard_normal((M, K)).astype(np.float32)
        op = DistSpMM(torch.from_numpy(rowptr), torch.from_numpy(col), None if val is None else torch.from_numpy(val),
                      N, device="cpu", block_spmm=oracle_block_spmm, arg_backward=oracle_arg_backward, overlap=False,
                      balance=balance)
        f = op.fwd
        r0, r1 = f.row_range()
        c0, c1 = f.col_range()
        assert f.local.nnz + f.remote.nnz == int(rowptr[r1] - rowptr[r0])
        if balance == "rows":
            assert f.R == (M + world - 1) // world and (r0, r1) == (min(rank * f.R, M), min((rank + 1) * f.R, M))
        else:
            assert (c0, c1) == (r0, r1)          # square graph: X rows are owned like A rows
        ok = {}
        for reduce in ("sum", "mean", "max", "min"):
            xs = f.pad_x(torch.from_numpy(x[c0:c1])).requires_grad_(True)
            out = op(xs, reduce)
            gpad = torch.zeros((f.R, K))
            gpad[: r1 - r0] = torch.from_numpy(go[r0:r1])
            out.backward(gpad)
            code = oracle.REDUCE_CODE[reduce]
            ref, ref_arg = oracle.spmm_c(rowptr, col, val, x, code)
            got = out.detach().numpy()[: r1 - r0]
            if reduce in ("max", "min"):
                ok[reduce + "_fwd"] = bool(np.array_equal(got, ref[r0:r1]))
                gref, _ = oracle.arg_backward(col, val, None, ref_arg, go, N)
            else:
                ok[reduce + "_fwd"] = bool(np.allclose(got, ref[r0:r1], rtol=1e-5, atol=1e-5))
                bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
                gref = bw(rowptr, col, val, go, N)
            ok[reduce + "_bwd"] = bool(np.allclose(xs.grad.numpy()[: c1 - c0], gref[c0:c1], rtol=1e-4, atol=1e-4))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", ["nnz", "rows"])
@pytest.mark.parametrize("with_value", [True, False])
def test_row_partitioned_spmm_world2_gloo(with_value, balance):
    world = 2
    port = 29500 + (os.getpid() % 2000) + (1 if with_value else 0) + (2 if balance == "nnz" else 0)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(world, port, with_value, balance, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def epilogue_worker(rank, world, port, results):
    """relu?(A x + scale * addend + bias) through the row-partitioned operator: the epilogue rides in the LAST
    launch (after the ACCUMULATE merge and the mean division), gradients reach x, bias and addend."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isplib_b200.dist import DistSpMM
        from oracle import oracle
        M = N = 101
        K = 6
        rowptr, col, val = make_graph(9, M, N, True)
        x = np.random.default_rng(1).integers(-3, 4, size=(N, K)).astype(np.float32)
        bias = np.random.default_rng(3).standard_normal(K).astype(np.float32)
        go = np.random.default_rng(2).standard_normal((M, K)).astype(np.float32)
        ok = {}
        for k_chunk in (None, 4):
            op = DistSpMM(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), N, device="cpu",
                          block_spmm=oracle_block_spmm, arg_backward=oracle_arg_backward, overlap=False)
            op.fwd.k_chunk = k_chunk
            f = op.fwd
            r0, r1 = f.row_range()
            for reduce, relu, use_addend in (("sum", True, False), ("mean", False, True), ("sum", True, True), ("max", True, False)):
                xs = f.pad_x(torch.from_numpy(x[r0:r1])).requires_grad_(True)
                b = torch.from_numpy(bias.copy()).requires_grad_(True)
                scale = 1.25
                out = op(xs, reduce, bias=b, addend=xs if use_addend else None, addend_scale=scale, relu=relu)
                gpad = torch.zeros((f.R, K))
                gpad[: r1 - r0] = torch.from_numpy(go[r0:r1])
                out.backward(gpad)
                code = oracle.REDUCE_CODE[reduce]
                plain, ref_arg = oracle.spmm_c(rowptr, col, val, x, code)
                ref = oracle.apply_epilogue(plain, bias=bias, addend=x if use_addend else None, addend_scale=scale, relu=relu)
                tag = f"{reduce}_relu{int(relu)}_add{int(use_addend)}_chunk{k_chunk}"
                ok[tag + "_fwd"] = bool(np.allclose(out.detach().numpy()[: r1 - r0], ref[r0:r1], rtol=1e-5, atol=1e-5))
                g = go * (ref > 0) if relu else go
                if reduce == "max":
                    gx, _ = oracle.arg_backward(col, val, None, ref_arg, g, N)
                else:
                    bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
                    gx = bw(rowptr, col, val, g, N)
                if use_addend:
                    gx = gx + scale * g
                ok[tag + "_gx"] = bool(np.allclose(xs.grad.numpy()[: r1 - r0], gx[r0:r1], rtol=1e-4, atol=1e-4))
                ok[tag + "_gbias"] = bool(np.allclose(b.grad.numpy(), g[r0:r1].sum(0), rtol=1e-4, atol=1e-4))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_epilogue_rides_in_the_last_launch_world2_gloo():
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(epilogue_worker, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def local_ingest_worker(rank, world, port, with_value, results):
    """Every rank builds the operator from ITS OWN rows only (DistSpMM.from_local_rows) -- cut unevenly on
    purpose -- and must produce what the whole-graph constructor and the single-process oracle produce."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isplib_b200.dist import DistSpMM, distributed_transpose
        from oracle import oracle
        M, N, K = 101, 101, 6
        rowptr, col, val = make_graph(7, M, N, with_value)
        x = np.random.default_rng(1).integers(-3, 4, size=(N, K)).astype(np.float32)
        go = np.random.default_rng(2).standard_normal((M, K)).astype(np.float32)
        cuts = [0, 37, M] if world == 2 else [0, 20, 75, M]
        r0, r1 = cuts[rank], cuts[rank + 1]
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        rp_loc = torch.from_numpy(rowptr[r0:r1 + 1] - e0)
        col_loc = torch.from_numpy(col[e0:e1].copy())
        val_loc = None if val is None else torch.from_numpy(val[e0:e1].copy())
        op = DistSpMM.from_local_rows(rp_loc, col_loc, val_loc, N, device="cpu", block_spmm=oracle_block_spmm,
                                      arg_backward=oracle_arg_backward, overlap=False)
        f = op.fwd
        ok = {"bounds": f.row_bounds == cuts and f.col_bounds == cuts and f.nnz == col.shape[0],
              "rowptr": bool(np.array_equal(op.rowptr.numpy(), rowptr))}
        # the distributed transpose = this rank's rows of the single-process transpose (same order: by source row)
        for mean in (False, True):
            colptr, csr2csc, row_t = oracle.build_csc(rowptr, col, N)
            w = np.ones(col.shape[0], np.float32) if val is None else val
            w = w[csr2csc]
            if mean:
                w = w / np.maximum(np.diff(rowptr), 1)[row_t].astype(np.float32)
            cp, rt, vt, p0 = distributed_transpose(op.rowptr, col_loc, val_loc, r0, r1, N, f.col_bounds, mean)
            a, b = int(colptr[cuts[rank]]), int(colptr[cuts[rank + 1]])
            good = np.array_equal(cp.numpy(), colptr) and p0 == a and np.array_equal(rt.numpy(), row_t[a:b])
            if vt is not None:
                good = good and np.allclose(vt.numpy(), w[a:b], rtol=1e-6)
            else:
                good = good and val is None and not mean
            ok[f"transpose_mean{int(mean)}"] = bool(good)
        for reduce in ("sum", "mean", "max", "min"):
            xs = f.pad_x(torch.from_numpy(x[r0:r1])).requires_grad_(True)
            out = op(xs, reduce)
            gpad = torch.zeros((f.R, K))
            gpad[: r1 - r0] = torch.from_numpy(go[r0:r1])
            out.backward(gpad)
            code = oracle.REDUCE_CODE[reduce]
            ref, ref_arg = oracle.spmm_c(rowptr, col, val, x, code)
            got = out.detach().numpy()[: r1 - r0]
            if reduce in ("max", "min"):
                ok[reduce + "_fwd"] = bool(np.array_equal(got, ref[r0:r1]))
                _, a = f.forward(f.pad_x(torch.from_numpy(x[r0:r1])), reduce)
                ok[reduce + "_arg"] = bool(np.array_equal(a.numpy()[: r1 - r0], ref_arg[r0:r1]))     # GLOBAL edge ids
                gref, _ = oracle.arg_backward(col, val, None, ref_arg, go, N)
            else:
                ok[reduce + "_fwd"] = bool(np.allclose(got, ref[r0:r1], rtol=1e-5, atol=1e-5))
                bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
                gref = bw(rowptr, col, val, go, N)
            ok[reduce + "_bwd"] = bool(np.allclose(xs.grad.numpy()[: r1 - r0], gref[r0:r1], rtol=1e-4, atol=1e-4))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("with_value", [True, False])
def test_rank_local_ingest_matches_whole_graph_gloo(with_value, world):
    port = 31500 + (os.getpid() % 2000) + (1 if with_value else 0) + 2 * world
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(local_ingest_worker, args=(world, port, with_value, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_offset_array_only_serves_the_resident_range():
    from isplib_b200.dist import _OffsetArray
    a = _OffsetArray(torch.arange(10, 20), 100, 500)
    assert a.numel() == 500 and torch.equal(a[103:106], torch.tensor([13, 14, 15]))
    with pytest.raises(AssertionError):
        a[0:5]                      # outside the rank's own edge range: not resident


def test_split_row_block_partitions_every_entry_once():
    from isplib_b200.dist import split_row_block
    rowptr, col, val = make_graph(9, 57, 44, True)
    rp, co, va = torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val)
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            loc, rem, deg, R = split_row_block(rp, co, va, rank, world, 44)
            assert loc.rowptr.numel() == R + 1 and rem.rowptr.numel() == R + 1
            Rc = (44 + world - 1) // world
            assert loc.col.numel() == 0 or int(loc.col.max()) < Rc
            assert not bool(((rem.col >= rank * Rc) & (rem.col < (rank + 1) * Rc)).any())
            # edge ids are increasing inside each block row (the tie-break relies on it)
            for b in (loc, rem):
                e = b.edge_ids.numpy().astype(np.int64)
                for i in range(R):
                    seg = e[int(b.rowptr[i]):int(b.rowptr[i + 1])]
                    assert (np.diff(seg) > 0).all()
            seen += loc.edge_ids.tolist() + rem.edge_ids.tolist()
        assert sorted(seen) == list(range(col.shape[0]))


def test_nnz_balanced_bounds_even_out_a_skewed_graph():
    """contiguous row ranges balanced by stored entries, not by rows."""
    from isplib_b200.dist import nnz_balanced_bounds, even_bounds, split_row_block, slice_position, bounds_width
    M = 400
    deg = np.concatenate([np.full(40, 200), np.full(360, 3)])        # ids sorted by degree: heavy head
    rowptr = np.zeros(M + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    rng = np.random.default_rng(3)
    col = np.concatenate([np.sort(rng.choice(M, size=d, replace=False)) for d in deg]).astype(np.int64)
    rp, co = torch.from_numpy(rowptr), torch.from_numpy(col)
    nnz = int(rowptr[-1])
    for world in (2, 4, 8):
        nb, eb = nnz_balanced_bounds(rp, world), even_bounds(M, world)
        assert nb[0] == 0 and nb[-1] == M and all(nb[i] <= nb[i + 1] for i in range(world))
        load = lambda b: max(int(rowptr[b[p + 1]] - rowptr[b[p]]) for p in range(world))
        assert load(nb) <= 1.15 * nnz / world + 200          # within one heavy row of the ideal share
        assert load(nb) < load(eb)
        seen = []
        Rc = bounds_width(nb)
        for rank in range(world):
            loc, rem, _, R = split_row_block(rp, co, None, rank, world, M, nb, nb)
            assert R == bounds_width(nb)
            assert loc.col.numel() == 0 or int(loc.col.max()) < nb[rank + 1] - nb[rank]
            # remote columns point into the width-padded gathered layout, never into the own slice
            assert not bool(((rem.col >= rank * Rc) & (rem.col < (rank + 1) * Rc)).any())
            seen += loc.edge_ids.tolist() + rem.edge_ids.tolist()
        assert sorted(seen) == list(range(nnz))
        # the layout map is a bijection onto [owner * width, owner * width + size)
        pos = slice_position(torch.arange(M), nb, Rc)
        assert pos.unique().numel() == M and int(pos.max()) < world * Rc
    assert nnz_balanced_bounds(torch.zeros(6, dtype=torch.int64), 4) == even_bounds(5, 4)   # empty graph


# ----------------------------------------------------------------------------------------
# the drop-in surface of the multi-GPU mode: patch_pyg(group=...) + a partitioned adjacency
# ----------------------------------------------------------------------------------------
def plugin_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import isplib_b200  # noqa: F401
        import torch_sparse
        from isplib import iSpLibPlugin
        from isplib_b200 import nn as gnn
        from oracle import oracle
        M = N = 77
        K = 5
        rowptr, col, val = make_graph(11, M, N, True)
        x = np.random.default_rng(1).standard_normal((N, K)).astype(np.float32)
        adj = torch_sparse.SparseTensor(rowptr=torch.from_numpy(rowptr), col=torch.from_numpy(col),
                                        value=torch.from_numpy(val), sparse_sizes=(M, N), is_sorted=True)
        iSpLibPlugin.patch_pyg(group=dist.group.WORLD)
        try:
            padj = iSpLibPlugin.partition(adj, device="cpu", block_spmm=oracle_block_spmm,
                                          arg_backward=oracle_arg_backward, overlap=False, mode="nccl")
            r0, r1 = padj.row_range()
            xs = padj.local_slice(torch.from_numpy(x))
            ok = {}
            # the SAME call a single-process script makes: torch_sparse.matmul(adj_t, x, reduce)
            for reduce in ("sum", "mean", "max"):
                out = torch_sparse.matmul(padj, xs, reduce)
                ref = oracle.spmm_c(rowptr, col, val, x, oracle.REDUCE_CODE[reduce])[0]
                ok[reduce] = bool(np.allclose(out.numpy()[: r1 - r0], ref[r0:r1], rtol=1e-5, atol=1e-5))
            # SAGE / GIN drop the values: set_value(None) must hand back a partitioned twin
            nov = padj.set_value(None)
            out = torch_sparse.matmul(nov, xs, "sum")
            ref = oracle.spmm_c(rowptr, col, None, x, oracle.SUM)[0]
            ok["novalue"] = bool(np.allclose(out.numpy()[: r1 - r0], ref[r0:r1], rtol=1e-5, atol=1e-5)) and not nov.has_value()
            # a whole layer through the unchanged model code: GCNConv.forward calls matmul(adj_t, ...)
            torch.manual_seed(0)
            conv = gnn.GCNConv(K, 3, order="linear_first")
            y = conv(xs, padj)
            with torch.no_grad():
                h = (torch.from_numpy(x) @ conv.lin.weight.t()).numpy()
            ref = oracle.spmm_c(rowptr, col, val, h, oracle.SUM)[0] + conv.bias.detach().numpy()
            ok["gcn_layer"] = bool(np.allclose(y.detach().numpy()[: r1 - r0], ref[r0:r1], rtol=1e-4, atol=1e-4))
            # rank-local ingest: a SparseTensor holding ONLY this rank's rows gives the same partitioned operator
            e0, e1 = int(rowptr[r0]), int(rowptr[r1])
            mine = torch_sparse.SparseTensor(rowptr=torch.from_numpy(rowptr[r0:r1 + 1] - e0), col=torch.from_numpy(col[e0:e1].copy()),
                                             value=torch.from_numpy(val[e0:e1].copy()), sparse_sizes=(r1 - r0, N), is_sorted=True)
            ladj = iSpLibPlugin.partition(mine, device="cpu", local_rows=True, block_spmm=oracle_block_spmm,
                                          arg_backward=oracle_arg_backward, overlap=False, mode="nccl")
            ok["local_rows_range"] = ladj.row_range() == (r0, r1)
            for reduce in ("sum", "max"):
                a = torch_sparse.matmul(ladj, xs, reduce)
                b = torch_sparse.matmul(padj, xs, reduce)
                ok["local_rows_" + reduce] = bool(torch.equal(a, b))
            ok["local_rows_novalue"] = bool(torch.equal(torch_sparse.matmul(ladj.set_value(None), xs, "sum"),
                                                        torch_sparse.matmul(nov, xs, "sum")))
            results[rank] = ok
        finally:
            iSpLibPlugin.unpatch_pyg()
    finally:
        dist.destroy_process_group()


def test_patched_matmul_dispatches_partitioned_adjacency_world2_gloo():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(plugin_worker, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_owner_groups_and_runs_of_the_fused_gather():
    """Host logic of the fused gather's owner mode: arrival groups by ring distance (own slice = group 0,
    every group non-empty, sizes within one of each other), and the contiguous owner runs handed to
    isplib_b200_plan_build_grouped (run starts ascending from column 0, adjacent runs differ in group)."""
    from isplib_b200.dist import owner_groups, group_runs
    for world in (1, 2, 3, 4, 8, 16):
        for G in (1, 2, 3, 7):
            for rank in range(world):
                groups, n_groups = owner_groups(world, rank, G)
                assert len(groups) == world and groups[rank] == 0
                assert n_groups == (min(G, world - 1) + 1 if world > 1 else 1)
                remote = [groups[(rank + d) % world] for d in range(1, world)]
                assert remote == sorted(remote)                       # nearer peers land earlier
                if world > 1:
                    sizes = [remote.count(g) for g in range(1, n_groups)]
                    assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
                starts, grp = group_runs(groups, 10)
                assert starts[0] == 0 and starts == sorted(starts) and len(starts) == len(grp)
                assert all(a != b for a, b in zip(grp, grp[1:]))
                # expanding the runs gives back the per-owner groups
                ends = starts[1:] + [world * 10]
                expanded = [g for s, e, g in zip(starts, ends, grp) for _ in range((e - s) // 10)]
                assert expanded == groups


# ------------------------------------------------------------------------------------------------------------
# partitioned ingest (isplib_b200/dist_io.py): byte-range Matrix Market shards -> row owners -> operator
# ------------------------------------------------------------------------------------------------------------
def numpy_csr_builder(row, col, val, m, n):
    """Stand-in for isplib_b200_coo_to_csr in the CPU tests: stable sort by (row, col)."""
    r, c = row.numpy(), col.numpy()
    order = np.lexsort((c, r))            # numpy's lexsort is stable
    rowptr = np.zeros(m + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.bincount(r, minlength=m))
    return (torch.from_numpy(rowptr), torch.from_numpy(c[order].copy()),
            None if val is None else torch.from_numpy(val.numpy()[order].copy()))


def write_test_mtx(path, kind):
    import scipy.io
    import scipy.sparse
    rng = np.random.default_rng(21)
    M = N = 83
    if kind == "symmetric":
        a = scipy.sparse.random(M, N, density=0.06, random_state=3, format="coo", dtype=np.float64)
        a = scipy.sparse.triu(a + a.T).tocoo()
        a.data = np.round(a.data * 8) / 8 + 0.125
        scipy.io.mmwrite(path, a, symmetry="symmetric", comment="lower half implied")
    else:
        deg = rng.integers(0, 14, size=M)
        deg[5] = 70                       # a heavy row: the nnz-balanced cut is not the even one
        row = np.repeat(np.arange(M), deg)
        col = np.concatenate([rng.choice(N, size=d, replace=False) for d in deg]) if row.size else row
        perm = rng.permutation(row.size)  # file order is NOT row order: entries really have to travel
        data = np.round(rng.random(row.size) * 8) / 8 + 0.125
        a = scipy.sparse.coo_matrix((data[perm], (row[perm], col[perm])), shape=(M, N))
        scipy.io.mmwrite(path, a, field="pattern" if kind == "pattern" else "real", symmetry="general",
                         comment="lines of\nvery different length follow")
    return M, N


def mtx_ingest_worker(rank, world, port, path, kind, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.io
        from isplib_b200 import dist_io
        from isplib_b200.dist import nnz_balanced_bounds
        from oracle import oracle
        coo = scipy.io.mmread(path).tocoo()
        M, N = coo.shape
        order = np.lexsort((coo.col, coo.row))
        g_row, g_col = coo.row[order].astype(np.int64), coo.col[order].astype(np.int64)
        g_val = None if kind == "pattern" else coo.data[order].astype(np.float32)
        rowptr = np.zeros(M + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(np.bincount(g_row, minlength=M))
        padj = dist_io.read_mtx_partitioned(path, device="cpu", csr_builder=numpy_csr_builder,
                                            block_spmm=oracle_block_spmm, arg_backward=oracle_arg_backward,
                                            overlap=False, mode="nccl")
        f = padj.op.fwd
        r0, r1 = padj.row_range()
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        ok = {"bounds": f.row_bounds == nnz_balanced_bounds(torch.from_numpy(rowptr), world),
              "rowptr": bool(np.array_equal(padj.op.rowptr.numpy(), rowptr)),
              "col": bool(np.array_equal(padj.op.col.local.numpy(), g_col[e0:e1])),
              "has_value": padj.has_value() == (kind != "pattern")}
        if g_val is not None:
            ok["val"] = bool(np.array_equal(padj.op.value.local.numpy(), g_val[e0:e1]))
        # every line was parsed by exactly one rank
        n_lines = torch.tensor([dist_io.read_mtx_shard(path, rank, world)[0].numel()])
        dist.all_reduce(n_lines)
        ok["every_entry_once"] = int(n_lines) == g_col.shape[0]
        K = 5
        x = np.random.default_rng(1).standard_normal((N, K)).astype(np.float32)
        xs = padj.local_slice(torch.from_numpy(x)).requires_grad_(True)
        for reduce in ("sum", "max"):
            out = padj.matmul(xs, reduce)
            ref = oracle.spmm_c(rowptr, g_col, g_val, x, oracle.REDUCE_CODE[reduce])[0]
            ok[reduce] = bool(np.allclose(out.detach().numpy()[: r1 - r0], ref[r0:r1], rtol=1e-5, atol=1e-5))
        go = np.random.default_rng(2).standard_normal((M, K)).astype(np.float32)
        gpad = torch.zeros((f.R, K))
        gpad[: r1 - r0] = torch.from_numpy(go[r0:r1])
        padj.matmul(xs, "sum").backward(gpad)
        gref = oracle.spmm_backward_sum(rowptr, g_col, g_val, go, N)
        ok["sum_bwd"] = bool(np.allclose(xs.grad.numpy()[: r1 - r0], gref[r0:r1], rtol=1e-4, atol=1e-4))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("real", 2), ("real", 3), ("pattern", 2), ("symmetric", 3)])
def test_matrix_market_byte_range_ingest_gloo(tmp_path, kind, world):
    path = str(tmp_path / f"g_{kind}.mtx")
    write_test_mtx(path, kind)
    port = 35500 + (os.getpid() % 2000) + 3 * world + {"real": 0, "pattern": 1, "symmetric": 2}[kind]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(mtx_ingest_worker, args=(world, port, path, kind, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_mtx_shards_cover_every_line_once_for_any_world(tmp_path):
    from isplib_b200 import dist_io
    path = str(tmp_path / "g.mtx")
    write_test_mtx(path, "real")
    m, n, nnz, field, symmetry, _ = dist_io.mtx_header(path)
    assert (m, n, field, symmetry) == (83, 83, "real", "general")
    whole = dist_io.read_mtx_shard(path, 0, 1)
    assert whole[0].numel() == nnz
    for world in (2, 5, 16, 200, 5000):      # more ranks than lines: some shards are empty, none overlaps
        parts = [dist_io.read_mtx_shard(path, r, world) for r in range(world)]
        assert torch.equal(torch.cat([p[0] for p in parts]), whole[0])
        assert torch.equal(torch.cat([p[1] for p in parts]), whole[1])
        assert torch.equal(torch.cat([p[2] for p in parts]), whole[2])
    with pytest.raises(ValueError):
        bad = tmp_path / "dense.mtx"
        bad.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
        dist_io.mtx_header(str(bad))


def test_partitioned_forward_refuses_malformed_slices():
    """A wrong slice shape would make the block kernels read outside the gathered matrix: refuse it up front."""
    from isplib_b200.dist import RowPartitionedSpMM, make_epilogue
    rowptr, col, val = make_graph(3, 40, 40, True)
    op = RowPartitionedSpMM(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), 40, device="cpu",
                            block_spmm=oracle_block_spmm, overlap=False, emulate=(2, 0, {}), mode="nccl")
    good = torch.zeros(op.Rc, 4)
    with pytest.raises(ValueError, match="padded fp32 slice"):
        op.forward(torch.zeros(op.Rc - 1, 4))
    with pytest.raises(ValueError, match="padded fp32 slice"):
        op.forward(good.double())
    with pytest.raises(ValueError, match="reduce must be"):
        op.forward(good, "prod")
    with pytest.raises(ValueError, match="bias must have shape"):
        op.forward(good, "sum", epilogue=make_epilogue(bias=torch.zeros(5)))
    with pytest.raises(ValueError, match="addend must have shape"):
        op.forward(good, "sum", epilogue=make_epilogue(addend=torch.zeros(op.R + 1, 4)))


def bad_shard_worker(rank, world, port, path, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isplib_b200 import dist_io
        seen = {}
        # (1) an endpoint outside the matrix in ONE rank's edge shard
        row = torch.tensor([0, 1, 2]) if rank == 0 else torch.tensor([1, 99])
        col = torch.tensor([1, 2, 0]) if rank == 0 else torch.tensor([0, 1])
        try:
            dist_io.partition_edges(row, col, None, 4, 4, device="cpu", csr_builder=numpy_csr_builder)
            seen["edges"] = "no error"
        except ValueError as e:
            seen["edges"] = "own: " + str(e)
        except RuntimeError as e:
            seen["edges"] = "peer: " + str(e)
        # (2) a malformed line in ONE rank's byte range of the file
        try:
            dist_io.read_mtx_partitioned(path, device="cpu", csr_builder=numpy_csr_builder)
            seen["mtx"] = "no error"
        except RuntimeError as e:
            seen["mtx"] = "peer: " + str(e)
        except Exception as e:
            seen["mtx"] = "own: " + type(e).__name__
        # the group is still usable afterwards: nobody is stuck in a half-entered collective
        t = torch.ones(1)
        dist.all_reduce(t)
        seen["alive"] = int(t)
        results[rank] = seen
    finally:
        dist.destroy_process_group()


def test_a_bad_shard_on_one_rank_stops_every_rank_gloo(tmp_path):
    """Validation errors are agreed on by the whole group before the first data collective: the rank with
    the bad shard raises its own error, the others a RuntimeError naming a peer -- none is left waiting."""
    path = tmp_path / "broken.mtx"
    lines = ["%%MatrixMarket matrix coordinate real general", "6 6 8"]
    lines += [f"{i + 1} {(i * 5) % 6 + 1} 0.5" for i in range(7)] + ["6 oops 0.25"]     # the last line is rank 1's
    path.write_text("\n".join(lines) + "\n")
    world, port = 2, 36900 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(bad_shard_worker, args=(world, port, str(path), results), nprocs=world, join=True)
    assert len(results) == world
    assert results[1]["edges"].startswith("own: ") and "outside" in results[1]["edges"]
    assert results[0]["edges"].startswith("peer: ")
    assert results[1]["mtx"].startswith("own: ")
    assert results[0]["mtx"].startswith("peer: ")
    assert results[0]["alive"] == results[1]["alive"] == world


def test_epilogue_addend_of_any_row_count_single_rank():
    """The addend may be shorter than the padded output slice (zero padded), exactly R rows, or longer (cut):
    forward and the gradient w.r.t. the addend keep the addend's own shape."""
    from isplib_b200.dist import DistSpMM
    from oracle import oracle
    M = N = 37
    K = 5
    rowptr, col, val = make_graph(11, M, N, True)
    op = DistSpMM(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), N, device="cpu",
                  block_spmm=oracle_block_spmm, arg_backward=oracle_arg_backward, overlap=False, emulate=(1, 0, {}),
                  mode="nccl")
    R = op.fwd.R
    assert R == M
    x = np.random.default_rng(3).integers(-3, 4, size=(N, K)).astype(np.float32)
    ref = oracle.spmm_c(rowptr, col, val, x, oracle.SUM)[0]
    go = torch.from_numpy(np.random.default_rng(4).standard_normal((R, K)).astype(np.float32))
    for rows in (R - 7, R, R + 9):
        add = torch.from_numpy(np.random.default_rng(rows).standard_normal((rows, K)).astype(np.float32)).requires_grad_(True)
        xs = torch.from_numpy(x).requires_grad_(True)
        out = op(xs, "sum", addend=add, addend_scale=0.5)
        want = ref.copy()
        n = min(rows, R)
        want[:n] += 0.5 * add.detach().numpy()[:n]
        np.testing.assert_allclose(out.detach().numpy(), want, rtol=1e-5, atol=1e-5)
        out.backward(go)
        assert add.grad.shape == add.shape
        np.testing.assert_allclose(add.grad.numpy()[:n], 0.5 * go.numpy()[:n], rtol=1e-6, atol=1e-6)
        assert not add.grad.numpy()[n:].any()
