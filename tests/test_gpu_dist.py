"""2-GPU tests of the row-partitioned path with the real CUDA kernels (skipped on boxes
with fewer than 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`).
Forward and backward of every reduction against the single-process oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, results, mode="fused"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from isplib_b200 import synth
        from isplib_b200.dist import DistSpMM
        from oracle import oracle
        g = synth.make_graph(4001, 300_000, law="lognormal", param=1.3, values="uniform", seed=3)
        val = torch.round(g.value * 4) / 4            # coarse values: exact ties for max/min
        K = 32
        x = torch.randint(-3, 4, (g.n, K), generator=torch.Generator().manual_seed(1)).float()
        go = torch.randn(g.m, K, generator=torch.Generator().manual_seed(2))
        pipelined = mode == "pipelined"
        op = DistSpMM(g.rowptr.to(dev), g.col.to(dev), val.to(dev), g.n, device=dev, pipelined=pipelined,
                      mode=mode if mode in ("fused", "auto") else "nccl")
        assert op.fwd.pipelined == pipelined and op.fwd.mode == (mode if mode in ("fused", "auto") else "nccl")
        f = op.fwd
        r0, r1 = f.row_range()
        c0, c1 = f.col_range()
        rp, co, va = g.rowptr.numpy(), g.col.numpy(), val.numpy()
        ok = {}
        for reduce in ("sum", "mean", "max", "min"):
            xs = f.pad_x(x[c0:c1].to(dev)).requires_grad_(True)
            out = op(xs, reduce)
            gpad = torch.zeros((f.R, K), device=dev)
            gpad[: r1 - r0] = go[r0:r1].to(dev)
            out.backward(gpad)
            torch.cuda.synchronize()
            code = oracle.REDUCE_CODE[reduce]
            ref, ref_arg = oracle.spmm_c(rp, co, va, x.numpy(), code)
            got = out.detach().cpu().numpy()[: r1 - r0]
            if reduce in ("max", "min"):
                ok[reduce + "_fwd"] = bool(np.array_equal(got, ref[r0:r1]))
                gref, _ = oracle.arg_backward(co, va, None, ref_arg, go.numpy(), g.n)
            else:
                ok[reduce + "_fwd"] = bool(np.allclose(got, ref[r0:r1], rtol=1e-4, atol=1e-4))
                bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
                gref = bw(rp, co, va, go.numpy(), g.n)
            ok[reduce + "_bwd"] = bool(np.allclose(xs.grad.cpu().numpy()[: c1 - c0], gref[c0:c1], rtol=1e-3, atol=1e-3))
            if reduce in ("max", "min"):
                # arg_out carries GLOBAL edge ids, bit-identical to the single-GPU answer
                _, a = f.forward(f.pad_x(x[c0:c1].to(dev)), reduce)
                ok[reduce + "_arg"] = bool(np.array_equal(a.cpu().numpy()[: r1 - r0], ref_arg[r0:r1]))
        f.check_status()
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_row_partitioned_spmm_two_gpus_fused_gather():
    """The default multi-GPU path: ONE kernel pulls the peer's slice of X over NVLink (symmetric
    memory) and multiplies; no collective call in the forward or in the sum/mean backward."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(2, 29350 + os.getpid() % 300, results, "fused"), nprocs=2, join=True)
    for rank in range(2):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_row_partitioned_spmm_two_gpus_auto_mode():
    """mode='auto' (the default): both paths are built and the faster one is measured per width."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(2, 29450 + os.getpid() % 300, results, "auto"), nprocs=2, join=True)
    for rank in range(2):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_row_partitioned_spmm_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(2, 29650 + os.getpid() % 300, results, "nccl"), nprocs=2, join=True)
    for rank in range(2):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def test_row_partitioned_spmm_two_gpus_pipelined_p2p():
    """Same, with the per-source pipelined gather: X slices move peer-to-peer over NVLink by the
    copy engines (torch symmetric memory) and each column-owner block is multiplied as it lands."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(2, 29950 + os.getpid() % 300, results, "pipelined"), nprocs=2, join=True)
    for rank in range(2):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def local_ingest_worker(rank, world, port, results, mode):
    """Rank-local ingest on the real kernels: every rank hands DistSpMM.from_local_rows ITS rows only."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from isplib_b200 import synth
        from isplib_b200.dist import DistSpMM
        from oracle import oracle
        g = synth.make_graph(4001, 300_000, law="lognormal", param=1.3, values="uniform", seed=3)
        val = torch.round(g.value * 4) / 4
        K = 32
        x = torch.randint(-3, 4, (g.n, K), generator=torch.Generator().manual_seed(1)).float()
        go = torch.randn(g.m, K, generator=torch.Generator().manual_seed(2))
        cuts = [0, g.m] if world == 1 else [0, 1700, g.m]          # uneven on purpose
        r0, r1 = cuts[rank], cuts[rank + 1]
        e0, e1 = int(g.rowptr[r0]), int(g.rowptr[r1])
        op = DistSpMM.from_local_rows((g.rowptr[r0:r1 + 1] - e0).to(dev), g.col[e0:e1].to(dev), val[e0:e1].to(dev),
                                      g.n, device=dev, mode=mode)
        f = op.fwd
        assert f.row_bounds == cuts and f.row_range() == (r0, r1)
        rp, co, va = g.rowptr.numpy(), g.col.numpy(), val.numpy()
        ok = {}
        for reduce in ("sum", "mean", "max", "min"):
            xs = f.pad_x(x[r0:r1].to(dev)).requires_grad_(True)
            out = op(xs, reduce)
            gpad = torch.zeros((f.R, K), device=dev)
            gpad[: r1 - r0] = go[r0:r1].to(dev)
            out.backward(gpad)
            torch.cuda.synchronize()
            ref, ref_arg = oracle.spmm_c(rp, co, va, x.numpy(), oracle.REDUCE_CODE[reduce])
            got = out.detach().cpu().numpy()[: r1 - r0]
            if reduce in ("max", "min"):
                ok[reduce + "_fwd"] = bool(np.array_equal(got, ref[r0:r1]))
                gref, _ = oracle.arg_backward(co, va, None, ref_arg, go.numpy(), g.n)
            else:
                ok[reduce + "_fwd"] = bool(np.allclose(got, ref[r0:r1], rtol=1e-4, atol=1e-4))
                bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
                gref = bw(rp, co, va, go.numpy(), g.n)
            ok[reduce + "_bwd"] = bool(np.allclose(xs.grad.cpu().numpy()[: r1 - r0], gref[r0:r1], rtol=1e-3, atol=1e-3))
        f.check_status()
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(1, "nccl"), (2, "nccl"), (2, "fused")])
def test_rank_local_ingest_on_the_gpu(world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(local_ingest_worker, args=(world, 30250 + os.getpid() % 300 + 7 * world, results, mode), nprocs=world, join=True)
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def epilogue_worker(rank, world, port, results, mode):
    """relu(A x + bias) through DistSpMM on real ranks, forward and gradients, and a GCN layer through the
    patched matmul with a partitioned adjacency (the fused path of nn.GCNConv)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import isplib_b200  # noqa: F401
        from isplib import iSpLibPlugin
        from isplib_b200 import nn as gnn, synth
        from isplib_b200.dist import PartitionedAdj
        from oracle import oracle
        g = synth.make_graph(4001, 300_000, law="lognormal", param=1.3, values="uniform", seed=3)
        K = 128 if mode == "fused" else 32
        x = torch.randn(g.n, K, generator=torch.Generator().manual_seed(1))
        bias = torch.randn(K, generator=torch.Generator().manual_seed(4))
        go = torch.randn(g.m, K, generator=torch.Generator().manual_seed(2))
        rp, co, va = g.rowptr.numpy(), g.col.numpy(), g.value.numpy()
        iSpLibPlugin.patch_pyg(group=dist.group.WORLD)
        try:
            padj = PartitionedAdj(g.rowptr.to(dev), g.col.to(dev), g.value.to(dev), g.n, device=dev, mode=mode)
            r0, r1 = padj.row_range()
            ok = {}
            xs = padj.local_slice(x).requires_grad_(True)
            b = bias.to(dev).requires_grad_(True)
            out = isplib_b200.fused_matmul(padj, xs, "sum", bias=b, relu=True)
            gpad = torch.zeros_like(out)
            gpad[: r1 - r0] = go[r0:r1].to(dev)
            out.backward(gpad)
            torch.cuda.synchronize()
            plain, _ = oracle.spmm_c(rp, co, va, x.numpy(), oracle.SUM)
            ref = oracle.apply_epilogue(plain, bias=bias.numpy(), relu=True)
            ok["fwd"] = bool(np.allclose(out.detach().cpu().numpy()[: r1 - r0], ref[r0:r1], rtol=1e-4, atol=1e-4))
            gm = go.numpy() * (ref > 0)
            ok["gx"] = bool(np.allclose(xs.grad.cpu().numpy()[: r1 - r0], oracle.spmm_backward_sum(rp, co, va, gm, g.n)[r0:r1],
                                        rtol=1e-3, atol=1e-3))
            ok["gbias"] = bool(np.allclose(b.grad.cpu().numpy(), gm[r0:r1].sum(0), rtol=1e-3, atol=1e-3))
            # the model code, unchanged: GCNConv -> fused_matmul -> the partitioned operator
            torch.manual_seed(0)
            conv = gnn.GCNConv(K, 16, order="linear_first", relu=True).to(dev)
            with torch.no_grad():
                conv.bias.normal_()
            y = conv(xs.detach(), padj)
            with torch.no_grad():
                h = (x.to(dev) @ conv.lin.weight.t()).cpu().numpy()
            ref = oracle.apply_epilogue(oracle.spmm_c(rp, co, va, h, oracle.SUM)[0], bias=conv.bias.detach().cpu().numpy(), relu=True)
            ok["gcn_layer"] = bool(np.allclose(y.detach().cpu().numpy()[: r1 - r0], ref[r0:r1], rtol=1e-3, atol=1e-3))
            padj.op.fwd.check_status()
            results[rank] = ok
        finally:
            iSpLibPlugin.unpatch_pyg()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fused", "nccl"])
def test_epilogue_through_the_partitioned_operator_two_gpus(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(epilogue_worker, args=(2, 30650 + os.getpid() % 300, results, mode), nprocs=2, join=True)
    for rank in range(2):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"


def mtx_ingest_worker(rank, world, port, path, results):
    """Partitioned ingest with the real device CSR build (isplib_b200_coo_to_csr per rank) against io.read_mtx."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from isplib_b200 import dist_io, io as gio
        whole = gio.read_mtx(path, device=dev)
        rowptr, col, val = whole.csr()
        padj = dist_io.read_mtx_partitioned(path, device=dev, mode="nccl")
        r0, r1 = padj.row_range()
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        ok = {"rowptr": bool(torch.equal(padj.op.rowptr, rowptr.to(torch.int64))),
              "col": bool(torch.equal(padj.op.col.local, col[e0:e1].to(torch.int64))),
              "val": bool(torch.equal(padj.op.value.local, val[e0:e1]))}
        x = torch.randn(whole.sparse_sizes()[1], 32, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        import isplib_b200
        from isplib import iSpLibPlugin
        iSpLibPlugin.patch_pyg(group=dist.group.WORLD)
        try:
            import torch_sparse
            for reduce in ("sum", "max"):
                a = torch_sparse.matmul(padj, padj.local_slice(x), reduce)[: r1 - r0]
                b = torch_sparse.matmul(whole, x, reduce)[r0:r1]
                ok[reduce] = bool(torch.allclose(a, b, rtol=1e-4, atol=1e-4))
        finally:
            iSpLibPlugin.unpatch_pyg()
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_matrix_market_partitioned_ingest_on_the_gpu(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import scipy.io
    import scipy.sparse
    a = scipy.sparse.random(3001, 3001, density=0.004, random_state=5, format="coo", dtype=np.float64)
    a.data = np.round(a.data * 16) / 16 + 0.0625
    path = str(tmp_path / "g.mtx")
    scipy.io.mmwrite(path, a)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(mtx_ingest_worker, args=(world, 30950 + os.getpid() % 300 + world, path, results), nprocs=world, join=True)
    for rank in range(world):
        bad = [k for k, v in results[rank].items() if not v]
        assert not bad, f"rank {rank}: {bad}"
