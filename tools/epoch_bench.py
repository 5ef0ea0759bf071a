#!/usr/bin/env python
"""Full-batch GNN training epoch on a synthetic graph, patched through iSpLibPlugin.

Epoch definition = /root/reference/tests/cpu/gcn-sparse.py:82-93: zero_grad -> forward ->
nll_loss(train mask) -> backward -> Adam step -> a second forward for train accuracy.
Prints one JSON line: epoch ms with and without the accuracy forward, and the share of the
epoch spent in isplib SpMM calls (forward + backward).

    python tools/epoch_bench.py --model gcn --shape products --feat 100 --hidden 256 --classes 47
    torchrun --nproc-per-node 2 tools/epoch_bench.py ...        (row-partitioned)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def run(a, init_dist=True):
    """a: namespace with model, shape, feat, hidden, classes, epochs, warmup, scale, stock.
    Returns the result dict on rank 0 (None elsewhere)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1 and init_dist and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    import isplib_b200  # noqa: F401
    from isplib import iSpLibPlugin
    from isplib_b200 import nn as gnn, synth

    torch.manual_seed(0)
    values = "gcn" if a.model == "gcn" else None
    g = synth.make_graph(a.shape, values=values, seed=0, device=dev, scale=a.scale)
    N = g.n
    gen = torch.Generator(device=dev).manual_seed(0)
    x_all = torch.randn(N, a.feat, device=dev, generator=gen)
    y_all = torch.randint(0, a.classes, (N,), device=dev, generator=gen)
    train_all = torch.rand(N, device=dev, generator=gen) < 0.5

    if a.model == "gcn":
        model = gnn.GCN(a.feat, a.hidden, a.classes, order=getattr(a, "order", "linear_first"))
    elif a.model.startswith("sage"):
        model = gnn.GraphSAGE(a.feat, a.hidden, a.classes, aggr=a.model.split("-")[1])
    else:
        model = gnn.GIN(a.feat, a.hidden, a.classes)
    model = model.to(dev)
    use_graph = bool(getattr(a, "cuda_graph", False)) and world == 1
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4, capturable=use_graph)

    if world == 1:
        adj = g.sparse_tensor()
        spmm = None
        x, y, train = x_all, y_all, train_all
        n_train = int(train.sum())
    else:
        # the drop-in multi-GPU path: the model still calls torch_sparse.matmul(adj_t, x, reduce); the
        # patched matmul recognises the partitioned adjacency and runs the row-partitioned operator
        iSpLibPlugin.dist_group = None
        adj = iSpLibPlugin.partition(g.sparse_tensor(), device=dev)
        op = adj.op
        f = op.fwd
        spmm = None
        r0, r1 = f.col_range()
        x = f.pad_x(x_all[r0:r1])
        y = torch.zeros(f.Rc, dtype=torch.long, device=dev)
        y[: r1 - r0] = y_all[r0:r1]
        train = torch.zeros(f.Rc, dtype=torch.bool, device=dev)
        train[: r1 - r0] = train_all[r0:r1]
        n_train = int(train_all.sum())
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    del x_all

    if not a.stock:
        iSpLibPlugin.patch_pyg()

    train_idx = train.nonzero(as_tuple=True)[0]     # index form: no host sync inside the step
    y_train = y[train_idx]

    def loss_fn(out):
        lp = out if a.model != "gin" else F.log_softmax(out, dim=1)
        return F.nll_loss(lp.index_select(0, train_idx), y_train, reduction="sum") / n_train

    def epoch(with_acc: bool):
        model.train()
        opt.zero_grad()
        out = model(x, adj, spmm)
        loss = loss_fn(out)
        loss.backward()
        if world > 1:
            for p in model.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
        opt.step()
        acc = None
        if with_acc:                                       # gcn-sparse.py:88-90
            pred = model(x, adj, spmm).max(dim=1)[1]
            acc = pred[train].eq(y[train]).sum()
        return loss, acc

    def timed(with_acc, n):
        for _ in range(a.warmup):
            epoch(with_acc)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss, acc = epoch(with_acc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        loss = loss.detach().clone()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.all_reduce(loss)        # local partial sums / global count -> global mean
        return ms, float(loss)

    graph_ms = None
    if use_graph:
        # Launch-bound small graphs: capture one whole training step (forward, loss, backward,
        # Adam) in a CUDA graph and replay it.  The ops are capture-safe after their first call
        # per graph (plan build / variant selection happen during the warm-up below).
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                epoch(False)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss, _ = epoch(False)
        for _ in range(a.warmup):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.epochs):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / a.epochs

    ms_full, loss = timed(True, a.epochs)
    ms_train, _ = timed(False, a.epochs)
    if not a.stock:
        iSpLibPlugin.unpatch_pyg()
    res = None
    if rank == 0:
        res = {"model": a.model, "shape": a.shape, "nodes": g.m, "nnz": g.nnz, "feat": a.feat,
               "hidden": a.hidden, "classes": a.classes, "n_gpus": world,
               "mode": "stock-torch" if a.stock else "isplib_b200", "gcn_order": getattr(a, "order", "linear_first"),
               "epoch_ms_with_accuracy_forward": round(ms_full, 3),
               "epoch_ms_train_only": round(ms_train, 3), "final_loss": round(loss, 5),
               "epochs_timed": a.epochs}
        if graph_ms is not None:
            res["epoch_ms_train_only_cuda_graph"] = round(graph_ms, 4)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gcn", choices=["gcn", "sage-mean", "sage-sum", "gin"])
    ap.add_argument("--shape", default="products")
    ap.add_argument("--feat", type=int, default=100)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--stock", action="store_true", help="do not patch: stock torch-op matmul (the 'pt1' mode)")
    ap.add_argument("--order", default="linear_first", choices=["linear_first", "aggregate_first", "auto"],
                    help="GCN only: PyG's order (linear then propagate) or the cheaper equivalent")
    ap.add_argument("--cuda-graph", action="store_true", help="also time the training step replayed from a CUDA graph")
    a = ap.parse_args()
    res = run(a)
    if res is not None:
        print(json.dumps(res), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
