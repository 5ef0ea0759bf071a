#!/usr/bin/env python
"""Per-call latency of the plugin path on a tiny graph (Cora-shape): where does the time go
between `torch_sparse.matmul(...)` and the 5 us kernel?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isplib import iSpLibPlugin
from isplib_b200 import capi, synth
import torch_sparse

dev = "cuda:0"
g = synth.make_graph("cora", values="gcn", seed=0).to(dev)
adj = g.sparse_tensor()
x = torch.randn(g.n, 64, device=dev)
xg = x.clone().requires_grad_(True)
rp, co = capi.narrow_i64_to_i32(g.rowptr), capi.narrow_i64_to_i32(g.col)
plan = capi.Plan(rp, g.nnz)
out = torch.empty(g.m, 64, device=dev)


def wall(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6


iSpLibPlugin.patch_pyg()
print(f"plugin matmul (no grad)        {wall(lambda: torch_sparse.matmul(adj, x, 'sum')):8.1f} us/call")
print(f"plugin matmul (requires_grad)  {wall(lambda: torch_sparse.matmul(adj, xg, 'sum')):8.1f} us/call")
iSpLibPlugin.unpatch_pyg()
ops = torch.ops.isplib
print(f"torch op direct                {wall(lambda: ops.fusedmm_spmm(None, g.rowptr, g.col, g.value, None, None, x, None, None)):8.1f} us/call")
print(f"C ABI via ctypes               {wall(lambda: capi.spmm_csr('sum', rp, co, g.value, x, plan, 1, out=out)):8.1f} us/call")
print(f"stock torch matmul (unpatched) {wall(lambda: torch_sparse.matmul(adj, x, 'sum'), 500):8.1f} us/call")
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(10):
        y = ops.fusedmm_spmm(None, g.rowptr, g.col, g.value, None, None, x, None, None)
print(f"CUDA graph replay (10 SpMMs)   {wall(graph.replay, 500) / 10:8.1f} us/SpMM")
