"""bench.py's reference arm runs on the host cores only, so its JSON contract can be checked
without a GPU: one line on stdout, the driver's keys, the tier's cpu_baseline / e2e objects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    # --no-gcn: the reference-arm GCN epoch (tens of seconds per epoch on a laptop-class host) is the GPU
    # box's business; the contract keys of the line do not depend on it
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--no-gcn"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                       # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "spmm_sum_effective_gbs" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "reddit-shape SpMM-sum forward, K=128" in d["config"]["workload"]
    assert d["config"]["full_graph"] == ("REDUCED" not in d["config"]["workload"])     # a shrunk sample must say so
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["unit"] == "GB/s" and cb["sample"]
    assert abs(cb["value"] - d["value"]) < 1e-6
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and abs(e2e["value"] - d["value"]) < 1e-6


def test_workload_name_is_shared_by_both_arms():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = bench.workload_name("reddit", "sum", 128)
    assert "232965 nodes" in w and "114615892 nnz" in w and "K=128" in w
