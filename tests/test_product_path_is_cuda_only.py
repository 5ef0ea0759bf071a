"""The oracle is test infrastructure: nothing the package ships may import, link or execute it, and the
product path has no CPU fallback -- it must fail loudly without the CUDA library / a CUDA tensor."""
import ast
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "isplib_b200")


def package_sources():
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                yield os.path.join(base, f)


def test_nothing_in_the_package_references_the_oracle():
    for path in package_sources():
        text = open(path, encoding="utf-8", errors="replace").read()
        if path.endswith(".py"):
            for node in ast.walk(ast.parse(text)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
            code = "\n".join(ln.split("#")[0] for ln in text.splitlines())
        else:
            code = re.sub(r"//.*|#.*", "", text)
        assert "oracle/" not in code and "fusedmm_oracle" not in code and "_ref/" not in code, path


def test_importing_the_package_does_not_load_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import isplib_b200, isplib_b200.dist, isplib_b200.dist_io, "
            "isplib_b200.nn, isplib_b200.io; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "import ctypes; maps = open('/proc/self/maps').read(); "
            "assert not bad, bad; assert 'fusedmm_oracle' not in maps and 'oracle/_ref' not in maps") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine WITHOUT a GPU")
def test_cpu_tensors_are_refused_not_computed():
    """No silent CPU path: the ops raise for host tensors (the reference's CPU arithmetic is not shipped)."""
    import isplib_b200  # noqa: F401
    rowptr = torch.tensor([0, 2, 3])
    col = torch.tensor([0, 1, 1])
    x = torch.ones(2, 4)
    with pytest.raises(Exception):
        torch.ops.isplib.fusedmm_spmm_max(rowptr, col, None, x)
    from isplib_b200 import dist_io
    with pytest.raises(RuntimeError, match="CUDA"):
        dist_io._cuda_csr_builder(torch.tensor([0, 1]), torch.tensor([1, 0]), None, 2, 2)
