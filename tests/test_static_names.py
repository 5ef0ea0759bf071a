"""Static check that runs without a GPU: every global name a function of the package, the tools, the
bench or the GPU-only tests refers to is bound somewhere (module level, import, builtin).  The code paths
behind `torch.cuda.is_available()` never execute in the CPU suite, so a removed import there would first
show on the GPU box."""
import builtins
import os
import subprocess
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__spec__"}


def undefined_globals(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}

    def declared_global(tab):
        for s in tab.get_symbols():
            if s.is_declared_global() and s.is_assigned():
                module_names.add(s.get_name())
        for c in tab.get_children():
            declared_global(c)

    declared_global(top)
    missing = []

    def walk(tab):
        for s in tab.get_symbols():
            if s.is_referenced() and s.is_global() and not s.is_assigned() and not s.is_imported():
                if s.get_name() not in module_names and s.get_name() not in BUILTINS:
                    missing.append(f"{os.path.relpath(path, ROOT)}:{tab.get_lineno()}: {tab.get_name()} -> {s.get_name()}")
        for c in tab.get_children():
            walk(c)

    walk(top)
    return missing


def test_the_checker_sees_a_missing_import(tmp_path):
    p = tmp_path / "m.py"
    p.write_text("import os\ndef f():\n    import json\n    return json\ndef g():\n    return json.dumps(1), os\n")
    assert undefined_globals(str(p)) == [f"{os.path.relpath(str(p), ROOT)}:5: g -> json"]


def test_no_function_refers_to_an_unbound_global():
    files = subprocess.run(["git", "ls-files", "*.py"], cwd=ROOT, capture_output=True, text=True).stdout.split()
    if not files:                                       # a snapshot without .git (the GPU box)
        files = [os.path.relpath(os.path.join(d, f), ROOT) for d, _, fs in os.walk(ROOT) for f in fs
                 if f.endswith(".py") and "/baseline/" not in d + "/" and "/gpurun_out" not in d]
    assert files
    missing = [m for f in files for m in undefined_globals(os.path.join(ROOT, f))]
    assert not missing, "\n".join(missing)
