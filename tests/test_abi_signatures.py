"""The ctypes binding (isplib_b200/capi.py) against the prototypes of include/isplib_b200.h, argument by
argument, and the three structs against what gcc lays out for the header -- no GPU needed.  A binding that
drifts from the header passes garbage to the kernels without any error on the host side."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "isplib_b200.h")


def prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(isplib_b200_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def c_kind(decl: str) -> str:
    """'ptr' | 'i32' | 'i64' | 'f32' | 'f64' | 'size' for one C parameter declaration."""
    if "*" in decl or "isplib_stream_t" in decl:
        return "ptr"
    words = decl.replace("const", " ").split()
    ty = words[0] if len(words) > 1 else decl.strip()
    return {"int": "i32", "int32_t": "i32", "uint32_t": "i32", "int64_t": "i64", "uint64_t": "i64",
            "float": "f32", "double": "f64", "size_t": "size"}[ty]


def ctypes_kind(t) -> str:
    if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    if t is ctypes.c_size_t:
        return "size"
    if t is ctypes.c_float:
        return "f32"
    if t is ctypes.c_double:
        return "f64"
    if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_uint32):
        return "i32"
    if t in (ctypes.c_int64, ctypes.c_uint64, ctypes.c_longlong):
        return "i64"
    raise AssertionError(f"unclassified ctypes type {t}")


def test_every_prototype_is_parsed():
    from isplib_b200 import capi
    protos = prototypes()
    assert sorted(protos) == sorted(capi.EXPORTS)


def test_argtypes_match_the_header_argument_by_argument():
    from isplib_b200 import capi
    lib = capi.lib()
    for name, (ret, params) in prototypes().items():
        fn = getattr(lib, name)
        want = [c_kind(p) for p in params]
        if fn.argtypes is None:
            assert not want, f"{name}: the header takes {len(want)} arguments, the binding declares none"
            continue
        got = [ctypes_kind(t) for t in fn.argtypes]
        # size_t and a 64-bit integer travel the same way on this ABI; keep them distinct anyway
        assert got == want, f"{name}: binding {got} != header {want}"
        if "char" in ret:
            assert fn.restype is ctypes.c_char_p, f"{name} returns a string"
        else:
            assert fn.restype in (ctypes.c_int, ctypes.c_int32), f"{name} returns a status int"


def test_struct_layouts_match_what_gcc_lays_out(tmp_path):
    from isplib_b200 import capi
    structs = {"isplib_b200_plan_info": capi.PlanInfo, "isplib_b200_epilogue": capi.Epilogue,
               "isplib_b200_gather_desc": capi.GatherDesc}
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    fields = {}
    for cname, cls in structs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part.replace("*", " ").split()[-1]))
        fields[cname] = names
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in names:
            lines.append(f'printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    got = {}
    for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        s, f, v = line.split()
        got[(s, f)] = int(v)
    for cname, cls in structs.items():
        assert ctypes.sizeof(cls) == got[(cname, "size")], cname
        assert [f for f, _ in cls._fields_] == fields[cname], f"{cname}: field names / order differ from the header"
        for f, _ in cls._fields_:
            assert getattr(cls, f).offset == got[(cname, f)], f"{cname}.{f}"
