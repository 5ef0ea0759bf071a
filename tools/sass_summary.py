#!/usr/bin/env python
"""Static evidence for the kernels in libisplib_b200.so, obtainable without a GPU (B200_PROFILING.md:
"check -Xptxas -v and cuobjdump -sass before spending GPU time"): per kernel template the registers, stack
frame (spills), shared memory and the count of the memory / reduction mnemonics that characterise it --
256-bit gathers (LDG.E.ENL2.256), 128-bit loads, TMA bulk copies (UBLKCP), reductions (RED / ATOM),
warp shuffles, FFMA / FMUL / FMNMX / FSETP / SEL.

    python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.md
    make -C isplib_b200/csrc -j8 ab TAG=ptxasv PTXAS_V="-Xptxas -v" > ptxas.log 2>&1     # optional: spill bytes
    python tools/sass_summary.py --ptxas-log ptxas.log > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = sys.argv[1:]
PTXAS_LOG = None
if "--ptxas-log" in ARGS:
    i = ARGS.index("--ptxas-log")
    PTXAS_LOG = ARGS[i + 1]
    del ARGS[i:i + 2]
LIB = ARGS[0] if ARGS else os.path.join(ROOT, "isplib_b200", "libisplib_b200.so")
# longest prefix first: an opcode is counted under the first entry it starts with
MNEMONICS = ["LDG.E.ENL2.256", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.ENL2.256", "STG.E.128", "STG.E", "ST.E", "UBLKCP",
             "RED.", "ATOM", "SHFL", "FFMA", "FMUL", "FMNMX", "FSETP", "FSEL", "SEL", "LDS", "STS", "BAR.SYNC", "MEMBAR"]
COLUMNS = ["LDG.E.ENL2.256", "LDG.E.128", "UBLKCP", "RED.", "ATOM", "SHFL", "FFMA", "FMUL", "FMNMX", "FSETP", "SEL", "FSEL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("isplib::", "").replace("(anonymous namespace)::", "")


def spills_from_ptxas_log(path):
    """mangled name -> (spill store bytes, spill load bytes) from an `nvcc -Xptxas -v` build log."""
    out, cur = {}, None
    for line in open(path, errors="replace"):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            out[cur] = (int(m.group(2)), int(m.group(3)))
            cur = None
    return out


def main():
    spills = spills_from_ptxas_log(PTXAS_LOG) if PTXAS_LOG else {}
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    n_inst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            n_inst[cur] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
                    break
    names = demangle(sorted(usage))
    rows = sorted(usage, key=lambda k: short(names[k]))
    print(f"# SASS / resource summary of `{os.path.relpath(LIB, ROOT)}` (sm_100a)\n")
    print("Produced by `python tools/sass_summary.py` from `cuobjdump --dump-resource-usage` and `cuobjdump -sass` "
          "(no GPU needed).")
    print("Static counts over each kernel's whole body, not executed-instruction counts.  STACK > 0 means a stack "
          "frame (spills or a local array).\n")
    print("| kernel | REG | STACK | spill st/ld B | SHARED | SASS instr | " + " | ".join(c.rstrip(".") for c in COLUMNS) + " |")
    print("|---|---|---|---|---|---|" + "---|" * len(COLUMNS))
    tot = collections.Counter()
    for k in rows:
        u, c = usage[k], counts[k]
        for mn in MNEMONICS:
            tot[mn] += c[mn]
        sp = spills.get(k)
        sp_txt = "" if sp is None or sp == (0, 0) else f"{sp[0]}/{sp[1]}"
        print(f"| `{short(names[k])}` | {u.get('REG', 0)} | {u.get('STACK', 0)} | {sp_txt} | {u.get('SHARED', 0)} | {n_inst[k]} | "
              + " | ".join(str(c[x]) if c[x] else "" for x in COLUMNS) + " |")
    print(f"\n{len(rows)} kernels; totals: " + ", ".join(f"{mn.rstrip('.')} {tot[mn]}" for mn in MNEMONICS if tot[mn]))
    frames = [f"{short(names[k])} ({usage[k]['STACK']} B)" for k in rows if usage[k].get("STACK", 0) > 0]
    if spills:
        spilled = [k for k in rows if spills.get(k, (0, 0)) != (0, 0)]
        print(f"\nKernels with register spills (ptxas -v): {len(spilled)} of {len(rows)}; the rest of the stack frames are "
              "local arrays indexed at run time.")
    print(f"\nKernels with a stack frame: {len(frames)}" + ("" if not frames else " -- " + "; ".join(frames[:30])))


if __name__ == "__main__":
    main()
