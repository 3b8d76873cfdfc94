#!/usr/bin/env python3
"""Per-variant resource usage and SASS opcode histogram of the built libraries -> Markdown (profiles/rNN_sass_summary.md).

    python tools/sass_summary.py > profiles/r02_sass_summary.md

For every library of noahmp_b200/build.py: `cuobjdump -res-usage` (registers, stack frame, static shared memory) of each
kernel, and for the benchmark's kernel instantiation the opcode histogram of `cuobjdump -sass` (code size, MUFU / LDL / STL /
BAR counts, whether any tensor-core or TMA opcode is present - none is expected: the path has no contraction)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from noahmp_b200.build import VARIANTS, lib_path  # noqa: E402

BENCH_KERNEL = "column_step_kernelINS_10OptDefaultELb0ELb0ELb0"


def short(name: str) -> str:
    m = re.search(r"column_step_kernelINS_(\d+)(Opt\w+?)ELb(\d)ELb(\d)ELb(\d)", name)
    if m:
        return f"column_step_kernel<{m.group(2)}, EXT={m.group(3)}, DIAG={m.group(4)}, BUDGET={m.group(5)}>"
    m = re.search(r"_ZN?\d*\w*?(\d+)([a-z_]+kernel)", name)
    return m.group(2) if m else name[:60]


def res_usage(lib):
    txt = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    out, fn = [], None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
        if m and fn:
            out.append((fn, int(m.group(1)), int(m.group(2)), int(m.group(3))))
            fn = None
    return out


def histogram(lib, kernel_sub):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", kernel_sub, lib], capture_output=True, text=True).stdout
    if "/*0000*/" not in txt:   # -fun wants the full mangled name: fall back to scanning the whole dump
        txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        keep, on = [], False
        for ln in txt.splitlines():
            if "Function :" in ln:
                on = kernel_sub in ln
            if on:
                keep.append(ln)
        txt = "\n".join(keep)
    ops = collections.Counter()
    for ln in txt.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            ops[m.group(1)] += 1
    return ops


def main():
    print("# SASS summary of the built libraries (sm_100a cubins; `tools/sass_summary.py`)\n")
    print("| library | kernel | registers | stack frame B | static smem B |\n|---|---|---|---|---|")
    libs = {v: lib_path(v) for v in VARIANTS}
    for v, lib in libs.items():
        if not os.path.exists(lib):
            continue
        for fn, reg, stack, smem in res_usage(lib):
            print(f"| {os.path.basename(lib)} | `{short(fn)}` | {reg} | {stack} | {smem} |")
    print("\n## Opcode histogram of the benchmark's instantiation `column_step_kernel<OptDefault, 0, 0, 0>`\n")
    names = [v for v in libs if os.path.exists(libs[v])]
    hists = {v: histogram(libs[v], BENCH_KERNEL) for v in names}
    groups = [("total SASS instructions", None), ("MUFU (exp2 / log2 / rcp / rsq / sqrt / sin / cos)", ("MUFU",)),
              ("FFMA FMUL FADD", ("FFMA", "FMUL", "FADD")), ("FSETP FSEL FMNMX", ("FSETP", "FSEL", "FMNMX")),
              ("DADD DMUL DFMA (fp64 island of groundwater, fp64 libm of the strict builds)", ("DADD", "DMUL", "DFMA")),
              ("LDL (local loads: spills + snow scratch)", ("LDL",)), ("STL (local stores)", ("STL",)),
              ("LDG STG (global)", ("LDG", "STG")), ("LDS STS (shared: tables)", ("LDS", "STS")),
              ("BAR (phase barriers)", ("BAR",)), ("BRA BSSY BSYNC", ("BRA", "BSSY", "BSYNC")),
              ("CALL RET (combine / divide, libm slow paths)", ("CALL", "RET")),
              ("tensor core / TMA / TMEM (UTCMMA* UTMA* HMMA* LDTM*)", None)]
    print("| | " + " | ".join(names) + " |\n|---|" + "---|" * len(names))
    for title, ops in groups:
        row = []
        for v in names:
            h = hists[v]
            if title.startswith("total"):
                row.append(str(sum(h.values())))
            elif title.startswith("tensor"):
                row.append(str(sum(n for o, n in h.items() if o.startswith(("UTC", "UTMA", "HMMA", "LDTM", "STTM", "UBLKCP")))))
            else:
                row.append(str(sum(h.get(o, 0) for o in ops)))
        print(f"| {title} | " + " | ".join(row) + " |")
    print("\nCode size = 16 B x instructions.  No tensor-core, TMA or TMEM opcode: the column step has no contraction and "
          "its memory traffic is one coalesced 128-byte line per warp access (DESIGN.md section 4).")
    top = hists.get("fast")
    if top:
        print("\nTop 25 opcodes, fast build: " + ", ".join(f"{o} {n}" for o, n in top.most_common(25)))


if __name__ == "__main__":
    main()
