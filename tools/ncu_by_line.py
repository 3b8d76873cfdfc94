#!/usr/bin/env python3
"""Attribute the per-SASS-instruction counters of an ncu report to source lines and to physics phases.

ncu's CSV export of the source page is SASS-only, so the file:line of every instruction is taken from
`nvdisasm -g` of the SAME library (`cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin`) and joined on the
instruction offset inside the kernel.

  python tools/ncu_by_line.py REPORT.ncu-rep LIB.so [KERNEL_SUBSTRING] [--top N] [--csrc DIR]

--csrc: the csrc/ directory the library was built from (default: the working tree), for the line -> phase table.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

# physics phases by (file, first line) in source order; filled from the marker comments of the sources
PHASE_MARKERS = {
    "energy.cuh": [(r"void esat\(", "esat"), (r"float tdc\(", "tdfcnd"), (r"void sfcdif1\(", "sfcdif1"),
                   (r"void sfcdif2\(", "sfcdif2"), (r"struct LeafParams", "stomata"), (r"void canres\(", "canres"),
                   (r"struct TwoStreamCommon", "twostream"), (r"struct Loop1", "vege_flux.loop1"), (r"void energy\(", "energy")],
    "water.cuh": [(r"void combo\(", "snow.combine/divide"), (r"void pack_snow\(", "water"), (r"void wdfcnd1\(", "wdfcnd"),
                  (r"void water\(", "water"), (r"void carbon\(", "carbon")],
}


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, chain, fresh = None, (), True
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur = m.group(1)
                out[cur] = {}
                continue
            m = re.match(r'\s*//## File "(.*?)", line (\d+)(?: inlined at "(.*?)", line (\d+))?', ln)
            if m:   # consecutive annotation lines spell the inline chain, innermost frame first
                if fresh:
                    chain, fresh = [(os.path.basename(m.group(1)), int(m.group(2)))], False
                if m.group(3):
                    chain.append((os.path.basename(m.group(3)), int(m.group(4))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur:
                out[cur][int(m.group(1), 16)] = (tuple(chain), m.group(2).strip())
                fresh = True
    ks = [k for k in out if kernel_sub in k]
    if len(ks) != 1:
        sys.exit(f"kernel substring {kernel_sub!r} matches {ks}")
    return out[ks[0]]


def ncu_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    return [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]


def phase_table(csrc):
    """(file, line) -> phase name: the last marker above the line; finer split inside energy()/water() by the
    `// ---- name` section comments."""
    tab = {}
    for f in ("energy.cuh", "water.cuh", "column_step.cuh", "common.cuh", "forcing_prep.cuh"):
        p = os.path.join(csrc, f)
        if not os.path.exists(p):
            continue
        cur = f.split(".")[0]
        marks = PHASE_MARKERS.get(f, [])
        for i, ln in enumerate(open(p), 1):
            for pat, name in marks:
                if re.search(r"__device__.*" + pat, ln) or re.match(pat, ln.strip()):
                    cur = name
            m = re.match(r"\s*// -{2,} ([A-Za-z0-9_ /+,().:-]+?)(,| -|$)", ln)
            if m and cur.split(":")[0] in ("energy", "water"):
                cur = cur.split(":")[0] + ":" + m.group(1).strip()[:28]
            tab[(f, i)] = cur
    return tab


def main():
    argv = sys.argv[1:]
    top = 25
    if "--top" in argv:
        i = argv.index("--top")
        top = int(argv[i + 1])
        del argv[i:i + 2]
    csrc = None
    if "--csrc" in argv:
        i = argv.index("--csrc")
        csrc = argv[i + 1]
        del argv[i:i + 2]
    args = argv
    rep, lib = args[0], args[1]
    ksub = args[2] if len(args) > 2 else "OptDefaultELb0ELb0ELb0"
    sass = sass_lines(lib, ksub)
    rows = ncu_rows(rep)
    base = int(rows[0]["Address"], 16)
    here = os.path.dirname(os.path.abspath(__file__))
    ptab = phase_table(csrc or os.path.join(here, "..", "noahmp_b200", "csrc"))
    NV = 9   # instr, thread-instr, samples, stall_long_sb, stall_barrier, stall_wait, stall_short_sb, local loads, local stores
    by_line, by_phase, by_leaf = (defaultdict(lambda: [0] * NV) for _ in range(3))
    tot = [0] * NV
    mism = 0
    for r in rows:
        off = int(r["Address"], 16) - base
        chain, text = sass.get(off, ((), ""))
        if text.split()[:1] != r["Source"].split()[:1] and not (text.startswith("@") or r["Source"].strip().startswith("@")):
            mism += 1
        op = r["Source"].replace("@", " ").split()
        op = next((o for o in op if not o.startswith(("P", "!P", "UP", "!UP"))), "")
        nexe = int(r["Instructions Executed"])
        v = (nexe, int(r["Thread Instructions Executed"]), int(r["# Samples"]), int(r.get("stall_long_sb", 0) or 0),
             int(r.get("stall_barrier", 0) or 0), int(r.get("stall_wait", 0) or 0), int(r.get("stall_short_sb", 0) or 0),
             nexe if op.startswith("LDL") else 0, nexe if op.startswith("STL") else 0)
        body = [c for c in chain if c[0] != "common.cuh"] or list(chain)      # skip the libm / static_for wrappers
        loc = body[0] if body else None                                         # innermost physics line
        leaf = ptab.get(loc, loc[0] if loc else "?")
        # phase = the innermost frame that lies in the body of energy()/water()/the kernel (not in a leaf routine).  The
        # OUTER frames of the debug info cannot be trusted: instantiations of static_for with different lambdas share
        # one "inlined at" record, so an outermost-frame rule files phasechange under the snow compaction loop.
        ph = next((ptab[c] for c in body if c in ptab and (":" in ptab[c] or ptab[c] == "column_step")), leaf)
        for acc in (by_line[loc], by_phase[ph], by_leaf[leaf], tot):
            for i in range(NV):
                acc[i] += v[i]
    if mism > len(rows) // 50:
        print(f"WARNING: {mism} of {len(rows)} instructions do not match the library's SASS - different build?")
    print(f"total warp-instructions {tot[0]:.4g}, lanes/instr {tot[1] / max(tot[0], 1):.2f}, samples {tot[2]}")
    for title, tab in (("phase of the step (callees included)", by_phase),
                       ("routine (innermost frame)", by_leaf)):
        print(f"\n| {title} | warp-instr % | lanes active | stall samples % | long_sb % | barrier % | wait % | short_sb % | "
              f"local ld+st % of its instr |\n|---|---|---|---|---|---|---|---|---|")
        for ph, v in sorted(tab.items(), key=lambda kv: -kv[1][0]):
            if v[0] * 1000 < tot[0]:
                continue
            print(f"| {ph} | {100 * v[0] / tot[0]:.1f} | {v[1] / max(v[0], 1):.1f} | {100 * v[2] / max(tot[2], 1):.1f} | "
                  f"{100 * v[3] / max(tot[2], 1):.1f} | {100 * v[4] / max(tot[2], 1):.1f} | {100 * v[5] / max(tot[2], 1):.1f} | "
                  f"{100 * v[6] / max(tot[2], 1):.1f} | {100 * (v[7] + v[8]) / max(v[0], 1):.1f} |")
    print(f"\ntop {top} source lines\n| file:line | warp-instr % | lanes | samples % |\n|---|---|---|---|")
    for loc, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {loc[0]}:{loc[1]} | {100 * v[0] / tot[0]:.2f} | {v[1] / max(v[0], 1):.1f} | {100 * v[2] / max(tot[2], 1):.2f} |" if loc else f"| ? | {100 * v[0] / tot[0]:.2f} | | |")


if __name__ == "__main__":
    main()
