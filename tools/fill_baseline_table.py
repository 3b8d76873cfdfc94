#!/usr/bin/env python3
"""Fill the result table of BASELINE.md section 5 from bench.py's JSON lines (never from prose).

    python tools/fill_baseline_table.py N1.json [N2.json N4.json N8.json] [--write]

N1.json: the default line of `python bench.py` (N = 1: carries other_workloads, parity objects, cpu_baseline.per_config);
the optional files are the lines of `bench.py --gpus 2 / 4 / 8`.  Prints the Markdown table; with --write replaces the
table under "## 5." in BASELINE.md and records the source files below it."""
import json
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
ROWS = [("single", "1 single point (`run/case.nml`)", "1 × 96"), ("uniform100k", "2 uniform 100 k, 1 day @ 30 min", "1×10⁵ × 48"),
        ("global025", "3 global 0.25°, 1 yr hourly", "2.8×10⁵ × 8 760"), ("conus1km", "4 CONUS 1 km", "1.7×10⁷ × 24"),
        ("snow1m", "5 snow-heavy ensemble", "1×10⁶ × 720")]


def sci(x):
    return "—" if x is None else f"{x:.3g}".replace("e+0", "e").replace("e+", "e")


def bpc(x) -> float:
    """Bytes that crossed the host link per cell-step of the e2e leg (counted by the library; 32 for dense records)."""
    c = x["config"]
    return x["e2e"]["h2d_bytes_per_step"] / (c["columns"] * c["model_steps_per_step"])


def scaling_table(by_n) -> str:
    """Section 6: device-timed and end-to-end scaling beside the measured host link."""
    one = by_n[1]
    out = ["| GPUs (which) | device-timed cell-steps/s | scaling | e2e cell-steps/s (partition) | e2e scaling | e2e, equal blocks | host link while all ranks copy, GB/s: "
           "slowest rank x N / sum of the ranks' own rates | bytes over the host link per cell-step (records) | link traffic of e2e, GB/s |", "|---|---|---|---|---|---|---|---|---|"]
    for n in sorted(by_n):
        x = by_n[n]
        e, h = x["e2e"], x["e2e"]["host_link"]
        g = x["config"].get("gpus", list(range(n)))
        eq = (e.get("equal_blocks") or {}).get("value")
        part = "link-weighted" if eq else "equal"
        rates = h.get("h2d_gbs_sum_while_all_copy")
        packed = "zero-suppressed" in e.get("records", "")
        # like with like: a line with dense records is set against the one-GPU line's dense-records leg
        base = one["e2e"]["value"] if packed or "dense_records" not in one["e2e"] else one["e2e"]["dense_records"]["value"]
        out.append(f"| {n} ({', '.join(map(str, g))}) | {sci(x['value'])} | {x['value'] / one['value']:.2f}x | {sci(e['value'])} ({part}) | "
                   f"{e['value'] / base:.2f}x{'' if packed or n == 1 else ' of the dense one-GPU leg'} | {sci(eq) if eq else 'same'} | {h['h2d_gbs_all_gpus']:.0f} / "
                   f"{(f'{rates:.0f}' if rates else 'n/a')} | {bpc(x):.1f} ({'packed' if 'zero-suppressed' in e.get('records', '') else 'dense'}) | "
                   f"{e['value'] * bpc(x) / 1e9:.0f} |")
    return "\n".join(out)


def main():
    argv = [a for a in sys.argv[1:] if a != "--write"]
    lines = [json.load(open(a)) for a in argv]
    n1 = next(d for d in lines if d["n_gpus"] == 1)
    by_n = {d["n_gpus"]: d for d in lines}
    cpu = (n1.get("cpu_baseline") or {})
    per = cpu.get("per_config", {})
    others = n1.get("other_workloads") or {}
    out = ["| config | columns × steps | CPU port cell-steps/s (cores; sample) | 1×B200 | 2× | 4× | 8× | % HBM roofline (1×) | issue-slot util (1×) | "
           "state error vs fp64 twin, tolerance units: worst p95 / worst max, build (fp32 reference) | loop1 iter-count mismatch, GPU vs oracle "
           "(two CPU builds) |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for key, label, size in ROWS:
        src = n1 if key == "conus1km" else others.get(key, {})
        val = src.get("value")
        if key == "conus1km":
            frac, iss, par = n1["roofline"]["frac"], (n1.get("roofline_issue") or {}).get("frac"), n1.get("parity")
            c = cpu if cpu.get("value") else per.get(key, {})
            scale = [sci(by_n[n]["value"]) + f" (e2e {sci(by_n[n]['e2e']['value'])})" if n in by_n and by_n[n].get("e2e") else
                     (sci(by_n[n]["value"]) if n in by_n else "") for n in (2, 4, 8)]
            one = sci(val) + (f" (e2e {sci(n1['e2e']['value'])})" if n1.get("e2e") else "")
        else:
            frac, iss, par = src.get("roofline_frac"), src.get("issue_frac_at_conus_instruction_count"), src.get("parity")
            c = per.get(key, {})
            scale = ["—"] * 3
            one = sci(val)
        cputxt = f"{sci(c.get('value'))} ({c.get('cores', '?')}; {c.get('sample', '').split(' (')[0]})" if c.get("value") else ""
        if par and "worst_p95" in par:
            err = (f"{par['worst_p95']['build_tol_units']:.2g} ({par['worst_p95']['ref32_tol_units']:.2g}) {par['worst_p95']['plane']} / "
                   f"{par['worst_max']['build_tol_units']:.3g} ({par['worst_max']['ref32_tol_units']:.3g}) {par['worst_max']['plane']}; "
                   f"{'pass' if par.get('pass') else 'FAIL'}, {par['columns']} columns")
        else:
            err = "strict build bit-exact vs oracle (tests)" if key == "single" else ""
        mis = (per.get(key) or {}).get("iteration_count_mismatch")
        mistxt = (f"{100 * mis['gpu_vs_oracle']['loop1']:.2f} % ({100 * mis['two_cpu_builds_of_the_oracle']['loop1']:.2f} %), "
                  f"{mis['columns']} columns") if mis else ("n/a" if key == "single" else "")
        pct = lambda x: "n/a" if key == "single" else ("" if x is None else f"{100 * x:.1f} %")  # noqa: E731
        out.append(f"| {label} | {size} | {cputxt} | {one} | {scale[0]} | {scale[1]} | {scale[2]} | {pct(frac)} | {pct(iss)} | {err} | {mistxt} |")
    table = "\n".join(out)
    note = ("\nSources: " + ", ".join(f"`{os.path.relpath(a, ROOT)}`" for a in argv) + " (lines of `bench.py`; filled by "
            "`tools/fill_baseline_table.py`).  GPU figures: fast build, CUDA events around the step launches.  Issue-slot utilisation of "
            "configs 1-3 and 5 uses the warp-instruction count of the CONUS capture (an approximation: their instruction mix differs).  "
            "The state-error column is in units of the north star's tolerance (rel 1e-4, abs floor 1e-6) against the fp64 twin of the "
            "translated reference after the config's FULL step count; the figure in brackets is the fp32 reference's own distance "
            "from that twin.  `single` has one column: no percentile, no roofline.\n")
    print(table + "\n" + note)
    scal = scaling_table(by_n) if len(by_n) > 1 and all(d.get("e2e") for d in by_n.values()) else None
    if scal:
        print(scal)
    if "--write" in sys.argv:
        p = os.path.join(ROOT, "BASELINE.md")
        s = open(p).read()
        a, b = s.index("<!-- table5 -->"), s.index("<!-- /table5 -->")
        s = s[:a] + "<!-- table5 -->\n" + table + "\n" + note + s[b:]
        if scal and "<!-- table6 -->" in s:
            a, b = s.index("<!-- table6 -->"), s.index("<!-- /table6 -->")
            s = s[:a] + "<!-- table6 -->\n" + scal + "\n" + s[b:]
        open(p, "w").write(s)


if __name__ == "__main__":
    main()
