#!/usr/bin/env python
"""Where a kernel spends its instructions, by SOURCE LINE, from an `ncu --set full --import-source on` report.

ncu's CSV source page lists SASS instructions with executed counts and stall samples but no line numbers; nvdisasm -g
lists the same instructions of the same cubin with `//## File ..., line N` markers (compile with -lineinfo).  The two
are joined by instruction index.  The library must be the build the report was taken from.

  python tools/sass_lines.py report.ncu-rep <mangled-kernel-substring> [--lib faucet_b200/libfaucet_gpu.so] [--top 40]
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import tempfile


def disasm_lines(lib, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    lines, cur, on = [], ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            lines.append((cur, ln.split("*/", 1)[1].strip()))
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--lib", default="faucet_b200/libfaucet_gpu.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ranges", default="", help="file:lo-hi[=name],... : also sum over these line ranges")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    i_s, i_x, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    dl = disasm_lines(a.lib, a.kernel)
    if len(dl) != len(data):
        print(f"warning: {len(data)} instructions in the report, {len(dl)} in the library (different build?)")
    by = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_x = 0
    for (loc, _), r in zip(dl, data):
        s, x = int(r[i_s] or 0), int(r[i_x] or 0)
        by[loc][0] += s
        by[loc][1] += x
        tot_s += s
        tot_x += x
    print(f"{tot_x} warp instructions, {tot_s} samples")
    for loc, (s, x) in sorted(by.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{loc[0]}:{loc[1]:<6} instr {100 * x / tot_x:5.1f} %   samples {100 * s / tot_s:5.1f} %")
    for spec in filter(None, a.ranges.split(",")):
        name = spec
        if "=" in spec:
            spec, name = spec.split("=")
        f, r = spec.split(":")
        lo, hi = map(int, r.split("-"))
        s = sum(v[0] for k, v in by.items() if k[0] == f and lo <= k[1] <= hi)
        x = sum(v[1] for k, v in by.items() if k[0] == f and lo <= k[1] <= hi)
        print(f"range {name:<28} instr {100 * x / tot_x:5.1f} %   samples {100 * s / tot_s:5.1f} %")


if __name__ == "__main__":
    main()
