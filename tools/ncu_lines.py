#!/usr/bin/env python3
"""Attribute ncu warp-stall samples to CUDA source lines.

    tools/ncu_lines.py <report.ncu-rep> <libfacgpu.so> [kernel-substring] [top-n]

`ncu --page source --csv` lists samples per SASS instruction; `nvdisasm -g` maps SASS offsets to
file:line (innermost inlined frame).  The two are joined on the instruction offset inside the kernel.
The .so must be the build that was profiled."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, so = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else "k_expand"
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the report may hold several kernels: split on "Kernel Name" rows
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    k = [k for k in kernels if sub in k["name"]][0]
    ci = {h: i for i, h in enumerate(k["hdr"])}
    base = int(k["rows"][0][ci["Address"]], 16)
    samples = {}
    for r in k["rows"]:
        samples[int(r[ci["Address"]], 16) - base] = (int(r[ci["# Samples"]]), r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]),
                                                    int(r[ci["Thread Instructions Executed"]]))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    # mangled name from the demangled one: match template args by the bool list
    targs = re.findall(r"\((bool|int)\)(\d+)", k["name"])
    fn = re.match(r"void (\w+)", k["name"]).group(1)
    pat = re.compile(r"\.text\._Z\d+%s%s" % (fn, ("I" + "".join("L%s%sE" % ("b" if t == "bool" else "i", v) for t, v in targs) + "E") if targs else ""))
    line_of, cur_line, active = {}, None, False
    for cb in cubins:
        dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        for ln in dis.splitlines():
            if ln.startswith("\t.section"):
                active = bool(pat.search(ln))
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m:
                line_of[int(m.group(1), 16)] = cur_line
    agg = defaultdict(int)
    iagg, tagg = defaultdict(int), defaultdict(int)
    tot = itot = 0
    for off, (s, _, ni, nt) in samples.items():
        agg[line_of.get(off)] += s
        iagg[line_of.get(off)] += ni
        tagg[line_of.get(off)] += nt
        tot += s
        itot += ni
    print("kernel:", k["name"], "total samples", tot, "warp instructions", itot)
    print("stall%  inst%  thr/inst  line")
    srcs = {}
    for (key, s) in sorted(agg.items(), key=lambda kv: -kv[1])[:topn]:
        text = ""
        if key:
            f, l = key
            if f not in srcs:
                p = os.path.join(os.path.dirname(os.path.abspath(so)), f)
                srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
            if 0 < l <= len(srcs[f]):
                text = srcs[f][l - 1].strip()[:100]
        print("%5.2f%% %5.2f%% %5.1f  %-22s %s" % (100.0 * s / max(tot, 1), 100.0 * iagg[key] / max(itot, 1), tagg[key] / max(iagg[key], 1),
                                             "%s:%d" % key if key else "?", text))


if __name__ == "__main__":
    main()
