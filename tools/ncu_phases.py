#!/usr/bin/env python3
"""Group the executed warp instructions of one kernel by basic-block execution count.

    tools/ncu_phases.py <report.ncu-rep> <libfacgpu.so> <kernel-substring> <units> [top-n]

Every SASS instruction of a loop body executes the same number of times, so grouping `ncu --page source --csv` rows by
their "Instructions Executed" value separates the phases of a persistent kernel (pop rounds, item rounds, walk steps,
loop control ...) without needing inline call chains.  `units` = work units of the launch (e.g. start windows): the
table is printed per unit.  The .so must be the build that was profiled (nvdisasm -g maps offsets to file:line)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, sub, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, start = None, 0
    for i, r in enumerate(rows):
        if "Address" in r and "Source" in r:
            hdr, start = r, i + 1
            break
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[start:]
    base = int(data[0][ci["Address"]], 16)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    line_of, cur, inside = {}, None, False
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        for ln in sass.splitlines():
            if ln.startswith(".text."):
                inside = sub in ln and not line_of
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m:
                line_of[int(m.group(1), 16)] = cur
    groups, tot = collections.OrderedDict(), 0
    for r in data:
        off = int(r[ci["Address"]], 16) - base
        ni = int(r[ci["Instructions Executed"]])
        nt = int(r[ci["Thread Instructions Executed"]]) if "Thread Instructions Executed" in ci else 0
        groups.setdefault(ni, []).append((off, nt, line_of.get(off)))
        tot += ni
    print("total warp instructions %d = %.1f per unit" % (tot, tot / units))
    print("instr/unit  (static instr x executions/unit, active threads)  dominant source lines (static instr)")
    for ni, lst in sorted(groups.items(), key=lambda kv: -kv[0] * len(kv[1]))[:topn]:
        lines = collections.Counter(l for _, _, l in lst if l)
        top = ", ".join("%s:%d(%d)" % (k[0], k[1], v) for k, v in lines.most_common(6))
        thr = sum(nt for _, nt, _ in lst) / max(1, ni * len(lst))
        print("%8.1f  (%4d x %7.3f, %4.1f thr)  %s" % (ni * len(lst) / units, len(lst), ni / units, thr, top))


if __name__ == "__main__":
    main()
