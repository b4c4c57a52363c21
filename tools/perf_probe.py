#!/usr/bin/env python3
"""Quick device-side throughput probe (not the bench): cfg1 / cfg2 slices, haystack resident in HBM."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
import torch  # noqa: E402
from fac_b200 import GpuBackend, SearchOptions, workload  # noqa: E402


def probe(name, cfg, reps=3, prefilter=False, order=0, overlap=0):
    gpu = GpuBackend()
    t0 = time.time()
    eng = workload.build_engine(cfg, gpu)
    build_s = time.time() - t0
    text = cfg["text"]
    d = torch.from_numpy(text).cuda()
    best = None
    for _ in range(reps):
        arr, st = gpu.search_device(eng._h, d.data_ptr(), d.numel(), cfg["threshold"], order, overlap, prefilter)
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
            nm = len(arr)
    n = d.numel()
    print("%s: %d B, nodes %d, build %.2fs, %d matches, device %.2f ms (expand %.2f ms, %d launches), %.4f GB/s, "
          "%.1f states/window, %.3e states/s" %
          (name, n, eng.num_nodes(), build_s, nm, best["device_ms"], best["expand_ms"], best["kernel_launches"],
           n / best["device_ms"] / 1e6, best["states_pushed"] / n, best["states_pushed"] / best["expand_ms"] * 1e3),
          flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "cfg1"):
        probe("cfg1", workload.cfg1(int(os.environ.get("CFG1_BYTES", 16 << 20))))
    if which in ("cfg4",):
        c4 = workload.cfg4(int(os.environ.get("CFG4_BYTES", 256 << 20)))
        probe("cfg4+prefilter(sorted,non_overlapping)", c4, prefilter=True, order=1, overlap=1)
    if which in ("all", "cfg2"):
        probe("cfg2", workload.cfg2(int(os.environ.get("CFG2_BYTES", 1 << 20)), int(os.environ.get("CFG2_PATTERNS", 10000))))
    if which in ("cfg3",):
        probe("cfg3 (unicode, mappings)", workload.cfg3(int(os.environ.get("CFG3_BYTES", 8 << 20))), reps=2)
    if which in ("cfg2u",):  # cfg2 with one non-ASCII grapheme every ~1 KiB: the whole haystack takes the K1 + grapheme-stream route
        import numpy as np
        c = workload.cfg2(int(os.environ.get("CFG2_BYTES", 8 << 20)), int(os.environ.get("CFG2_PATTERNS", 10000)))
        t = c["text"].copy()
        for p in range(512, len(t) - 2, 1024):
            t[p] = 0xC3; t[p + 1] = 0xA9   # e-acute
        c["text"] = t
        probe("cfg2 + non-ASCII graphemes", c)
