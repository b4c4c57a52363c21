#!/usr/bin/env python3
"""Exact engines (no FuzzyLimits): cfg2's patterns without any edit budget."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
import torch
from fac_b200 import FuzzyAhoCorasickBuilder, GpuBackend, workload
n = int(os.environ.get("EXACT_BYTES", 64 << 20))
cfg = workload.cfg2(n, 10000)
gpu = GpuBackend()
eng = FuzzyAhoCorasickBuilder.new(gpu).build(cfg["patterns"])
d = torch.from_numpy(cfg["text"]).cuda()
for _ in range(3):
    arr, st = gpu.search_device(eng._h, d.data_ptr(), d.numel(), 0.8, 0, 0, False)
print("exact engine, 10k patterns: %d B, %d matches, device %.2f ms (expand %.2f ms) -> %.3f GB/s" %
      (n, len(arr), st["device_ms"], st["expand_ms"], n / st["device_ms"] / 1e6))
