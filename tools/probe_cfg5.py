#!/usr/bin/env python3
"""cfg5-shaped probe: 1000 patterns, edits(2), case-insensitive, auto_beam(200000, 100); 256 KiB stream windows
searched with sorted().non_overlapping() (src/stream.rs:262-297)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, GpuBackend, SearchOptions, workload  # noqa: E402

n = int(os.environ.get("CFG5_BYTES", 1 << 20))
cfg = workload.cfg2(n, 1000, seed=0xFAC00005)
gpu = GpuBackend()
for label, mk in (("auto_beam(200000,100)", lambda b: b.auto_beam(200000, 100)), ("no beam", lambda b: b)):
    eng = mk(FuzzyAhoCorasickBuilder.new(gpu).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True)).build(cfg["patterns"])
    text = bytes(cfg["text"])
    opts = SearchOptions.new().threshold(0.8).sorted().non_overlapping()
    for rep in range(2):
        t0 = time.time()
        tot = 0
        for a in range(0, n, 256 << 10):
            r = eng.search(text[a:a + (256 << 10)], opts)
            tot += len(r)
        dt = time.time() - t0
    print("%s: %d B in 256 KiB windows: %.3f s -> %.4f GB/s, %d matches" % (label, n, dt, n / dt / 1e9, tot), flush=True)

# streaming API throughput (fac_search_stream): windows cut by the library, searched in batches
import io
for label, mk in (("stream no beam", lambda b: b),):
    eng = mk(FuzzyAhoCorasickBuilder.new(gpu).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True)).build(cfg["patterns"])
    big = bytes(cfg["text"]) * max(1, (32 << 20) // n)
    for rep in range(2):
        cnt = [0]
        t0 = time.time()
        eng.search_stream(io.BytesIO(big), 0.8, lambda m: cnt.__setitem__(0, cnt[0] + 1))
        dt = time.time() - t0
    print("%s: %d B: %.3f s -> %.4f GB/s, %d matches" % (label, len(big), dt, len(big) / dt / 1e9, cnt[0]), flush=True)
