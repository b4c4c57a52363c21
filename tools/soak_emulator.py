#!/usr/bin/env python3
"""One-off CPU soak: the emulator's fast paths (product headers run sequentially, tests/emu) against the oracle on fresh
random engines / haystacks, beyond the fixed seeds of tests/test_emulator_vs_oracle.py.  TEST INFRASTRUCTURE.

    tools/soak_emulator.py succ|flat <first seed> <seconds>

Round 2 (deep survivor / productivity tables, root productivity tables): 4 processes x 15 min, 87 000 cases, no mismatch."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fac_b200 import SearchOptions
from oracle_backend import OracleBackend
from emu_backend import EmuBackend
from fuzzgen import rand_case, rand_dense_case
mode = sys.argv[1]; seed0 = int(sys.argv[2]); budget = float(sys.argv[3])
oracle = OracleBackend()
emu = EmuBackend(tile=16)
if mode == "succ": emu.succinct = True
else: emu.flat = True
t0 = time.time(); n = 0; used = 0
seed = seed0
while time.time() - t0 < budget:
    seed += 1
    r1, r2 = random.Random(seed), random.Random(seed)
    dense = (seed % 3 == 0)
    uni = (mode == "flat" and seed % 2 == 0)
    if dense:
        eo, hay, thr, desc = rand_dense_case(r1, oracle); ee, _, _, _ = rand_dense_case(r2, emu); hay = hay[:300]
    else:
        eo, hay, thr, desc = rand_case(r1, oracle, uni); ee, _, _, _ = rand_case(r2, emu, uni)
    o = eo.search(hay, SearchOptions.new().threshold(thr))
    e = ee.search(hay, SearchOptions.new().threshold(thr))
    if o.tuples() != e.tuples():
        print("MISMATCH seed", seed, desc, flush=True)
        sys.exit(1)
    n += 1
print(mode, "ok", n, "cases; fast path used", emu.succinct_used if mode == "succ" else emu.flat_used, flush=True)
