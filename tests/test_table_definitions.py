"""CPU check of the builder's productivity / survivor tables against their definition.

The builder (csrc/fac_builder.cpp: build_deep_tables, build_flat_pm) constructs the tables bit-parallel / as wildcard
patterns; tests/emu/fac_emu.cpp re-evaluates the definition written in the builder's header comment cell by cell with plain
recursion.  Engines are small (alphabets of 2-5 symbols) so that every cell of every table is visited.  Result-neutrality of
what the kernels do with the tables is covered by test_emulator_vs_oracle.py / test_gpu_parity.py; this test pins the
tables themselves."""
import ctypes as C
import random

import pytest

from emu_backend import EmuBackend
from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits
from fac_b200._abi import fac_config, fac_pattern


class _Capture:
    """Backend stub: records the C structs build() would hand to fac_engine_create."""
    name = "capture"

    def create(self, cfg, pats, n, device=None):
        self.args = (cfg, pats, n)
        return C.c_void_p(1)

    def free(self, h):
        pass

    def max_match_graphemes(self, h):
        return 0

    def prefilter_active(self, h):
        return False

    def num_nodes(self, h):
        return 0


def _check(fn_name, builder, patterns):
    cap = _Capture()
    builder(cap).build(patterns)
    lib = EmuBackend().lib
    fn = getattr(lib, fn_name)
    fn.argtypes = [C.POINTER(fac_config), C.POINTER(fac_pattern), C.c_size_t, C.POINTER(C.c_uint64)]
    info = (C.c_uint64 * 8)()
    cfg, pats, n = cap.args
    rc = fn(C.byref(cfg), pats, n, info)
    return rc, list(info)


def _words(r, alpha, n, lo, hi):
    n = min(n, sum(len(alpha) ** k for k in range(lo, hi + 1)) // 2 + 1)   # never ask for more words than exist
    out = set()
    while len(out) < n:
        out.add("".join(r.choice(alpha) for _ in range(r.randrange(lo, hi + 1))))
    return sorted(out)


@pytest.mark.parametrize("sizes", [None, ("3", "9", "2"), ("100000", "100000", "100000")])
def test_succinct_deep_tables_equal_their_definition(monkeypatch, sizes):
    if sizes:
        monkeypatch.setenv("FAC_GM3_NODES", sizes[0])
        monkeypatch.setenv("FAC_PM2_NODES", sizes[1])
        monkeypatch.setenv("FAC_PM4_NODES", sizes[2])
    r = random.Random(90210)
    cells = 0
    for t in range(14):
        alpha = r.choice(["ab", "abc", "abcd", "abcde"])
        pats = _words(r, alpha, r.choice([3, 8, 20, 40]), 1, r.choice([3, 5, 7]))
        edits = r.choice([1, 2, 2, 3])
        rc, info = _check("emu_check_deep_tables", lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)), pats)
        assert rc == 0, (rc, pats)
        assert info[1] == 0, ("mismatching cells", info, alpha, pats)
        cells += info[0]
    assert cells > (10000 if sizes and sizes[0] == "3" else 100000)


def test_flat_root_productivity_tables_equal_their_definition(monkeypatch):
    r = random.Random(777)
    seen_k, cells = set(), 0
    for t in range(30):
        alpha = r.choice(["abx", "aeks", "abcsx", "aeksxy"])
        pats = _words(r, alpha, r.choice([3, 8, 20]), 2, r.choice([3, 5, 6]))
        monkeypatch.setenv("FAC_FLAT_ROOT_PM", r.choice(["2", "3", "3"]))
        maps = [("ks", "x"), ("ae", "y")] if "y" in alpha else [("ks", "x")] if "k" in alpha else [("ab", "x")]
        edits = r.choice([2, 2, 3])

        def mk(b, maps=maps, edits=edits):
            bld = FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits))
            for m in maps:
                bld = bld.mapping(*m)
            return bld
        rc, info = _check("emu_check_flat_pm", mk, pats)
        if rc == -3:      # no table (the mapping did not apply to any pattern, or the root has a single edge)
            continue
        assert rc == 0, (rc, pats)
        assert info[1] == 0, ("mismatching cells", info, alpha, pats)
        seen_k.add(info[3])
        cells += info[0]
    assert seen_k == {2, 3} and cells > 20000
