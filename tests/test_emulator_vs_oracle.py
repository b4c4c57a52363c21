"""CPU differential: the product's flattened automaton + slot formulation (run sequentially by
tests/emu) must reproduce the oracle bit-for-bit, including per-window pushed-state counts."""
import random

import pytest

from emu_backend import EmuBackend
from fac_b200 import SearchOptions
from fuzzgen import rand_case, rand_dense_case


@pytest.fixture(scope="module")
def emu():
    return EmuBackend(tile=7)


def _cmp(oracle, emu, seed, unicode_, trials):
    r1, r2 = random.Random(seed), random.Random(seed)
    for t in range(trials):
        eo, hay, thr, desc = rand_case(r1, oracle, unicode_)
        ee, _, _, _ = rand_case(r2, emu, unicode_)
        o = eo.search(hay, SearchOptions.new().threshold(thr))
        e = ee.search(hay, SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, desc)
        assert o.stats["states_pushed"] == e.stats["states_pushed"], (t, desc)
        assert eo.max_match_graphemes() == ee.max_match_graphemes(), (t, desc)
        assert oracle.prefilter_active(eo._h) == emu.prefilter_active(ee._h), (t, desc)


def test_emu_matches_oracle_ascii(oracle, emu):
    _cmp(oracle, emu, 1, False, 1500)


def test_emu_matches_oracle_unicode(oracle, emu):
    _cmp(oracle, emu, 2, True, 1500)


def test_emu_tile_size_independent(oracle):
    r = random.Random(5)
    e1, e2 = EmuBackend(tile=1), EmuBackend(tile=64)
    for t in range(200):
        seed = r.randrange(1 << 30)
        a, hay, thr, desc = rand_case(random.Random(seed), e1, t % 2 == 1)
        b, _, _, _ = rand_case(random.Random(seed), e2, t % 2 == 1)
        o = SearchOptions.new().threshold(thr)
        assert a.search(hay, o).tuples() == b.search(hay, o).tuples(), desc


def test_succinct_fast_path_matches_oracle(oracle):
    """The succinct-trie formulation of the fast kernel (csrc/fac_succinct.h): order-independent
    enumeration + tie detection + faithful redo must equal the oracle wherever it applies."""
    emu = EmuBackend(tile=16)
    emu.succinct = True
    r1, r2 = random.Random(31), random.Random(31)
    for t in range(2500):
        eo, hay, thr, desc = rand_case(r1, oracle, False)
        ee, _, _, _ = rand_case(r2, emu, False)
        o = eo.search(hay, SearchOptions.new().threshold(thr))
        e = ee.search(hay, SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, desc)
    assert emu.succinct_used > 300, emu.succinct_used


def test_succinct_ties(oracle):
    emu = EmuBackend(tile=16)
    emu.succinct = True
    r = random.Random(99)
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits
    for t in range(400):
        pats = ["".join(r.choice("ab") for _ in range(r.randrange(3, 7))) for _ in range(r.randrange(1, 5))]
        hay = "".join(r.choice("ab ") for _ in range(r.randrange(0, 50)))
        edits = r.choice([1, 2, 2, 3])
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).build(pats)
        o = mk(oracle).search(hay, SearchOptions.new().threshold(0.3))
        e = mk(emu).search(hay, SearchOptions.new().threshold(0.3))
        assert o.tuples() == e.tuples(), (t, pats, hay, edits)
    assert emu.succinct_dirty > 0


def test_succinct_dense_tries(oracle):
    emu = EmuBackend(tile=16)
    emu.succinct = True
    r1, r2 = random.Random(177), random.Random(177)
    modes = set()
    for t in range(40):
        eo, hay, thr, desc = rand_dense_case(r1, oracle)
        ee, _, _, _ = rand_dense_case(r2, emu)
        o = eo.search(hay[:400], SearchOptions.new().threshold(thr))
        e = ee.search(hay[:400], SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, desc)
        modes.add(desc["lim_mode"])
    assert emu.succinct_used == 40 and modes == {0, 1, 2, 3}


@pytest.mark.parametrize("sizes", [("0", "0", "0"), ("1", "1", "1"), ("7", "40", "3"), ("60", "300", "12"), ("100000", "100000", "100000")])
def test_succinct_deep_tables_are_result_neutral(oracle, monkeypatch, sizes):
    """The three-deep survivor masks and the 2/3/4-symbol productivity masks (build_deep_tables, fac_succinct.h) only drop
    states that cannot emit: whatever part of the trie they cover (none, the root only, table boundaries inside the trie,
    every node), the result equals the oracle's, and covering more nodes never visits more states."""
    monkeypatch.setenv("FAC_GM3_NODES", sizes[0])
    monkeypatch.setenv("FAC_PM2_NODES", sizes[1])
    monkeypatch.setenv("FAC_PM4_NODES", sizes[2])
    emu = EmuBackend(tile=16)
    emu.succinct = True
    r1, r2 = random.Random(4242), random.Random(4242)
    states = 0
    for t in range(24):
        eo, hay, thr, desc = rand_dense_case(r1, oracle)
        ee, _, _, _ = rand_dense_case(r2, emu)
        o = eo.search(hay[:300], SearchOptions.new().threshold(thr))
        e = ee.search(hay[:300], SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, desc, sizes)
        states += e.stats["states_pushed"]
    assert emu.succinct_used == 24
    _DEEP_STATES[sizes] = states
    if ("0", "0", "0") in _DEEP_STATES and ("100000", "100000", "100000") in _DEEP_STATES:
        assert _DEEP_STATES[("100000", "100000", "100000")] < _DEEP_STATES[("0", "0", "0")]


_DEEP_STATES = {}


def test_succinct_on_unicode_haystacks(oracle):
    """ASCII-alphabet engines on non-ASCII haystacks: the fast formulation reads the K1 first-char stream
    (first-char identity, src/structs.rs:512-519; dead-end filter blind to non-ASCII chars, :471-475)."""
    emu = EmuBackend(tile=16)
    emu.succinct = True
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits
    r = random.Random(515)
    words = ["hello", "world", "help", "cafe", "naive", "resume", "abc", "eclair", "senor", "uber"]
    fill = ["a", "e", "\u00e9", "e\u0301", "\u00f1", " ", "o", "l", "h", "\u4e2d", "\r\n", "\u0301", "E", "\U0001F600", "r", "s"]
    used0 = emu.succinct_used
    for t in range(600):
        pats = r.sample(words, r.randrange(1, 6))
        edits = r.choice([1, 2, 2, 3])
        ci = r.random() < 0.5
        hay = ""
        for _ in range(r.randrange(0, 50)):
            hay += (r.choice(pats) if r.random() < 0.5 else "".join(r.choice(fill) for _ in range(3))) if r.randrange(6) == 0 else r.choice(fill)
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).case_insensitive(ci).build(pats)
        thr = r.choice([0.3, 0.5, 0.7, 0.8])
        o = mk(oracle).search(hay, SearchOptions.new().threshold(thr))
        e = mk(emu).search(hay, SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, pats, edits, ci, thr, hay)
    assert emu.succinct_used - used0 == 600


def test_flat_fast_path_matches_oracle(oracle):
    """The general stack-machine formulation (csrc/fac_flat.h: merged node / edge records, push-time ceiling prune,
    exhausted children walked, order-independent reduction + tie detection + faithful redo) must equal the oracle on
    every fast engine: ASCII and Unicode alphabets, multi-character mappings, non-ASCII haystacks."""
    emu = EmuBackend(tile=16)
    emu.flat = True
    for seed, unicode_, trials in ((71, False, 1200), (72, True, 2000)):
        r1, r2 = random.Random(seed), random.Random(seed)
        for t in range(trials):
            eo, hay, thr, desc = rand_case(r1, oracle, unicode_)
            ee, _, _, _ = rand_case(r2, emu, unicode_)
            o = eo.search(hay, SearchOptions.new().threshold(thr))
            e = ee.search(hay, SearchOptions.new().threshold(thr))
            assert o.tuples() == e.tuples(), (t, desc)
    assert emu.flat_used > 600, emu.flat_used
    # dense tries: branching nodes on their last edit use the per-look-ahead-char survivor masks / output-children lists
    used0 = emu.flat_used
    r1, r2 = random.Random(277), random.Random(277)
    for t in range(60):
        eo, hay, thr, desc = rand_dense_case(r1, oracle)
        ee, _, _, _ = rand_dense_case(r2, emu)
        hay2 = hay[:300] + " \u00e9x\u4e2d " + hay[300:500]
        for h in (hay[:400], hay2):
            o = eo.search(h, SearchOptions.new().threshold(thr))
            e = ee.search(h, SearchOptions.new().threshold(thr))
            assert o.tuples() == e.tuples(), (t, desc)
    assert emu.flat_used - used0 >= 20


def test_flat_fast_path_mappings_and_ties(oracle):
    emu = EmuBackend(tile=16)
    emu.flat = True
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits
    r = random.Random(606)
    words = ["straße", "strasse", "cæsar", "caesar", "taxi", "taksi", "fußball", "æble", "keks", "fix", "maße",
             "москва", "東京都", "café", "éclair"]
    fill = ["a", "e", "s", "ss", "ß", "æ", "ae", "x", "ks", "k", " ", "é", "é", "́", "t", "r", "м", "о", "東", "京", "E", "S"]
    for t in range(700):
        pats = r.sample(words, r.randrange(1, 6))
        edits = r.choice([1, 2, 2, 3])
        ci = r.random() < 0.6
        hay = ""
        for _ in range(r.randrange(0, 40)):
            hay += (r.choice(words) if r.random() < 0.6 else "".join(r.choice(fill) for _ in range(3))) if r.randrange(4) == 0 else r.choice(fill)
        def mk(b):
            bb = FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).case_insensitive(ci)
            return bb.mapping("æ", "ae").mapping("ß", "ss").mapping_scored("ks", "x", 0.8).build(pats)
        thr = r.choice([0.3, 0.5, 0.7, 0.8])
        o = mk(oracle).search(hay, SearchOptions.new().threshold(thr))
        e = mk(emu).search(hay, SearchOptions.new().threshold(thr))
        assert o.tuples() == e.tuples(), (t, pats, edits, ci, thr, hay)
    assert emu.flat_used == 700


def test_flat_root_productivity_masks_are_result_neutral(oracle, monkeypatch):
    """cfg3 (Unicode patterns + mappings, edits(2)): the root productivity masks of the general stack-machine path
    (FlatView::pm_root: two symbols, or three symbols plus the second-level table) cut the visited states and change no
    result."""
    from fac_b200 import workload
    cfg = workload.cfg3(1 << 13)
    text = bytes(cfg["text"])
    opts = SearchOptions.new().threshold(cfg["threshold"])
    o = workload.build_engine(cfg, oracle).search(text, opts)
    assert len(o) > 500
    states = {}
    for flag in ("0", "2", "3"):
        monkeypatch.setenv("FAC_FLAT_ROOT_PM", flag)
        emu = EmuBackend(tile=16)
        emu.flat = True
        e = workload.build_engine(cfg, emu).search(text, opts)
        assert emu.flat_used == 1
        assert o.tuples() == e.tuples(), flag
        states[flag] = e.stats["states_pushed"]
    assert states["2"] * 2 < states["0"] and states["3"] * 2 < states["2"], states


def test_flat_faithful_mode_matches_oracle_push_for_push(oracle):
    """flat_make_ctx<false> / flat_eval_slot<false> (what the shared-memory beamed kernel evaluates states with): the
    order-faithful emulation run on them must reproduce the oracle's matches AND its queue.len() totals -- the
    output-children lists / survivor masks may only skip slots that push nothing, in unchanged order."""
    emu = EmuBackend(tile=5)
    emu.lib.emu_set_faithful_flat(1)
    try:
        for seed, unicode_, trials in ((81, False, 900), (82, True, 900)):
            r1, r2 = random.Random(seed), random.Random(seed)
            for t in range(trials):
                eo, hay, thr, desc = rand_case(r1, oracle, unicode_)
                ee, _, _, _ = rand_case(r2, emu, unicode_)
                o = eo.search(hay, SearchOptions.new().threshold(thr))
                e = ee.search(hay, SearchOptions.new().threshold(thr))
                assert o.tuples() == e.tuples(), (t, desc)
                assert o.stats["states_pushed"] == e.stats["states_pushed"], (t, desc)
        r1, r2 = random.Random(377), random.Random(377)
        for t in range(40):
            eo, hay, thr, desc = rand_dense_case(r1, oracle)
            ee, _, _, _ = rand_dense_case(r2, emu)
            h = hay[:250] + " \u00e9x\u4e2d " + hay[250:350]
            o = eo.search(h, SearchOptions.new().threshold(thr))
            e = ee.search(h, SearchOptions.new().threshold(thr))
            assert o.tuples() == e.tuples(), (t, desc)
            assert o.stats["states_pushed"] == e.stats["states_pushed"], (t, desc)
    finally:
        emu.lib.emu_set_faithful_flat(0)
