"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.
Bit-exact on (start, end, pattern, similarity bits, ins, del, sub, swap, edits) and on the number of
pushed search states."""
import ctypes as C
import random

import numpy as np
import pytest

from fac_b200 import (FuzzyAhoCorasickBuilder, FuzzyLimits, Order, Overlap, SearchOptions, workload)
from fac_b200._abi import fac_match
from fuzzgen import rand_case, rand_dense_case

pytestmark = pytest.mark.gpu

# FAC_TEST_SEED shifts the seeds of the randomized differentials so that extra GPU runs cover fresh cases
SEED = int(__import__("os").environ.get("FAC_TEST_SEED", "0"))

ALL_OPTS = [(o, v) for o in (Order.Unsorted, Order.Default, Order.Greedy, Order.CoverageWeighted)
            for v in (Overlap.Keep, Overlap.NonOverlapping, Overlap.NonOverlappingUnique)]


def _fuzz(oracle, gpu, seed, unicode_, trials, faithful, monkeypatch):
    # FAC_FAITHFUL=1 forces the order-faithful kernel (its pushed-state count equals the reference's
    # queue.len()); the default FAST kernel walks exhausted chains in place, so only results are compared.
    monkeypatch.setenv("FAC_FAITHFUL", "1" if faithful else "0")
    r1, r2 = random.Random(seed), random.Random(seed)
    ropt = random.Random(seed + 7)
    for t in range(trials):
        eo, hay, thr, desc = rand_case(r1, oracle, unicode_)
        eg, _, _, _ = rand_case(r2, gpu, unicode_)
        order, overlap = ropt.choice(ALL_OPTS)
        opts = SearchOptions(thr, order, overlap)
        o = eo.search(hay, opts)
        g = eg.search(hay, opts)
        assert o.tuples() == g.tuples(), (t, order, overlap, desc)
        if faithful:
            assert o.stats["states_pushed"] == g.stats["states_pushed"], (t, desc)


@pytest.mark.parametrize("faithful", [True, False])
def test_fuzz_ascii(oracle, gpu, faithful, monkeypatch):
    _fuzz(oracle, gpu, 11 + SEED, False, 150 if faithful else 400, faithful, monkeypatch)


@pytest.mark.parametrize("faithful", [True, False])
def test_fuzz_unicode(oracle, gpu, faithful, monkeypatch):
    _fuzz(oracle, gpu, 12 + SEED, True, 400, faithful, monkeypatch)


def test_fast_kernel_on_non_ascii_haystack(oracle, gpu, monkeypatch):
    # ASCII-alphabet engine, haystack with accents / CJK / combining marks / emoji / CRLF sprinkled in: the fast
    # kernel reads the K1 grapheme stream (first chars), byte offsets come from the K1 offset table
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    cfg = workload.cfg2(1 << 14, n_patterns=2000)
    base = bytes(cfg["text"]).decode("ascii")
    r = random.Random(5)
    extra = ["\u00e9", "e\u0301", "\u4e2d\u6587", "\U0001F600", "\r\n", "\u00df", "\u0301", "\u00c9"]
    parts, pos = [], 0
    while pos < len(base):
        step = r.randrange(5, 120)
        parts.append(base[pos:pos + step])
        parts.append(r.choice(extra))
        pos += step
    text = "".join(parts)
    for ci, edits in ((False, 2), (True, 1)):
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).case_insensitive(ci).build(cfg["patterns"])
        eo, eg = mk(oracle), mk(gpu)
        for opts in (SearchOptions.new().threshold(0.8), SearchOptions.new().threshold(0.7).sorted().non_overlapping()):
            o, g = eo.search(text, opts), eg.search(text, opts)
            assert len(o) > 50
            assert o.tuples() == g.tuples()
        monkeypatch.setenv("FAC_FAITHFUL", "1")
        assert mk(gpu).search(text, SearchOptions.new().threshold(0.8)).tuples() == eo.search(text, SearchOptions.new().threshold(0.8)).tuples()
        monkeypatch.setenv("FAC_FAITHFUL", "0")


def test_tile_and_segment_boundaries(oracle, gpu, monkeypatch):
    # tiny tiles and tiny launch segments: matches whose look-ahead crosses tile / segment ends, ragged last tile
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    monkeypatch.setenv("FAC_SUCC_TILE", "32")
    monkeypatch.setenv("FAC_SEGMENT_WINDOWS", "3001")
    cfg = workload.cfg2(10007, n_patterns=1500)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    for n in (10007, 3001, 3002, 33, 32, 31, 5, 1):
        o = eo.search(text[:n], SearchOptions.new().threshold(0.75))
        g = eg.search(text[:n], SearchOptions.new().threshold(0.75))
        assert o.tuples() == g.tuples(), n
    # a pattern at the very end of the haystack, cut short (deletions at end of text) and complete
    tail = text[:200] + cfg["patterns"][3].encode()
    for cut in (0, 1, 2):
        t = tail[:len(tail) - cut]
        assert eo.search(t, SearchOptions.new().threshold(0.6)).tuples() == eg.search(t, SearchOptions.new().threshold(0.6)).tuples(), cut


def test_concurrent_searches_on_one_engine(oracle, gpu):
    # &self search from several host threads (src/stream.rs:395-402): per-call stream + pooled workspace
    import threading
    cfg = workload.cfg2(1 << 16, n_patterns=2000)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    slices = [text[i * 8192:(i + 1) * 8192 + 64] for i in range(8)]
    want = [eo.search(s, SearchOptions.new().threshold(0.8).sorted()).tuples() for s in slices]
    got = [None] * len(slices)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = eg.search(slices[i], SearchOptions.new().threshold(0.8).sorted()).tuples()
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(slices))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs
    assert got == want


def test_exact_engine_fast_path(oracle, gpu, monkeypatch):
    # engines without FuzzyLimits (plain multi-pattern matching): the fast kernel walks the exact chain only
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    cfg = workload.cfg2(1 << 16, n_patterns=3000)
    for ci in (False, True):
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).case_insensitive(ci).build(cfg["patterns"] + ["ab", "b", "the"])
        eo, eg = mk(oracle), mk(gpu)
        text = bytes(cfg["text"]) if not ci else bytes(cfg["text"]).upper()
        for opts in (SearchOptions.new().threshold(0.8), SearchOptions.new().threshold(0.5).sorted().non_overlapping()):
            o, g = eo.search(text, opts), eg.search(text, opts)
            assert len(o) > 100
            assert o.tuples() == g.tuples()


def test_fast_kernel_dense_tries(oracle, gpu, monkeypatch):
    # dense random tries over small alphabets: survivor masks, two-deep masks, walk queue, ties, many outputs
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    r1, r2 = random.Random(77 + SEED), random.Random(77 + SEED)
    ropt = random.Random(78 + SEED)
    for t in range(60):
        eo, hay, thr, desc = rand_dense_case(r1, oracle)
        eg, _, _, _ = rand_dense_case(r2, gpu)
        order, overlap = ropt.choice(ALL_OPTS)
        o = eo.search(hay, SearchOptions(thr, order, overlap))
        g = eg.search(hay, SearchOptions(thr, order, overlap))
        assert o.tuples() == g.tuples(), (t, order, overlap, desc)


def test_fast_kernel_tie_redo(oracle, gpu, monkeypatch):
    # patterns / texts built to tie: sub(sim 0) == ins + del == del + swap at default penalties
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    r = random.Random(99 + SEED)
    letters = "ab"
    for t in range(300):
        pats = ["".join(r.choice(letters) for _ in range(r.randrange(3, 7))) for _ in range(r.randrange(1, 5))]
        hay = "".join(r.choice(letters + " ") for _ in range(r.randrange(0, 50)))
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(2)).build(pats)
        o = mk(oracle).search(hay, SearchOptions.new().threshold(0.3))
        g = mk(gpu).search(hay, SearchOptions.new().threshold(0.3))
        assert o.tuples() == g.tuples(), (t, pats, hay)


def _engines(oracle, gpu, cfg):
    return workload.build_engine(cfg, oracle), workload.build_engine(cfg, gpu)


@pytest.mark.parametrize("faithful", [True, False])
def test_cfg1_slice_parity(oracle, gpu, faithful, monkeypatch):
    monkeypatch.setenv("FAC_FAITHFUL", "1" if faithful else "0")
    cfg = workload.cfg1(1 << 19)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    for opts in (SearchOptions.new().threshold(0.8), SearchOptions.new().threshold(0.8).sorted().non_overlapping()):
        o, g = eo.search(text, opts), eg.search(text, opts)
        assert len(o) > 100
        assert o.tuples() == g.tuples()
        if faithful:
            assert o.stats["states_pushed"] == g.stats["states_pushed"]


@pytest.mark.parametrize("faithful", [True, False])
def test_cfg2_slice_parity(oracle, gpu, faithful, monkeypatch):
    monkeypatch.setenv("FAC_FAITHFUL", "1" if faithful else "0")
    cfg = workload.cfg2(1 << 15, n_patterns=2000)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    o = eo.search(text, SearchOptions.new().threshold(0.8))
    g = eg.search(text, SearchOptions.new().threshold(0.8))
    assert len(o) > 50
    assert o.tuples() == g.tuples()
    if faithful:
        assert o.stats["states_pushed"] == g.stats["states_pushed"]
    for order, overlap in ALL_OPTS[1:]:
        assert eo.search(text, SearchOptions(0.8, order, overlap)).tuples() == \
            eg.search(text, SearchOptions(0.8, order, overlap)).tuples(), (order, overlap)


def test_low_threshold_many_matches(oracle, gpu):
    cfg = workload.cfg2(1 << 12, n_patterns=500)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    o = eo.search(text, SearchOptions.new().threshold(0.0))
    g = eg.search(text, SearchOptions.new().threshold(0.0))
    assert len(o) > 1000
    assert o.tuples() == g.tuples()


def test_shards_with_halo_equal_whole(gpu):
    cfg = workload.cfg1(1 << 18)
    eg = workload.build_engine(cfg, gpu)
    text = bytes(cfg["text"])
    whole = eg.search(text, SearchOptions.new().threshold(0.8)).tuples()
    halo = eg.max_match_graphemes() + 1
    n = len(text)
    cuts = [0, n // 3 + 5, 2 * n // 3 + 1, n]
    got = []
    for a, b in zip(cuts, cuts[1:]):
        end = min(n, b + halo)
        arr, _ = gpu.search_shard(eg._h, text[a:end], end - a, 0, b - a, a, 0.8, False)
        got += [(m.start, m.end, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value,
                 m.insertions, m.deletions, m.substitutions, m.swaps, m.edits) for m in arr]
    assert sorted(got) == sorted(whole)


def test_matches_apply_equals_oracle(oracle, gpu):
    cfg = workload.cfg1(1 << 17)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    raw, _ = gpu.search(eg._h, text, 0.8, 0, 0, False)
    for order, overlap in ALL_OPTS:
        a, _ = gpu.apply(eg._h, raw, len(raw), order, overlap)
        b, _ = oracle.apply(eo._h, raw, len(raw), order, overlap)
        assert bytes(a) == bytes(b), (order, overlap)


def test_device_resident_haystack(gpu):
    import torch
    cfg = workload.cfg1(1 << 18)
    eg = workload.build_engine(cfg, gpu)
    text = bytes(cfg["text"])
    host = eg.search(text, SearchOptions.new().threshold(0.8).sorted()).tuples()
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    arr, stats = gpu.search_device(eg._h, t.data_ptr(), t.numel(), 0.8, 1, 0, False)
    assert [(m.start, m.end, m.pattern_index) for m in arr] == [(x[0], x[1], x[2]) for x in host]
    assert stats["kernel_launches"] > 0


def test_invalid_utf8_and_empty(gpu):
    from fac_b200 import SearchError
    eg = FuzzyAhoCorasickBuilder.new(gpu).fuzzy(FuzzyLimits.new().edits(1)).build(["abc"])
    assert len(eg.search("", SearchOptions.new())) == 0
    with pytest.raises(SearchError) as ei:
        eg.search(b"ab\xffc\xe4", SearchOptions.new())
    assert ei.value.status == 2


def test_tile_failure_retry_path(oracle, gpu, monkeypatch):
    # a tiny queue forces tiles to overflow and exercises the one-window-per-tile retry
    monkeypatch.setenv("FAC_QCAP", "2048")
    monkeypatch.setenv("FAC_TILE", "64")
    monkeypatch.setenv("FAC_FAITHFUL", "1")
    cfg = workload.cfg2(1 << 12, n_patterns=1000)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    o = eo.search(text, SearchOptions.new().threshold(0.8))
    g = eg.search(text, SearchOptions.new().threshold(0.8))
    assert o.tuples() == g.tuples()
    assert o.stats["states_pushed"] == g.stats["states_pushed"]


def test_succinct_stack_overflow_goes_to_faithful_redo(oracle, gpu, monkeypatch):
    # edits(3) with a deliberately tiny warp stack: windows whose top state does not fit are marked
    # dirty by the fast kernel and must come back bit-exact from the order-faithful pass
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    monkeypatch.setenv("FAC_SUCC_STACK", "40")
    cfg = workload.cfg2(1 << 11, n_patterns=1500)
    from fac_b200 import FuzzyLimits as FL
    mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FL.new().edits(3)).build(cfg["patterns"])
    text = bytes(cfg["text"])
    o = mk(oracle).search(text, SearchOptions.new().threshold(0.75))
    g = mk(gpu).search(text, SearchOptions.new().threshold(0.75))
    assert len(o) > 20
    assert o.tuples() == g.tuples()


def test_succinct_deep_tables_result_neutral_on_gpu(oracle, gpu, monkeypatch):
    # the deep survivor / productivity tables (build_deep_tables) with none, partial and default coverage: identical
    # full lists on a 1 MiB cfg2 slice at 10 000 patterns, fewer visited states with more coverage, and the oracle's
    # result on a prefix; several start windows fed at once must not change anything either
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    cfg = workload.cfg2(1 << 20)
    text = bytes(cfg["text"])
    opts = SearchOptions.new().threshold(cfg["threshold"])
    res = {}
    for name, env in (("none", {"FAC_GM3_NODES": "0", "FAC_PM2_NODES": "0", "FAC_PM4_NODES": "0"}),
                      ("partial", {"FAC_GM3_NODES": "20", "FAC_PM2_NODES": "300", "FAC_PM4_NODES": "1", "FAC_SUCC_FEED": "1"}),
                      ("default", {}), ("feed32", {"FAC_SUCC_FEED": "32"}), ("mergesort", {"FAC_RADIX_UNSORTED": "0"})):
        for k in ("FAC_GM3_NODES", "FAC_PM2_NODES", "FAC_PM4_NODES", "FAC_SUCC_FEED", "FAC_RADIX_UNSORTED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g = workload.build_engine(cfg, gpu).search(text, opts)
        res[name] = (g.tuples(), g.stats["states_pushed"])
    assert len(res["none"][0]) > 100000
    assert res["none"][0] == res["partial"][0] == res["default"][0] == res["feed32"][0] == res["mergesort"][0]   # radix == merge order
    assert res["default"][1] < res["partial"][1] < res["none"][1]
    assert res["default"][1] * 3 < res["none"][1]
    n = 1 << 14
    o = workload.build_engine(cfg, oracle).search(text[:n], opts)
    g = workload.build_engine(cfg, gpu).search(text[:n], opts)
    assert o.tuples() == g.tuples()


@pytest.mark.parametrize("edits", [1, 3, 4])
def test_succinct_other_edit_budgets(oracle, gpu, edits, monkeypatch):
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    cfg = workload.cfg2(1 << 11 if edits > 2 else 1 << 14, n_patterns=600)
    from fac_b200 import FuzzyLimits as FL
    mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FL.new().edits(edits)).case_insensitive(True).build(cfg["patterns"])
    text = bytes(cfg["text"])
    thr = 0.8 if edits < 3 else 0.7
    o = mk(oracle).search(text, SearchOptions.new().threshold(thr))
    g = mk(gpu).search(text, SearchOptions.new().threshold(thr))
    assert len(o) > 10
    assert o.tuples() == g.tuples()


def test_prefilter_fuzz_ascii(oracle, gpu, monkeypatch):
    # Prefiltered::search (prefilter.rs:135): GPU bitap scan + slices vs the oracle's restatement
    r1, r2 = random.Random(41 + SEED), random.Random(41 + SEED)
    ropt = random.Random(48 + SEED)
    used = 0
    for t in range(400):
        eo, hay, thr, desc = rand_case(r1, oracle, False)
        eg, _, _, _ = rand_case(r2, gpu, False)
        order, overlap = ropt.choice(ALL_OPTS)
        opts = SearchOptions(thr, order, overlap)
        assert eo.with_prefilter().is_active() == eg.with_prefilter().is_active(), (t, desc)
        used += eo.with_prefilter().is_active()
        o = eo.with_prefilter().search(hay, opts)
        g = eg.with_prefilter().search(hay, opts)
        assert o.tuples() == g.tuples(), (t, order, overlap, desc)
    assert used > 50


def test_prefilter_non_ascii_haystack_per_slice_is_ascii(oracle, gpu):
    # K2 on a non-ASCII haystack: the bitap scan runs over K1's symbol-id stream (transcode, prefilter.rs:262-281) and
    # every merged slice is searched as its own haystack with its own is_ascii test (prefilter.rs:346-350, quirk Q8):
    # the all-ASCII slice around "abcd\r\nefgh" uses byte graphemes (CR, LF = two insertions), while the plain search
    # of the whole (non-ASCII) haystack sees one CR LF cluster (one insertion).  The reference is not self-consistent
    # here and neither are we: pre-filtered GPU == pre-filtered oracle != plain search.
    mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(2)).build(["abcdefgh", "zebra"])
    eo, eg = mk(oracle), mk(gpu)
    text = "lorem ipsum abcd\r\nefgh dolor sit amet " + "x" * 40 + " caf\u00e9 zebra"
    opts = SearchOptions.new().threshold(0.7)
    assert eg.with_prefilter().is_active()
    o, g = eo.with_prefilter().search(text, opts), eg.with_prefilter().search(text, opts)
    assert o.tuples() == g.tuples()
    assert any(m.insertions == 2 for m in g)                      # the ASCII-storage reading of the slice
    plain = eg.search(text, opts)
    assert plain.tuples() == eo.search(text, opts).tuples()
    assert plain.tuples() != g.tuples()                           # i.e. the device pre-filter really ran
    # randomized: ASCII-alphabet and Unicode engines over haystacks sprinkled with accents, CJK, CR LF, emoji
    r = random.Random(77 + SEED)
    extra = ["\u00e9", "e\u0301", "\u4e2d\u6587", "\U0001F600", "\r\n", "\u00df", "\r\n\r\n", "\u00c9", "\u043c\u043e\u0441\u043a\u0432\u0430"]
    words = ["vestibulum", "consectetur", "moskva", "\u043c\u043e\u0441\u043a\u0432\u0430", "na\u00efve", "stra\u00dfe", "lorem", "ipsum", "dolor"]
    used = 0
    for t in range(60):
        pats = r.sample(words, r.randrange(1, 5))
        ci = r.random() < 0.5
        edits = r.choice([1, 1, 2])
        mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).case_insensitive(ci).build(pats)
        eo, eg = mk(oracle), mk(gpu)
        parts = []
        for _ in range(r.randrange(5, 60)):
            w = r.choice(words + ["foo", "bar", "quux", "zzz", "amet"])
            if r.random() < 0.3 and len(w) > 3:
                k = r.randrange(1, len(w) - 1)
                w = r.choice([w[:k] + w[k + 1:], w[:k] + "\r\n" + w[k:], w[:k] + "q" + w[k:], w.upper()])
            parts.append(w)
            parts.append(r.choice([" ", " ", "  ", "\r\n", ", "]) if r.random() < 0.8 else r.choice(extra))
        hay = "".join(parts) + r.choice(extra)
        thr = r.choice([0.6, 0.7, 0.8, 0.9])
        order, overlap = r.choice(ALL_OPTS)
        opts = SearchOptions(thr, order, overlap)
        assert eo.with_prefilter().is_active() == eg.with_prefilter().is_active()
        used += eg.with_prefilter().is_active()
        assert eo.with_prefilter().search(hay, opts).tuples() == eg.with_prefilter().search(hay, opts).tuples(), (t, pats, ci, edits, thr, hay)
    assert used > 30


def test_non_ascii_haystack_beyond_4GiB_bytes(gpu, oracle):
    # the reference's limit is graphemes (u32 positions, search.rs:198-202, 296-300), not bytes: a 4 GiB + 64 MiB UTF-8
    # haystack of two-byte scalars (2.2 G graphemes) is segmented with 64-bit byte offsets and searched; matches planted
    # beyond 2^32 bytes come back with their absolute offsets
    import torch
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 32) + (64 << 20)
    if free < 80 * (1 << 30):
        pytest.skip("needs ~60 GiB of device memory for the grapheme streams of a 4 GiB haystack")
    pats = ["vestibulum", "tincidunt"]
    mk = lambda b: FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(1)).build(pats)
    eg, eo = mk(gpu), mk(oracle)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    dev[0::2] = 0xC3
    dev[1::2] = 0xA9                       # "\u00e9" repeated
    plants = [(1000, "caf\u00e9 vestibulum "), ((1 << 32) - 8, " vestibulm t\u00efncidunt "), ((1 << 32) + (48 << 20) + 4, " \u4e2d tincidunt\r\n")]
    for pos, txt in plants:
        b = txt.encode("utf-8")
        if len(b) % 2:
            b += b" "
        dev[pos:pos + len(b)] = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
    arr, st = gpu.search_device(eg._h, dev.data_ptr(), n, 0.8, 0, 0, False)
    got = sorted((m.start, m.end, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value, m.insertions, m.deletions,
                  m.substitutions, m.swaps) for m in arr)
    want = []
    for pos, txt in plants:
        lo = pos - 32
        sl = bytes(dev[lo:pos + 96].cpu().numpy())
        o, _ = oracle.search(eo._h, sl, 0.8, 0, 0, False)
        want += [(m.start + lo, m.end + lo, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value, m.insertions, m.deletions,
                  m.substitutions, m.swaps) for m in o]
    assert len(want) >= 4 and max(w[0] for w in want) > (1 << 32)
    assert got == sorted(want)
    del dev


@pytest.mark.parametrize("faithful", [True, False])
def test_cfg1_prefilter_parity(oracle, gpu, faithful, monkeypatch):
    monkeypatch.setenv("FAC_FAITHFUL", "1" if faithful else "0")
    cfg = workload.cfg1(1 << 17)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    for opts in (SearchOptions.new().threshold(0.8), SearchOptions.new().threshold(0.8).sorted().non_overlapping()):
        o, g = eo.with_prefilter().search(text, opts), eg.with_prefilter().search(text, opts)
        assert len(o) > 50
        assert o.tuples() == g.tuples()
        assert g.tuples() == eg.search(text, opts).tuples()   # the reference's own differential (prefilter.rs:465-546)


def test_cfg4_sparse_prefilter_parity(oracle, gpu):
    cfg = workload.cfg4(1 << 20, plant_every=1 << 13)
    eo, eg = _engines(oracle, gpu, cfg)
    assert eg.with_prefilter().is_active()
    text = bytes(cfg["text"])
    opts = SearchOptions.new().threshold(0.85).sorted().non_overlapping()
    o = eo.with_prefilter().search(text, opts)
    g = eg.with_prefilter().search(text, opts)
    assert len(o) > 20
    assert o.tuples() == g.tuples()
    raw_o = eo.with_prefilter().search(text, SearchOptions.new().threshold(0.85))
    raw_g = eg.with_prefilter().search(text, SearchOptions.new().threshold(0.85))
    assert raw_o.tuples() == raw_g.tuples()


def test_cfg3_unicode_mappings_slice_parity(oracle, gpu, monkeypatch):
    # BASELINE config 3: Cyrillic / CJK / combining marks, case-insensitive, mappings, edits(2)
    cfg = workload.cfg3(1 << 15, n_patterns=1000)
    eo, eg = _engines(oracle, gpu, cfg)
    text = bytes(cfg["text"])
    for opts in (SearchOptions.new().threshold(0.8), SearchOptions.new().threshold(0.8).sorted().non_overlapping(),
                 SearchOptions.new().threshold(0.6).greedy().non_overlapping_unique()):
        o, g = eo.search(text, opts), eg.search(text, opts)
        assert len(o) > 50
        assert o.tuples() == g.tuples()
    # the order-faithful kernel also reproduces the reference's queue.len() total
    monkeypatch.setenv("FAC_FAITHFUL", "1")
    ef = workload.build_engine(cfg, gpu)
    fo, fg = eo.search(text, SearchOptions.new().threshold(0.8)), ef.search(text, SearchOptions.new().threshold(0.8))
    assert fo.tuples() == fg.tuples() and fo.stats["states_pushed"] == fg.stats["states_pushed"]


def test_cfg3_4MiB_fast_equals_faithful_full_list(oracle, gpu, monkeypatch):
    # cfg3 at 4 MiB: the stack-machine kernel with the root productivity masks, the same kernel without them and the
    # order-faithful kernel return the same full list (bit-exact records); the oracle agrees on a 48 KiB slice
    from bench_configs import _cut_utf8
    cfg = workload.cfg3(4 << 20, n_patterns=1000)
    text = bytes(cfg["text"])
    opts = SearchOptions.new().threshold(0.8)
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    fast = workload.build_engine(cfg, gpu).search(text, opts)
    monkeypatch.setenv("FAC_FLAT_ROOT_PM", "0")
    nopm = workload.build_engine(cfg, gpu).search(text, opts)
    monkeypatch.delenv("FAC_FLAT_ROOT_PM")
    monkeypatch.setenv("FAC_FAITHFUL", "1")
    faithful = workload.build_engine(cfg, gpu).search(text, opts)
    assert len(fast) > 100000
    assert fast.tuples() == nopm.tuples() == faithful.tuples()
    assert fast.stats["states_pushed"] * 2 < nopm.stats["states_pushed"]
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    piece = _cut_utf8(text[2 << 20:(2 << 20) + (48 << 10)])
    o = workload.build_engine(cfg, oracle).search(piece, opts)
    g = workload.build_engine(cfg, gpu).search(piece, opts)
    assert len(o) > 1000 and o.tuples() == g.tuples()


def test_cfg5_streaming_replacer_parity(oracle, gpu):
    # BASELINE config 5 in miniature: repeating-block Read source, auto_beam, FuzzyReplacer::replace_stream,
    # absolute offsets; three 256 KiB windows
    import io
    cfg = workload.cfg5(total=700_000, n_pairs=200, block=1 << 18, auto_beam=(20_000, 50))
    ro, rg = workload.build_engine(cfg, oracle), workload.build_engine(cfg, gpu)
    out_o, out_g = io.BytesIO(), io.BytesIO()
    ro.replace_stream(workload.BlockReader(cfg["block"], cfg["total"]), out_o, 0.8)
    rg.replace_stream(workload.BlockReader(cfg["block"], cfg["total"]), out_g, 0.8)
    assert out_o.getvalue() == out_g.getvalue()
    assert b"<" in out_g.getvalue() and len(out_g.getvalue()) != cfg["total"]
    mo, mg = [], []
    ro.engine().search_stream(workload.BlockReader(cfg["block"], cfg["total"]), 0.8, lambda m: mo.append(m.as_tuple()))
    rg.engine().search_stream(workload.BlockReader(cfg["block"], cfg["total"]), 0.8, lambda m: mg.append(m.as_tuple()))
    assert mo == mg and len(mo) > 100
    assert max(t[0] for t in mg) > (1 << 18)   # absolute offsets beyond the first window


MATCH_DT = np.dtype([("start", "<u8"), ("end", "<u8"), ("pat", "<u4"), ("simbits", "<u4"), ("ins", "u1"), ("del", "u1"),
                     ("sub", "u1"), ("swp", "u1"), ("edits", "u1"), ("pad", "u1", (3,))])


def test_cfg2_32MiB_properties_and_sampled_parity(oracle, gpu):
    """BASELINE-size behaviour through size-independent properties plus oracle parity on sampled regions:
    the oracle needs ~4 s per KiB on this configuration, so it checks 150 random 384-byte regions
    (region + halo searched as its own haystack, ownership by start -- the reference's own streaming rule,
    src/stream.rs:262-297) against the GPU's whole-haystack result."""
    nbytes = 32 << 20
    cfg = workload.cfg2(nbytes, 10000)
    eo, eg = _engines(oracle, gpu, cfg)
    text = cfg["text"]
    arr, stats = gpu.search_host_ptr(eg._h, text.ctypes.data, nbytes, 0.8, 0, 0, False)
    m = np.frombuffer(arr, dtype=MATCH_DT)
    assert len(m) > 4_000_000
    # properties: keys unique and ascending (Unsorted == ascending (start, end, pattern)), bounds, limits, threshold
    key_lt = (m["start"][:-1] < m["start"][1:]) | ((m["start"][:-1] == m["start"][1:]) & (
        (m["end"][:-1] < m["end"][1:]) | ((m["end"][:-1] == m["end"][1:]) & (m["pat"][:-1] < m["pat"][1:]))))
    assert key_lt.all()
    assert (m["end"] <= nbytes).all() and (m["start"] <= m["end"]).all()
    assert (m["end"] - m["start"] <= eg.max_match_graphemes()).all()
    assert (m["edits"] <= 2).all() and (m["ins"].astype(int) + m["del"] + m["sub"] + m["swp"] == m["edits"]).all()
    sim = m["simbits"].view("<f4")
    assert (sim >= np.float32(0.8)).all() and (sim <= np.float32(1.0)).all()
    # sampled parity
    rng = np.random.default_rng(7)
    halo = eg.max_match_graphemes() + 1
    region = 384
    for p in [0, nbytes - region] + [int(x) for x in rng.integers(0, nbytes - region, 148)]:
        end = min(nbytes, p + region + halo)
        o, _ = oracle.search(eo._h, bytes(text[p:end]), 0.8, 0, 0, False)
        want = sorted((x.start + p, x.end + p, x.pattern_index, C.c_uint32.from_buffer(C.c_float(x.similarity)).value,
                       x.insertions, x.deletions, x.substitutions, x.swaps, x.edits) for x in o if x.start < region)
        lo, hi = np.searchsorted(m["start"], [p, p + region])
        got = sorted((int(r["start"]), int(r["end"]), int(r["pat"]), int(r["simbits"]), int(r["ins"]), int(r["del"]),
                      int(r["sub"]), int(r["swp"]), int(r["edits"])) for r in m[lo:hi])
        assert got == want, p


def test_cfg2_64MiB_fast_equals_faithful_full_list(gpu, monkeypatch):
    """The quoted configuration at size: cfg2 with all 10 000 patterns over 64 MiB, the fast kernel's COMPLETE match
    list (11 M records) byte for byte against the order-faithful kernel (FAC_FAITHFUL=1: FIFO order, per-level dedup,
    pushed-state count == the reference's queue.len()) -- two independent device paths, one of which the oracle pins at
    small sizes with exactly this engine (tests/golden/cfg2_64KiB_10000pat.npz)."""
    import torch
    nbytes = 64 << 20
    cfg = workload.cfg2(nbytes, 10000)
    dev = torch.from_numpy(cfg["text"]).cuda()
    monkeypatch.setenv("FAC_FAITHFUL", "0")
    fast = workload.build_engine(cfg, gpu)
    monkeypatch.setenv("FAC_FAITHFUL", "1")
    faithful = workload.build_engine(cfg, gpu)
    a, sa = gpu.search_device(fast._h, dev.data_ptr(), nbytes, 0.8, 0, 0, False)
    b, sb = gpu.search_device(faithful._h, dev.data_ptr(), nbytes, 0.8, 0, 0, False)
    assert len(a) == len(b) and len(a) > 10_000_000
    assert np.array_equal(np.frombuffer(a, dtype=np.uint8), np.frombuffer(b, dtype=np.uint8))
    # the faithful kernel counts the reference's pushed states, the fast kernel visits far fewer
    assert sb["states_pushed"] > 2000 * nbytes and sa["states_pushed"] < sb["states_pushed"] // 3


def _beam_cases(seed, trials):
    r = random.Random(seed)
    words = ["saddam", "hussein", "tincidunt", "porta", "vestibulum", "accumsan", "hello", "world", "help", "shell",
             "yellow", "abc", "abcd", "needle"]
    fill = list("abcdehlorstu   ")
    for t in range(trials):
        pats = r.sample(words, r.randrange(1, 7))
        edits = r.randrange(1, 5)
        kind = r.randrange(3)
        hay = ""
        for _ in range(r.randrange(0, 80)):
            hay += (r.choice(pats) + " ") if r.randrange(5) == 0 else r.choice(fill)
        thr = r.choice([0.3, 0.5, 0.6, 0.7, 0.8])
        yield t, pats, edits, kind, (r.choice([1, 2, 3, 8, 16, 100]), r.choice([1, 50, 500, 5000, 10 ** 9])), hay, thr


def test_beam_and_auto_beam_parity(oracle, gpu):
    for t, pats, edits, kind, (bw, budget), hay, thr in _beam_cases(21, 300):
        def mk(b):
            bb = FuzzyAhoCorasickBuilder.new(b).fuzzy(FuzzyLimits.new().edits(edits)).case_insensitive(True)
            if kind == 0:
                bb = bb.beam_width(bw)
            elif kind == 1:
                bb = bb.auto_beam(budget, bw)
            else:
                bb = bb.beam_width(bw).auto_beam(budget, 7)  # an explicit beam wins (search.rs:527)
            return bb.build(pats)
        o = mk(oracle).search(hay, SearchOptions.new().threshold(thr))
        g = mk(gpu).search(hay, SearchOptions.new().threshold(thr))
        assert o.tuples() == g.tuples(), (t, pats, edits, kind, bw, budget, thr, hay)
        assert o.stats["states_pushed"] == g.stats["states_pushed"], (t, pats, edits, kind, bw, budget, thr, hay)
