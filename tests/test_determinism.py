"""The reference's determinism suite (/root/reference/src/tests.rs:1351-1703): the same engine + haystack searched six
times must give identical match lists (every field, similarity bits included) in every mode -- unsorted, sorted, greedy,
non-overlapping, auto_beam(100, 500), Unicode, pre-filter, streams.  On the GPU backend this is the test that would
catch an atomicAdd-order leak (candidate slots, hash-table representatives, dirty-window redo order)."""
import io

import pytest

from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, SearchOptions

HAYSTACKS = ["hello world", "helo world", "helllo world", "hlelo world", "hwllo world",
             "She sells sea shells by the sea shore", "Why did the yellow bird help the shell?",
             "A quick brown fox jumps over the lazy dog"]


def B(backend):
    return FuzzyAhoCorasickBuilder.new(backend)


def _six_times(fn):
    first = fn()
    for _ in range(5):
        assert fn() == first
    return first


def _modes(thr):
    o = SearchOptions.new().threshold(thr)
    return [o, o.sorted(), o.greedy(), o.sorted().non_overlapping(), o.non_overlapping(), o.coverage_weighted().non_overlapping_unique()]


def test_deterministic_search(backend):  # tests.rs:1351-1453
    e = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(["hello", "world", "help", "held", "shell", "yellow"])
    total = 0
    for hay in HAYSTACKS:
        for thr in (0.5, 0.7, 0.9):
            for opts in _modes(thr):
                total += len(_six_times(lambda: e.search(hay, opts).tuples()))
    assert total > 100


def test_deterministic_search_beam(backend):  # tests.rs:1456-1498
    e = B(backend).fuzzy(FuzzyLimits.new().edits(3)).case_insensitive(True).auto_beam(100, 500).build(
        ["hello", "world", "help", "held", "shell", "yellow", "algorithms", "automaton", "abbreviations"])
    hays = ["hello world", "helo world", "She sells sea shells by the sea shore", "Why did the yellow bird help the shell?",
            "The quick brown fox jumps over the lazy dog", "algorithmic automata and abbreviated forms"]
    total = 0
    for hay in hays:
        for thr in (0.5, 0.7):
            total += len(_six_times(lambda: e.search(hay, SearchOptions.new().threshold(thr).sorted()).tuples()))
    assert total > 10


def test_deterministic_search_unicode(backend):  # tests.rs:1500-1580
    e = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(["café", "résumé", "naïve", "piñata", "jalapeño"])
    hays = ["J'aime le café", "Elle a un joli résumé", "Très naïve attitude", "La piñata est colorée", "Jalapeño poppers",
            "Café au lait avec du sucre", "Un café noir et un résumé clair", "No matches here at all", "Cafe without accent",
            "resume without accent"]
    total = 0
    for hay in hays:
        for thr in (0.5, 0.7, 0.9):
            for opts in _modes(thr)[:4]:
                total += len(_six_times(lambda: e.search(hay, opts).tuples()))
    assert total > 50


def test_deterministic_search_prefilter(backend):  # tests.rs:1582-1630
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["hello", "world", "help", "shell", "yellow"])
    pf = e.with_prefilter()
    assert pf.is_active()
    total = 0
    for hay in HAYSTACKS[:2] + HAYSTACKS[5:]:
        for thr in (0.6, 0.8):
            total += len(_six_times(lambda: pf.search(hay, SearchOptions.new().threshold(thr)).tuples()))
            total += len(_six_times(lambda: pf.search(hay, SearchOptions.new().threshold(thr).sorted().non_overlapping()).tuples()))
    assert total > 10


def test_deterministic_stream(backend):  # tests.rs:1632-1703
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["hello", "world"])
    hay = "hello world hello world"

    def stream_search():
        hits = []
        e.search_stream(io.BytesIO(hay.encode()), 0.8, lambda m: hits.append("%d:%d" % (m.start, m.end)))
        return hits

    def stream_iter():
        return ["%d:%d" % (m.start, m.end) for m in e.stream_matches(io.BytesIO(hay.encode()), 0.8)]

    def replace():
        out = io.BytesIO()
        n = e.replace_stream(io.BytesIO(hay.encode()), out, 0.8, lambda m: "X")
        return out.getvalue().decode(), n

    assert len(_six_times(stream_search)) == 4
    assert _six_times(stream_iter) == stream_search()
    assert _six_times(replace) == ("X X X X", 7)


@pytest.mark.gpu
def test_deterministic_dense_workload_gpu(gpu):
    """Run-to-run identity where races would show: 10k-pattern cfg2 text (hundreds of thousands of candidates through the
    atomic candidate counter and the hash reduction, tie-redo windows included), whole list compared byte for byte."""
    import ctypes as C
    from fac_b200 import workload
    cfg = workload.cfg2(1 << 21, 10000)
    eng = workload.build_engine(cfg, gpu)
    text = bytes(cfg["text"])
    ref = None
    for order, overlap in ((0, 0), (1, 1)):
        first = None
        for _ in range(4):
            arr, _ = gpu.search(eng._h, text, 0.8, order, overlap, False)
            blob = C.string_at(C.addressof(arr), len(arr) * 32) if len(arr) else b""
            if first is None:
                first = blob
            assert blob == first
        assert len(first) > 32 * 1000
