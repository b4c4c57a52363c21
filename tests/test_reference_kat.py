"""Known-answer tests transcribed from the reference's own test-suite
(/root/reference/src/tests.rs, src/prefilter.rs:437-562, doc examples).  Each test cites the
reference test it mirrors.  Parametrised over the CPU oracle (pins the oracle; `-m "not gpu"`)
and the GPU product path through the C ABI (`-m gpu`)."""
import io

import pytest

from fac_b200 import (FuzzyAhoCorasickBuilder, FuzzyLimits, FuzzyPenalties, Pattern, SearchOptions)

BIG_TEXT = ("Lorem ipsum dolor sit amet, consectetur adipiscing elit. Vestibulum eros ipsum, tincidutn eu metus ut, "
            "commodo accumsan mi. Vestibulum porta, orci nec ullamcorper posuere, eros tortor pharetra est, at "
            "porttitor mi leo a velit. Aenean sollicitudin mauris elit, ultricies congue dui vulputate in. In hac "
            "habitasse platea dictumst. Nam iaculis sagittis justo a condimentum. Curabitur sed rhoncus dolor. Lorem "
            "ipsum dolor sit amet, consectetur adipiscing elit. Vivamus egestas congue lorem, in convallis magna "
            "viverra quis. Maecenas fringilla mollis arcu quis maximus. Maecenas tincidunt semper vestibulum. Donec "
            "aliquet leo at molestie elementum. Nulla venenatis iaculis gravida. Phasellus at pulvinar odio. Etiam "
            "bibendum tempor purus at dignissim. Nam a turpis ante. Etiam imperdiet justo sit amet quam tristique "
            "porttitor. Cras ultrices tellus et dolor lobortis tempor. Suspendisse eu mi nec nisi sollicitudin "
            "pharetra. Proin imperdiet elementum ullamcorper. Nam imperdiet quis mi at vulputate. Vivamus pulvinar, "
            "quam et tempus sollicitudin, justo dolor venenatis lacus, sit amet dignissim ex quam ut est. Suspendisse "
            "feugiat libero a augue malesuada sagittis. Curabitur vel magna neque. Praesent eu nulla faucibus, egestas "
            "eros sit amet, elementum quam. Fusce porttitor et lacus vitae maximus. Ut viverra eu sem sed lobortis. "
            "Fusce feugiat vestibulum posuere. Integer erat mauris, tempor eu magna vitae, varius rutrum elit. Proin "
            "mattis, nunc at porta commodo, erat urna viverra ante, vitae feugiat velit dolor ac quam. Nulla semper "
            "elit in neque mollis molestie. Aenean a augue scelerisque, tincidunt odio ut, finibus erat. Integer "
            "feugiat eros ac dolor tempus, sed varius lectus ullamcorper. Orci varius natoque penatibus et magnis dis "
            "parturient montes, nascetur ridiculus mus.")


def B(backend):
    return FuzzyAhoCorasickBuilder.new(backend)


def O():
    return SearchOptions.new()


def has(result, pattern=None, text=None):
    return any((pattern is None or m.pattern.as_str() == pattern) and (text is None or m.text == text) for m in result)


def make_engine(backend):
    return B(backend).fuzzy(FuzzyLimits.new().edits(2)).build(["saddam", "hussein"])


# ---- tests.rs:14-93 -------------------------------------------------------------------------
def test_non_overlapping_regression_0(backend):
    fac = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(["NA", "MENA"])
    r = fac.search("NA MENA", O().threshold(0.6).sorted().non_overlapping())
    assert has(r, "MENA", "MENA")


def test_non_overlapping_regression_2(backend):
    fac = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["KO", "KO", "LWIN"])
    r = fac.search("KWO KO LWIN", O().threshold(0.6).sorted().non_overlapping())
    assert has(r, "KO", "KWO")


def test_non_overlapping_regression_3(backend):
    fac = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(
        ["AL", "WASEL", "AND", "BABEL", "GENERAL", "TRADING", "LLC"])
    r = fac.search("AL WASL ANT BBEL GNERAL TRATING LC", O().threshold(0.6).sorted().non_overlapping_unique())
    assert has(r, "WASEL", "WASL")
    assert has(r, "BABEL", "BBEL")


def test_case_insensitive_ascii(backend):
    e = B(backend).case_insensitive(True).build(["world"])
    r = e.search("HeLlO WoRlD", O().threshold(0.9).sorted())
    assert any(m.text.lower() == "world" for m in r)


def test_unicode_cyrillic(backend):  # tests.rs:98-118
    e = B(backend).case_insensitive(True).build(["юрий"])
    r = e.search("ЮРИЙ ГАГАРИН", O().threshold(0.9).sorted())
    assert any(m.text.lower() == "юрий" for m in r)
    assert e.segment_text("ЮРИЙГАГАРИН", O().threshold(0.9)) == "ЮРИЙ ГАГАРИН"


def test_greek_doctest(backend):  # builder.rs:175-179
    e = B(backend).case_insensitive(True).build([("Γειά", 1.0), ("σου", 1.0)])
    assert not e.search("γειά ΣΟΥ!", O().threshold(0.8).sorted()).is_empty()


# ---- tests.rs:121-228 -----------------------------------------------------------------------
def test_exact_match(backend):
    r = make_engine(backend).search("saddamhussein", O().threshold(0.5).sorted())
    assert has(r, "saddam", "saddam") and has(r, "hussein", "hussein")


def test_extra_letter(backend):
    r = make_engine(backend).search("saddammhussein", O().threshold(0.3).sorted())
    assert has(r, "saddam", "saddam")


def test_missing_letter(backend):
    r = make_engine(backend).search("saddmhussin", O().threshold(0.3).sorted())
    assert has(r, "saddam", "saddm")


def test_substitution(backend):
    r = make_engine(backend).search("saddamhuzein", O().threshold(0.2).sorted())
    assert has(r, "hussein", "huzein")


def test_swap(backend):
    fac = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(["ALI", "KONY"])
    r = fac.search("ALIKOYN", O().threshold(0.6).sorted().non_overlapping())
    assert has(r, "KONY", "KOYN")


def test_big(backend):  # tests.rs:210-228
    fac = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["tincidunt", "porta"])
    r = fac.search(BIG_TEXT, O().threshold(0.8).sorted().non_overlapping())
    assert has(r, text="tincidutn") and has(r, text="tincidunt") and has(r, text="porta")


# ---- tests.rs:231-336 -----------------------------------------------------------------------
def test_overlap_vs_nonoverlap(backend):
    e = B(backend).build([("saddam", 1.0, 2), ("ddamhu", 1.0, 2)])
    m = e.search("saddamddamhu", O().threshold(0.5).sorted())
    assert has(m, "saddam", "saddam") and has(m, "ddamhu", "ddamhu")
    assert len(e.search("saddamhussein", O().threshold(0.7).sorted().non_overlapping())) == 1
    two = e.search("sadam ddamhu", O().threshold(0.4).sorted().non_overlapping())
    assert len(two) == 2
    assert has(two, "saddam", "sadam") and has(two, "ddamhu", "ddamhu")


def test_adjustable_penalties(backend):
    strict = B(backend).build([("hussein", 1.0, 2)]).search("huzein", O().threshold(0.3).sorted())
    assert has(strict, "hussein", "huzein")
    e = B(backend).penalties(FuzzyPenalties.default().substitution(0.8).insertion(0.95).deletion(0.95)).build(
        [("hussein", 1.0, 3)])
    assert has(e.search("huzein", O().threshold(0.2).sorted()), "hussein", "huzein")


def test_regression_1(backend):
    e = B(backend).case_insensitive(True).build(["CO"])
    assert len(e.search("CA", O().threshold(0.8).sorted())) == 0


def test_regression_2(backend):  # tests.rs:339-354
    e = B(backend).build([Pattern.from_("TOLA").fuzzy(FuzzyLimits.new().edits(2))])
    assert has(e.search("TOL", O().threshold(0.5).sorted().non_overlapping()), text="TOL")


# ---- tests.rs:357-423 (exact strings) -----------------------------------------------------------
def test_segment_text(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(3)).build(["saddam", "hussein"])
    assert e.segment_text("sadamhusein", O().threshold(0.8)) == "sadam husein"
    assert e.segment_text("sadamhuseinaltikriti", O().threshold(0.8)) == "sadam husein altikriti"


def test_segment_readme(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).build(["input", "more"])
    m = e.search("someinptandm0re", O().threshold(0.75).sorted().non_overlapping())
    assert m.segment_text() == "some inpt and m0re"


def test_segment_name(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(3)).build(["SHANE", "DOMINIC", "CRAWFORD"])
    assert e.segment_text("SHANEDOM INICCRAWFORD", O().threshold(0.8)) == "SHANE DOM INIC CRAWFORD"


def test_segment_text2(backend):
    e = B(backend).case_insensitive(True).build(["HASAN", "JAMAL", "HUSSEIN", "ZEINIYE"])
    assert e.segment_text("ZEINIYEHussEINHASaNJAMAL", O().threshold(0.8)) == "ZEINIYE HussEIN HASaN JAMAL"


def test_fail(backend):
    e = B(backend).build(["saddam", "hussein"])
    assert e.segment_text("sadam husein", O().threshold(0.8)) == "sadam husein"


# ---- tests.rs:426-576 ---------------------------------------------------------------------------
def test_fuzzy_replace(backend):
    r = B(backend).case_insensitive(True).build_replacer([
        ("PUBLIC JOINT STOCK COMPANY", "PJSC"), ("PUBLIC JOINT STOCK", "PJSC"),
        ("LIMITED LIABILITY COMPANY", "LLC"), ("LIMITED LIABILITY", "LLC")])
    assert r.replace("PUBLIC JOINT STOCK COMPANY GAZPROM", O().threshold(0.8)) == "PJSC GAZPROM"


def test_fuzzy_replace_fn(backend):
    e = B(backend).case_insensitive(True).build(["hair", "bear", "wuzzy"])
    out = e.replace("Fuzzy Wuzzy was a hair. Fuzzy Wuzzy had no bear.", O().threshold(0.8),
                    lambda m: {"bear": "hair", "hair": "bear"}.get(m.text))
    assert out == "Fuzzy Wuzzy was a bear. Fuzzy Wuzzy had no hair."


def test_longer_match_preference(backend):
    e = B(backend).build(["JOINT STOCK COMPANY", "STOCK"])
    r = e.search("JOINT STOCK COMPANY GAZPROM", O().threshold(0.8).sorted().non_overlapping())
    assert has(r, "JOINT STOCK COMPANY") and not has(r, "STOCK")


def test_regression_0(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(2).substitutions(1)).case_insensitive(True).build(["zavod"])
    assert e.search("NARODNY", O().threshold(0.8).sorted().non_overlapping()).is_empty()


def test_readme(backend):
    r = B(backend).fuzzy(FuzzyLimits.new().substitutions(1)).case_insensitive(True).build_replacer(
        [("foo", "bar"), ("baz", "qux")])
    assert r.replace("fo0 and BAZ!", O().threshold(0.7)) == "bar and qux!"


def test_country(backend):
    r = B(backend).fuzzy(FuzzyLimits.new().edits(5)).case_insensitive(True).build_replacer(
        [("CZECHOSLOVAKIA", "SERBIA")])
    assert r.replace("CHEKHOSLOVAKIA", O().threshold(0.7)) == "SERBIA"


def _lorem(backend):
    return B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["LOREM", "IPSUM"])


def test_strip_prefix(backend):
    assert _lorem(backend).strip_prefix("LrEM ISuM Lorm ZZZ", O().threshold(0.8)) == "ZZZ"


def test_strip_postfix(backend):
    assert _lorem(backend).strip_suffix("ZZZ LrEM ISuM Lorm", O().threshold(0.8)) == "ZZZ"


def test_split(backend):
    assert _lorem(backend).split("ZZZLrEMISuMAAA", O().threshold(0.8)) == ["ZZZ", "AAA"]


def test_split_doc(backend):  # query.rs:144-154
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["FOO", "BAR"])
    assert e.split("xxFo0yyBAARzz", O().threshold(0.8)) == ["xx", "yy", "zz"]


def test_doc_matched_spans(backend):  # matches.rs:474-483, 503-510
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["HELLO", "WORLD"])
    m = e.search("helllo wolrd", O().threshold(0.8).sorted().non_overlapping())
    assert m.matched_spans() == [(0, 6), (7, 12)]
    assert m.matched_strings() == ["helllo", "wolrd"]


def test_doc_quickstart(backend):  # query.rs:19-28, builder.rs:12-21
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["hello", "world"])
    found = [m.pattern.as_str() for m in e.search("helllo wolrd", O().threshold(0.8).non_overlapping())]
    assert "hello" in found and "world" in found
    e = B(backend).case_insensitive(True).build(["hello", "world"])
    assert e.segment_text("justheLLowOrLd!", O().threshold(1.0)) == "just heLLo wOrLd!"


def test_doc_replace(backend):  # query.rs:73-85
    e = B(backend).build(["FOO", "BAR", "BAZ"])
    out = e.replace("FOO BAR BAZ", O().threshold(0.8), lambda m: "###" if m.pattern.pattern == "BAR" else None)
    assert out == "FOO ### BAZ"


# ---- tests.rs:578-809 ---------------------------------------------------------------------------
def test_beam_search(backend):
    nb = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(["saddam", "hussein"])
    wb = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).beam_width(100).build(
        ["saddam", "hussein"])
    o = O().threshold(0.7).sorted().non_overlapping()
    assert not nb.search("saddamhusein", o).is_empty()
    r = wb.search("saddamhusein", o)
    assert not r.is_empty() and has(r, "saddam")


def test_truncated_walijan(backend):
    e = B(backend).case_insensitive(True).build([Pattern.from_("WALIJAN").fuzzy(FuzzyLimits.new().edits(3))])
    assert has(e.search("alijan", O().threshold(0.7).sorted()), "WALIJAN")


def test_truncated_short(backend):
    e = B(backend).case_insensitive(True).build([Pattern.from_("TOLA").fuzzy(FuzzyLimits.new().edits(2))])
    assert has(e.search("OLA", O().threshold(0.5).sorted()), text="OLA")


def test_truncated_with_global_limits(backend):
    e = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(2)).build(["TOLA"])
    assert has(e.search("OLA", O().threshold(0.5).sorted()), text="OLA")


def test_truncated_walijan_with_global_limits(backend):
    e = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(3)).build(["WALIJAN"])
    assert has(e.search("alijan", O().threshold(0.7).sorted()), "WALIJAN")


def test_phonetic_td_substitution(backend):
    e = B(backend).case_insensitive(True).build([Pattern.from_("DJAMEL").fuzzy(FuzzyLimits.new().edits(3))])
    r = e.search("Tjamel", O().threshold(0.5).sorted())
    assert has(r, "DJAMEL")
    best = max(m.similarity for m in r)
    assert abs(best - (6 - 0.858) / 6) < 1e-3  # tests.rs:724-728


def test_missing_middle_char(backend):
    e = B(backend).case_insensitive(True).build([Pattern.from_("MOMIR").fuzzy(FuzzyLimits.new().edits(3))])
    r = e.search("Mmir", O().threshold(0.5).sorted())
    assert has(r, "MOMIR")
    assert abs(max(m.similarity for m in r) - (5 - 0.91) / 5) < 1e-3  # tests.rs:753-755


def test_aminullah_aminulah(backend):
    e = B(backend).case_insensitive(True).build([Pattern.from_("AMINULLAH").fuzzy(FuzzyLimits.new().edits(3))])
    assert len(e.search("Aminulah", O().threshold(0.7).sorted())) > 0


def test_long_token_no_blowup_regression(backend):  # tests.rs:816-864
    limits = FuzzyLimits.new().edits(3).substitutions(1).deletions(2).insertions(2).swaps(0)
    pats = ["SA", "LES", "CO", "JSC", "LTD", "BANK", "GROUP", "COMPANY", "CORPORATION", "JOINT STOCK COMPANY",
            "FEDERAL STATE BUDGETARY INSTITUTION OF SCIENCE"]
    e = B(backend).case_insensitive(True).build([Pattern.from_(p).fuzzy(limits) for p in pats])
    r = e.search("RUSSISCHE NATIONALE RUCKVERSICHERUNGSGESELLSCHAFT JSC", O().threshold(0.8).greedy())
    assert has(r, "JSC")


def test_auto_beam(backend):  # tests.rs:866-917
    pats = ["saddam", "hussein", "tincidunt", "porta", "vestibulum", "accumsan"]
    text = "this is a saddamhu example with multiple saddam and tincidutn matches"
    exact = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).build(pats)
    huge = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).auto_beam(2 ** 64 - 1, 8).build(pats)
    o = O().threshold(0.6).sorted()
    assert exact.search(text, o).tuples() == huge.search(text, o).tuples()
    beamed = B(backend).fuzzy(FuzzyLimits.new().edits(2)).case_insensitive(True).auto_beam(1, 16).build(pats)
    assert "saddam" in [m.pattern.as_str() for m in beamed.search(text, o)]


# ---- tests.rs:920-1056 (mappings) ---------------------------------------------------------------
def test_multi_char_mapping_bidirectional(backend):
    ae = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).mapping("æ", "ae").build(
        ["encyclopaedia"])
    m = ae.search("encyclopædia", O().threshold(0.95).sorted())
    assert len(m) == 1 and m[0].substitutions == 1 and m[0].similarity > 0.999
    ea = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).mapping("æ", "ae").build(
        ["encyclopædia"])
    assert len(ea.search("encyclopaedia", O().threshold(0.95).sorted())) == 1


def test_multi_char_mapping_many_to_one(backend):
    mk = lambda p: B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).mapping("ks", "x").build(p)
    assert len(mk(["alexandr"]).search("aleksandr", O().threshold(0.95).sorted())) == 1
    assert len(mk(["aleksandr"]).search("alexandr", O().threshold(0.95).sorted())) == 1


def test_multi_char_mapping_counts_as_edit(backend):
    build = lambda e: B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(e)).mapping("ß", "ss").build(
        ["strasse"])
    assert build(0).search("straße", O().threshold(0.9).sorted()).is_empty()
    assert len(build(1).search("straße", O().threshold(0.9).sorted())) == 1


def test_multi_char_mapping_scored_penalty(backend):
    exact = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).mapping("ks", "x").build(["alexandr"])
    scored = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).mapping_scored("ks", "x", 0.8).build(
        ["alexandr"])
    se = exact.search("aleksandr", O().threshold(0.5).sorted())[0].similarity
    ss = scored.search("aleksandr", O().threshold(0.5).sorted())[0].similarity
    assert se > 0.999 and ss < se


def test_no_mapping_is_unaffected(backend):
    e = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).build(["encyclopaedia"])
    assert e.search("encyclopædia", O().threshold(0.9).sorted()).is_empty()


# ---- tests.rs:1276-1343 -------------------------------------------------------------------------
def test_min_symbol_similarity_floor(backend):
    o = O().threshold(0.8).sorted().non_overlapping()
    no_floor = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["vestibulum"])
    assert len(no_floor.search("vxstibulum", o)) == 1
    floored = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).min_symbol_similarity(0.3).build(
        ["vestibulum"])
    assert floored.search("vxstibulum", o).is_empty()
    assert len(floored.search("vestibulom", o)) == 1
    assert len(floored.search("vestibulum", o)) == 1


# ---- README score example (README.md:61-62) -----------------------------------------------------
def test_prefilter_doc(backend):  # prefilter.rs:101-112
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).build(["vestibulum", "consectetur"])
    pf = e.with_prefilter()
    o = O().threshold(0.85).sorted()
    assert len(pf.search("lorem vestibulm ipsum", o)) == len(e.search("lorem vestibulm ipsum", o)) == 1


def test_prefilter_falls_back_when_not_reducible(backend):  # prefilter.rs:549-561
    assert not B(backend).mapping("ae", "æ").build(["caesar"]).with_prefilter().is_active()
    assert B(backend).fuzzy(FuzzyLimits.new().edits(1)).build(["caesar"]).with_prefilter().is_active()


# ---- streaming: tests.rs:1059-1273 --------------------------------------------------------------
def _needle_input():
    filler = "the quick brown fox " * 50
    s = ""
    while len(s) < 600_000:
        s += filler + "needle "
    return s


def test_streaming_apis_match_whole_input(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["needle"])
    inp = _needle_input()
    truth = sorted((m.start, m.end, m.pattern_index)
                   for m in e.search(inp, O().threshold(0.8).sorted().non_overlapping()))
    assert len(truth) > 300
    got = []
    n = e.search_stream(io.BytesIO(inp.encode()), 0.8, lambda m: got.append((m.start, m.end, m.pattern_index)))
    assert n == len(inp)
    assert sorted(got) == truth
    it = [(m.start, m.end, m.pattern_index) for m in e.stream_matches(io.BytesIO(inp.encode()), 0.8)]
    assert sorted(it) == truth
    par = []
    e.search_stream_parallel(io.BytesIO(inp.encode()), 0.8, 4, lambda m: par.append((m.start, m.end, m.pattern_index)))
    assert sorted(par) == truth


def test_streaming_empty_input(backend):
    e = B(backend).build(["x"])
    hits = []
    assert e.search_stream(io.BytesIO(b""), 0.8, hits.append) == 0 and not hits


def _run_replace(e, text, cb=lambda m: "X"):
    out = io.BytesIO()
    n = e.replace_stream(io.BytesIO(text.encode()), out, 0.8, cb)
    s = out.getvalue().decode()
    assert n == len(out.getvalue())
    return s


def test_replace_stream_small_cases(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["needle"])
    assert _run_replace(e, "a needle b") == "a X b"
    assert _run_replace(e, "needle b") == "X b"
    assert _run_replace(e, "a needle") == "a X"
    assert _run_replace(e, "needle needle") == "X X"
    assert _run_replace(e, "a neeedle b") == "a X b"
    assert _run_replace(e, "nothing here") == "nothing here"
    assert _run_replace(e, "a needle b", lambda m: None) == "a needle b"
    assert _run_replace(e, "") == ""


def test_replace_stream_matches_whole_input(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["needle"])
    inp = _needle_input()
    truth = e.replace(inp, O().threshold(0.8), lambda m: "<%d>" % m.pattern_index)
    streamed = _run_replace(e, inp, lambda m: "<%d>" % m.pattern_index)
    assert streamed == truth and "<0>" in streamed


def test_fuzzy_replacer_replace_stream(backend):
    r = B(backend).case_insensitive(True).fuzzy(FuzzyLimits.new().edits(1)).build_replacer(
        [("hello", "hi"), ("world", "earth")])
    out = io.BytesIO()
    r.replace_stream(io.BytesIO("hell0 w0rld!".encode()), out, 0.8)
    assert out.getvalue().decode() == "hi earth!"


def test_search_stream_doc(backend):  # stream.rs:300-315
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["needle"])
    hits = []
    e.search_stream(io.BytesIO(b"hay neeedle hay"), 0.8, hits.append)
    assert len(hits) == 1 and hits[0].pattern_index == 0


# ---- SURVEY 8a-Q quirks that the reference's behaviour implies --------------------------------------
def test_quirk_suffix_outputs_with_long_span(backend):  # builder.rs:264-268, search.rs:664-678
    e = B(backend).build(["abcd", "cd"])
    spans = sorted((m.start, m.end, m.pattern.as_str()) for m in e.search("xabcdx", O().threshold(0.9)))
    assert (1, 5, "cd") in spans and (3, 5, "cd") in spans and (1, 5, "abcd") in spans


def test_duplicate_patterns_both_reported(backend):  # tests.rs:38-58
    e = B(backend).build(["KO", "KO"])
    r = e.search("KO", O().threshold(0.9))
    assert sorted(m.pattern_index for m in r) == [0, 1]


# ---- tests.rs:758-797 (print-only in the reference: they pin "no panic" and, here, oracle == GPU) ----
@pytest.mark.parametrize("pattern,hay", [("SIMIC", "SIIC"), ("AMINULLAH", "Aminulah"), ("JAFAR", "Jaar")])
def test_missing_letter_edits3(backend, pattern, hay):
    e = B(backend).case_insensitive(True).build([Pattern(pattern).fuzzy(FuzzyLimits.new().edits(3))])
    r = e.search(hay, O().threshold(0.7).sorted())
    assert has(r, pattern)   # one deletion: similarity (n - 0.91) / n >= 0.7 for n >= 4


# ---- tests.rs:1239-1259 ----
def test_replace_stream_parallel_small_cases(backend):
    e = B(backend).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(["needle"])

    def run(inp, threads):
        out = io.BytesIO()
        e.replace_stream_parallel(io.BytesIO(inp.encode()), out, threads, 0.8, lambda m: "X")
        return out.getvalue().decode()
    assert run("a needle b", 8) == "a X b"
    assert run("needle needle", 4) == "X X"
    assert run("a neeedle b", 2) == "a X b"
    assert run("nothing here", 4) == "nothing here"
    assert run("", 4) == ""
