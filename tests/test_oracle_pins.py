"""Pins for the oracle beyond the KATs: the reference's own prefilter-vs-full differential
(src/prefilter.rs:442-546, same xorshift seeds and draw order) and the Unicode tables that stand
in for unicode-segmentation / to_lowercase (checked against the python `regex` \\X and str.lower)."""
import ctypes as C
import os
import random

import pytest
import regex

from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, FuzzyPenalties, SearchOptions

M64 = (1 << 64) - 1


class Rng:  # prefilter.rs:442-452
    def __init__(self, seed):
        self.s = seed

    def next(self):
        x = self.s
        x ^= (x << 13) & M64
        x ^= x >> 7
        x ^= (x << 17) & M64
        self.s = x
        return x


def _f32(x):
    return C.c_float(x).value


def differential(backend, seed, vocab, filler, trials, check):
    rng = Rng(seed)
    for trial in range(trials):
        npat = 1 + rng.next() % 3
        patterns = [vocab[rng.next() % len(vocab)] for _ in range(npat)]
        edits = rng.next() % 3
        ci = (rng.next() & 1) == 0
        b = FuzzyAhoCorasickBuilder.new(backend).case_insensitive(ci)
        if edits > 0:
            b = b.fuzzy(FuzzyLimits.new().edits(edits))
        if trial % 5 == 0:
            b = b.penalties(FuzzyPenalties.default().swap(0.6).insertion(0.5).deletion(0.8))
        engine = b.build(patterns)
        n = rng.next() % 60
        hay = ""
        for _ in range(n):
            if rng.next() % 7 == 0:
                hay += patterns[rng.next() % len(patterns)] + " "
            else:
                hay += filler[rng.next() % len(filler)]
        thr = _f32(_f32(0.6) + _f32(_f32(rng.next() % 4) * _f32(0.1)))
        check(trial, engine, hay, thr, patterns, edits, ci)


def _pf_check(trial, engine, hay, thr, patterns, edits, ci):
    key = lambda m: (m.start, m.end, m.pattern_index, m.as_tuple()[3], m.edits)
    exp = sorted(key(m) for m in engine.search(hay, SearchOptions.new().threshold(thr)))
    got = sorted(key(m) for m in engine.with_prefilter().search(hay, SearchOptions.new().threshold(thr)))
    assert exp == got, (trial, patterns, edits, ci, thr, hay)


ASCII_VOCAB = ["hello", "world", "vestibulum", "abc", "lorem", "cell"]
ASCII_FILLER = ["a", "b", "c", "d", "e", " ", "1", "o", "0", "l"]
UNI_VOCAB = ["caf\u00e9", "na\u00efve", "\u03a9\u03bc\u03ad\u03b3\u03b1", "\u041c\u043e\u0441\u043a\u0432\u0430", "se\u00f1or", "e\u0301cole"]
UNI_FILLER = ["a", "\u00e9", "\u00f1", "\u03c9", "\u043c", " ", "o", "0", "e\u0301"]


def test_prefilter_matches_full_search_ascii(oracle):  # prefilter.rs:531-536
    differential(oracle, 0x123456789ABCDEF1, ASCII_VOCAB, ASCII_FILLER, 4000, _pf_check)


def test_prefilter_matches_full_search_unicode(oracle):  # prefilter.rs:539-546
    differential(oracle, 0xDEADBEEF0BADF00D, UNI_VOCAB, UNI_FILLER, 4000, _pf_check)


# the same two seeded differentials on the GPU backend (Prefiltered::search == search, both through the C ABI).  Every
# trial builds an engine on the device (~40 ms of cudaMalloc / upload): the default run replays the first 1000 trials of
# each seeded sequence, FAC_TEST_FULL=1 all 4000 (the oracle always runs all 4000 above).
GPU_TRIALS = 4000 if os.environ.get("FAC_TEST_FULL") == "1" else 1000


@pytest.mark.gpu
def test_prefilter_matches_full_search_ascii_gpu(gpu):  # prefilter.rs:531-536
    differential(gpu, 0x123456789ABCDEF1, ASCII_VOCAB, ASCII_FILLER, GPU_TRIALS, _pf_check)


@pytest.mark.gpu
def test_prefilter_matches_full_search_unicode_gpu(gpu):  # prefilter.rs:539-546
    differential(gpu, 0xDEADBEEF0BADF00D, UNI_VOCAB, UNI_FILLER, GPU_TRIALS, _pf_check)


# ---- Unicode tables -----------------------------------------------------------------------------
TRICKY = (["a", "B", " ", "\r", "\n", "\t", "é", "́", "‍", "\U0001F468", "\U0001F469", "\U0001F467",
           "\U0001F1FA", "\U0001F1F8", "\U0001F3FB", "क", "्", "ष", "ि", "؀", "ᄀ",
           "ᅡ", "ᆨ", "가", "각", "é", "ß", "İ", "Σ", "я", "中", "ำ", "ः", "️",
           "©", "❤", "്", "ക", "ᬅ", "᭄"])


def _rand_text(r, n):
    return "".join(r.choice(TRICKY) for _ in range(n))


def test_grapheme_segmentation_matches_regex_X(oracle):
    r = random.Random(1234)
    for _ in range(3000):
        s = _rand_text(r, r.randint(0, 24))
        data = s.encode("utf-8")
        exp, off = [], 0
        for g in regex.findall(r"\X", s):
            exp.append(off)
            off += len(g.encode("utf-8"))
        assert oracle.grapheme_starts(data) == exp, [hex(ord(c)) for c in s]


def test_grapheme_segmentation_all_scalars_pairwise_sample(oracle):
    # every scalar next to a small set of neighbours (covers the whole class table)
    neigh = ["a", "́", "‍", "\U0001F468", "क", "्", "ᄀ", "가", "\U0001F1FA", "\r"]
    r = random.Random(99)
    cps = [c for c in range(0x20, 0x110000, 7) if not (0xD800 <= c <= 0xDFFF)]
    r.shuffle(cps)
    for c in cps[:6000]:
        s = r.choice(neigh) + chr(c) + r.choice(neigh) + chr(c)
        data = s.encode("utf-8")
        exp, off = [], 0
        for g in regex.findall(r"\X", s):
            exp.append(off)
            off += len(g.encode("utf-8"))
        assert oracle.grapheme_starts(data) == exp, [hex(ord(ch)) for ch in s]


def test_lowercase_matches_python_per_scalar(oracle):
    # str::to_lowercase on a single grapheme == per-scalar full lowercase (U+0130 is the only 1:n)
    for c in list(range(0x20, 0x3000)) + list(range(0xA000, 0xAC00, 3)) + list(range(0x10400, 0x10500)) + \
            list(range(0x1E900, 0x1E960)):
        if 0xD800 <= c <= 0xDFFF:
            continue
        ch = chr(c)
        exp = ch.lower()
        if ch == "Σ":
            exp = "σ"
        assert oracle.to_lowercase(ch.encode("utf-8")).decode("utf-8") == exp, hex(c)


def test_utf8_validation(oracle):
    lib = oracle.lib
    good = "aé中\U0001F600".encode()
    assert lib.orc_utf8_valid_up_to(good, len(good)) == len(good)
    for bad, upto in [(b"ab\xff", 2), (b"a\xc3", 1), (b"\xe4\xb8", 0), (b"ok\xed\xa0\x80", 2), (b"\xc0\xaf", 0),
                      (b"\xf4\x90\x80\x80", 0), (b"x\xf0\x9f\x98", 1)]:
        assert lib.orc_utf8_valid_up_to(bad, len(bad)) == upto, bad
