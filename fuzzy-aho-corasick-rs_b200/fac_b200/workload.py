"""Seeded synthetic workloads for the parity tests and bench.py.

The reference ships no data set; its examples generate text on the fly (examples/replace_bench.rs:19-36,
examples/bitap_prototype.rs:128-134).  The generator here is deterministic for a given seed
(numpy PCG64 streams keyed by the seed) and vectorised so that GiB-sized haystacks are produced in
seconds:

  vocabulary  V pseudo-words, length = clip(round(N(6, 2.5)), 2, 14), letters drawn by English frequency
  text        Zipf(s=1) words joined by one space; ". " + capitalised word every ~12 words
              (mixed-case variant: random words fully upper-cased / capitalised)
  planted     every ~`plant_every` bytes a pattern mutated by 0-2 random edits (ins/del/sub/swap)
"""
import numpy as np

LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
FREQ = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5,
                 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
FREQ = FREQ / FREQ.sum()


def make_vocab(seed, size=50000):
    """Returns (words2d uint8 [V, 16] space padded, lengths int32 [V]); distinct words, rank order."""
    rng = np.random.default_rng([seed, 1])
    words, seen = [], set()
    while len(words) < size:
        n = size - len(words)
        lens = np.clip(np.rint(rng.normal(6.0, 2.5, n * 2)), 2, 14).astype(np.int32)
        for ln in lens:
            w = bytes(LETTERS[rng.choice(26, size=int(ln), p=FREQ)])
            if w not in seen:
                seen.add(w)
                words.append(w)
                if len(words) == size:
                    break
    arr = np.full((size, 16), 32, dtype=np.uint8)
    lengths = np.zeros(size, dtype=np.int32)
    for i, w in enumerate(words):
        arr[i, :len(w)] = np.frombuffer(w, dtype=np.uint8)
        lengths[i] = len(w)
    return arr, lengths


def vocab_words(vocab):
    arr, lengths = vocab
    return [bytes(arr[i, :lengths[i]]) for i in range(len(lengths))]


def make_text(seed, nbytes, vocab, mixed_case=False, chunk_words=1 << 20):
    """English-like text of exactly `nbytes` bytes (ASCII)."""
    arr, lengths = vocab
    V = len(lengths)
    rng = np.random.default_rng([seed, 2])
    p = 1.0 / np.arange(1, V + 1)
    cdf = np.cumsum(p / p.sum())
    out = np.empty(nbytes, dtype=np.uint8)
    pos = 0
    while pos < nbytes:
        ids = np.searchsorted(cdf, rng.random(chunk_words)).clip(0, V - 1)
        sentence_end = rng.random(chunk_words) < (1.0 / 12.0)
        wl = lengths[ids] + 1 + sentence_end.astype(np.int32)  # word + ' ' (or '. ')
        ends = np.cumsum(wl)
        total = int(ends[-1])
        starts = ends - wl
        wid = np.repeat(np.arange(chunk_words), wl)
        off = np.arange(total) - np.repeat(starts, wl)
        buf = arr[ids[wid], np.minimum(off, 15)]
        # separators: last byte of each word slot is ' ', the one before is '.' at sentence ends
        wlen_b = np.repeat(lengths[ids], wl)
        is_sep = off >= wlen_b
        buf[is_sep] = 32
        dot = is_sep & np.repeat(sentence_end, wl) & (off == wlen_b)
        buf[dot] = 46
        # capitalise the word after a sentence end
        cap_word = np.zeros(chunk_words, dtype=bool)
        cap_word[1:] = sentence_end[:-1]
        if mixed_case:
            r = rng.random(chunk_words)
            cap_word |= r < 0.10
            upper_word = r > 0.97
            up = np.repeat(upper_word, wl) & ~is_sep
            buf[up] -= 32
        first = np.repeat(cap_word, wl) & (off == 0) & (buf >= 97)
        buf[first] -= 32
        take = min(total, nbytes - pos)
        out[pos:pos + take] = buf[:take]
        pos += take
    return out


def mutate(word, rng, max_edits=2):
    w = bytearray(word)
    for _ in range(int(rng.integers(0, max_edits + 1))):
        if len(w) < 2:
            break
        i = int(rng.integers(0, len(w)))
        op = int(rng.integers(0, 4))
        c = int(LETTERS[rng.integers(0, 26)])
        if op == 0:
            w.insert(i, c)
        elif op == 1:
            del w[i]
        elif op == 2:
            w[i] = c
        elif i + 1 < len(w):
            w[i], w[i + 1] = w[i + 1], w[i]
    return bytes(w)


def plant(text, patterns, seed, every=4096):
    """Overwrite a mutated pattern roughly every `every` bytes (keeps the length of `text`)."""
    rng = np.random.default_rng([seed, 3])
    n = len(text)
    pos = every // 2
    while pos + 64 < n:
        w = b" " + mutate(patterns[int(rng.integers(0, len(patterns)))], rng) + b" "
        text[pos:pos + len(w)] = np.frombuffer(w, dtype=np.uint8)
        pos += every + int(rng.integers(-every // 4, every // 4 + 1))
    return text


def random_words(seed, count, lo, hi):
    rng = np.random.default_rng([seed, 4])
    out, seen = [], set()
    while len(out) < count:
        ln = int(rng.integers(lo, hi + 1))
        w = bytes(LETTERS[rng.choice(26, size=ln, p=FREQ)])
        if w not in seen:
            seen.add(w)
            out.append(w)
    return out


def cfg1(nbytes=64 << 20, seed=0xFAC00001):
    """100 ASCII patterns, edits(1), case-insensitive, threshold 0.8, mixed-case English-like text."""
    vocab = make_vocab(seed)
    words = vocab_words(vocab)
    pats = [w for w in words[200:5000] if len(w) >= 4][:100]
    text = make_text(seed, nbytes, vocab, mixed_case=True)
    return {"patterns": [p.decode() for p in pats], "edits": 1, "case_insensitive": True, "threshold": 0.8,
            "text": text, "name": "cfg1: 100 ASCII patterns, edits(1), ci, thr 0.8"}


def cfg2(nbytes=1 << 30, n_patterns=10000, seed=0xFAC00002):
    """10k ASCII patterns (len 5-16), edits(2), default penalties, threshold 0.8, planted hits every ~4 KiB."""
    vocab = make_vocab(seed)
    words = vocab_words(vocab)
    half = n_patterns // 2
    from_vocab = [w for w in words[1000:] if 5 <= len(w) <= 16][:half]
    rnd = random_words(seed, n_patterns - len(from_vocab), 5, 16)
    pats = from_vocab + rnd
    text = make_text(seed, nbytes, vocab, mixed_case=False)
    text = plant(text, pats, seed)
    return {"patterns": [p.decode() for p in pats], "edits": 2, "case_insensitive": False, "threshold": 0.8,
            "text": text, "name": "cfg2: %d ASCII patterns, edits(2), thr 0.8" % n_patterns}


def cfg4(nbytes=4_000_000_000, n_patterns=100, seed=0xFAC00004, plant_every=1 << 20):
    """Sparse haystack for the bitap pre-filter: 100 random patterns (len 8-20) that do not occur in the
    text vocabulary, per-pattern weights {0.5, 1, 1.5, 2} and per-pattern limits (edits(1), edits(2),
    edits(2).swaps(0), substitutions(1).deletions(1)), one planted fuzzy hit every ~`plant_every` bytes,
    threshold 0.85, sorted().non_overlapping().  Per-pattern limits put the engine on the reference's
    generic MAX_EDITS_FAST=255 path (src/search.rs:205-247)."""
    vocab = make_vocab(seed)
    pats = random_words(seed, n_patterns, 8, 20)
    text = make_text(seed, nbytes, vocab, mixed_case=False)
    text = plant(text, pats, seed, every=plant_every)
    weights = [(0.5, 1.0, 1.5, 2.0)[i % 4] for i in range(n_patterns)]
    limits = [("edits1", "edits2", "edits2_noswap", "sub1_del1")[(i // 4) % 4] for i in range(n_patterns)]
    return {"patterns": [p.decode() for p in pats], "weights": weights, "limit_kinds": limits, "case_insensitive": False,
            "threshold": 0.85, "text": text, "name": "cfg4: %d weighted patterns with per-pattern limits, sparse text" % n_patterns}


def _limits_of(kind):
    from .api import FuzzyLimits
    if kind == "edits1":
        return FuzzyLimits.new().edits(1)
    if kind == "edits2":
        return FuzzyLimits.new().edits(2)
    if kind == "edits2_noswap":
        return FuzzyLimits.new().edits(2).swaps(0)
    return FuzzyLimits.new().substitutions(1).deletions(1)


def build_engine(cfg, backend=None, device=None):
    from .api import FuzzyAhoCorasickBuilder, FuzzyLimits, Pattern
    b = FuzzyAhoCorasickBuilder.new(backend).case_insensitive(cfg["case_insensitive"])
    if "edits" in cfg:
        b = b.fuzzy(FuzzyLimits.new().edits(cfg["edits"]))
    if device is not None:
        b = b.device(device)
    if "limit_kinds" in cfg:
        pats = [Pattern(p, w, _limits_of(k)) for p, w, k in zip(cfg["patterns"], cfg["weights"], cfg["limit_kinds"])]
        return b.build(pats)
    return b.build(cfg["patterns"])
