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


def make_text(seed, nbytes, vocab, mixed_case=False, chunk_words=1 << 20, stream=0):
    """English-like text of exactly `nbytes` bytes (ASCII).  `stream` selects an independent PCG64 stream of the
    same seed (block k of a blocked haystack, see cfg2_range)."""
    arr, lengths = vocab
    V = len(lengths)
    rng = np.random.default_rng([seed, 2] + ([stream] if stream else []))
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


def plant(text, patterns, seed, every=4096, stream=0):
    """Overwrite a mutated pattern roughly every `every` bytes (keeps the length of `text`)."""
    rng = np.random.default_rng([seed, 3] + ([stream] if stream else []))
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


CFG2_BLOCK = 32 << 20   # haystacks larger than this are generated block-wise (independent streams per block)


def cfg2_patterns(n_patterns=10000, seed=0xFAC00002, vocab=None):
    vocab = vocab or make_vocab(seed)
    words = vocab_words(vocab)
    half = n_patterns // 2
    from_vocab = [w for w in words[1000:] if 5 <= len(w) <= 16][:half]
    rnd = random_words(seed, n_patterns - len(from_vocab), 5, 16)
    return from_vocab + rnd


def _cfg2_block(args):
    seed, k, size, n_patterns = args
    vocab = make_vocab(seed)
    pats = cfg2_patterns(n_patterns, seed, vocab)
    return plant(make_text(seed, size, vocab, mixed_case=False, stream=k), pats, seed, stream=k)


def cfg2_range(a, b, total=1 << 30, n_patterns=10000, seed=0xFAC00002, procs=1):
    """Bytes [a, b) of the cfg2 haystack of `total` bytes.  A haystack of at most CFG2_BLOCK bytes is one
    make_text + plant stream; a larger one is the concatenation of CFG2_BLOCK-sized blocks, block k drawn from
    stream k of the same seed (block 0 = the unblocked stream), so that a rank can produce its own shard (+ halo)
    without generating the whole haystack and blocks can be produced by a process pool."""
    b = min(b, total)
    if b <= a:
        return np.empty(0, dtype=np.uint8)
    if total <= CFG2_BLOCK:
        blocks = [(seed, 0, total, n_patterns)]
    else:
        blocks = [(seed, k, min(CFG2_BLOCK, total - k * CFG2_BLOCK), n_patterns) for k in range(a // CFG2_BLOCK, (max(b, a + 1) - 1) // CFG2_BLOCK + 1)]
    if procs > 1 and len(blocks) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, len(blocks))) as pool:
            parts = pool.map(_cfg2_block, blocks)
    else:
        parts = [_cfg2_block(x) for x in blocks]
    first = blocks[0][1] * CFG2_BLOCK
    out = np.empty(b - a, dtype=np.uint8)
    pos = first
    for part in parts:
        lo, hi = max(a, pos), min(b, pos + len(part))
        if hi > lo:
            out[lo - a:hi - a] = part[lo - pos:hi - pos]
        pos += len(part)
    return out


def cfg2(nbytes=1 << 30, n_patterns=10000, seed=0xFAC00002, procs=1, text_range=None):
    """10k ASCII patterns (len 5-16), edits(2), default penalties, threshold 0.8, planted hits every ~4 KiB.
    text_range = (a, b): only bytes [a, b) of the nbytes-byte haystack are generated (a rank's shard + halo)."""
    pats = cfg2_patterns(n_patterns, seed)
    a, b = text_range if text_range else (0, nbytes)
    text = cfg2_range(a, b, nbytes, n_patterns, seed, procs)
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


CYR = "абвгдежзийклмнопрстуфхцчшщыьэюя"
CJK = "東京都大阪府北海道日本語中国語漢字文書検索機械学習情報処理技術開発研究計算速度性能評価"
LAT = "abcdefghijklmnopqrstuvwxyz"
COMB = ["e\u0301", "a\u0300", "o\u0308", "u\u0308", "n\u0303", "c\u0327"]   # NFD: base + combining mark = one grapheme
NORDIC = ["straße", "cæsar", "smørbrød", "fußball", "größe", "ærlig", "maße", "æble", "taxi", "boks", "fix", "keks"]


def _uni_word(rng, kind, lo, hi):
    n = int(rng.integers(lo, hi + 1))
    if kind == 0:
        return "".join(CYR[int(i)] for i in rng.integers(0, len(CYR), n))
    if kind == 1:
        return "".join(CJK[int(i)] for i in rng.integers(0, len(CJK), max(2, n // 2)))
    if kind == 2:  # Latin with combining marks
        out = []
        for _ in range(n):
            out.append(COMB[int(rng.integers(0, len(COMB)))] if rng.random() < 0.2 else LAT[int(rng.integers(0, 26))])
        return "".join(out)
    w = NORDIC[int(rng.integers(0, len(NORDIC)))]
    return w + "".join(LAT[int(i)] for i in rng.integers(0, 26, int(rng.integers(0, 4))))


def cfg3(nbytes=256 << 20, n_patterns=1000, seed=0xFAC00003):
    """Unicode workload: Cyrillic / CJK / Latin+combining marks (NFD) / German-Nordic words, case-insensitive,
    mappings ae<->ae-ligature, ss<->sharp-s, ks<->x, edits(2), threshold 0.8.  Text = Zipf words over a 20k-word
    mixed-script vocabulary (some upper-cased), patterns = 1k vocabulary words of 4..12 graphemes; every ~2 KiB a
    pattern is planted with one of its mapped spellings or a character edit.  Returns UTF-8 bytes (numpy uint8)."""
    rng = np.random.default_rng([seed, 1])
    vocab, seen = [], set()
    while len(vocab) < 20000:
        w = _uni_word(rng, int(rng.integers(0, 4)), 3, 12)
        if w not in seen:
            seen.add(w)
            vocab.append(w)
    pats = [w for w in vocab[500:] if 4 <= len(w) <= 14][:n_patterns]
    p = 1.0 / np.arange(1, len(vocab) + 1)
    cdf = np.cumsum(p / p.sum())
    swaps = [("æ", "ae"), ("ß", "ss"), ("x", "ks"), ("ae", "æ"), ("ss", "ß"), ("ks", "x")]
    parts, size = [], 0
    while size < nbytes:
        ids = np.searchsorted(cdf, rng.random(4096)).clip(0, len(vocab) - 1)
        ups = rng.random(4096) < 0.1
        for k, i in enumerate(ids):
            w = vocab[int(i)]
            if k % 97 == 0:  # planted pattern, respelled or edited
                w = pats[int(rng.integers(0, len(pats)))]
                a, b = swaps[int(rng.integers(0, len(swaps)))]
                w = w.replace(a, b, 1) if a in w else (w[:1] + w[2:] if rng.random() < 0.5 and len(w) > 3 else w)
            if ups[k]:
                w = w.upper()
            b_ = (w + " ").encode("utf-8")
            parts.append(b_)
            size += len(b_)
    raw = b"".join(parts)[:nbytes]
    while raw and (raw[-1] & 0xC0) == 0x80:   # do not end inside a scalar
        raw = raw[:-1]
    if raw and raw[-1] >= 0xC0:
        raw = raw[:-1]
    return {"patterns": pats, "edits": 2, "case_insensitive": True, "threshold": 0.8,
            "mappings": [("æ", "ae"), ("ß", "ss"), ("ks", "x")],
            "text": np.frombuffer(raw, dtype=np.uint8).copy(), "name": "cfg3: %d Unicode patterns, ci, mappings, edits(2)" % n_patterns}


class BlockReader:
    """io::Read over a repeating block, handing out at most `max_read` bytes per call and a short read at every
    block end (the contract of examples/streaming.rs:43-82)."""

    def __init__(self, block, total, max_read=64 * 1024):
        self.block, self.total, self.pos, self.max_read = bytes(block), total, 0, max_read

    def read(self, cap=-1):
        if self.pos >= self.total:
            return b""
        off = self.pos % len(self.block)
        n = min(len(self.block) - off, self.total - self.pos, self.max_read, cap if cap and cap > 0 else self.max_read)
        self.pos += n
        return self.block[off:off + n]


def cfg5(total=16 << 30, n_pairs=1000, seed=0xFAC00005, block=1 << 20, auto_beam=(200_000, 100)):
    """Streaming find-and-replace: a Read source repeating a 1 MiB block, 1000 (pattern, replacement) pairs,
    edits(2), case-insensitive, auto_beam(budget, width), threshold 0.8, absolute u64 offsets."""
    vocab = make_vocab(seed)
    words = vocab_words(vocab)
    pats = [w for w in words[1000:] if 5 <= len(w) <= 14][:n_pairs]
    text = plant(make_text(seed, block, vocab, mixed_case=True), pats, seed)
    pairs = [(p.decode(), "<%d>" % i) for i, p in enumerate(pats)]
    return {"pairs": pairs, "edits": 2, "case_insensitive": True, "threshold": 0.8, "auto_beam": auto_beam,
            "block": bytes(text), "total": total, "name": "cfg5: %d-pair replacer, auto_beam%s, streaming" % (n_pairs, auto_beam)}


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
    for a, c in cfg.get("mappings", ()):
        b = b.mapping(a, c)
    if cfg.get("auto_beam"):
        b = b.auto_beam(*cfg["auto_beam"])
    if device is not None:
        b = b.device(device)      # an index, or a list of indices (one replica per device for the stream entry points)
    if "pairs" in cfg:
        return b.build_replacer(cfg["pairs"])
    if "limit_kinds" in cfg:
        pats = [Pattern(p, w, _limits_of(k)) for p, w, k in zip(cfg["patterns"], cfg["weights"], cfg["limit_kinds"])]
        return b.build(pats)
    return b.build(cfg["patterns"])
