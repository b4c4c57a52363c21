"""Seeded random engine configurations + haystacks shared by the differential tests."""
import random

from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, FuzzyPenalties, Pattern

ASCII_WORDS = ["hello", "world", "help", "held", "shell", "yellow", "abc", "abcd", "cd", "lorem", "ipsum", "cell",
               "KO", "NA", "MENA", "saddam", "hussein", "vestibulum", "a", "ab", "aab", "abb", "needle", "o0o", "l1l"]
ASCII_FILL = list("abcdehlorsu 01AB\r\n.,")
UNI_WORDS = ["café", "naïve", "Ωμέγα", "Москва",
             "señor", "école", "straße", "strasse", "encyclopædia", "encyclopaedia", "alexandr",
             "aleksandr", "中文字", "İstanbul", "क्ष", "\U0001F468‍\U0001F469",
             "σου", "Γειά"]
UNI_FILL = ["a", "é", "ñ", "ω", "м", " ", "o", "0", "é", "ß", "ss", "æ", "ae", "x",
            "ks", "中", "\r\n", "́", "\U0001F1FA", "İ", "Σ", "E", "K"]


def rand_limits(r):
    k = r.randrange(6)
    if k == 0:
        return FuzzyLimits.new().edits(r.randrange(0, 4))
    if k == 1:
        return FuzzyLimits.new().edits(r.randrange(1, 4)).swaps(0)
    if k == 2:
        return FuzzyLimits.new().substitutions(1).deletions(1)
    if k == 3:
        return FuzzyLimits.new().insertions(r.randrange(0, 3)).deletions(r.randrange(0, 2)).swaps(1)
    if k == 4:
        return FuzzyLimits.new().edits(2).substitutions(1)
    return FuzzyLimits.new()


def rand_case(r, backend, unicode_=False, allow_mappings=True, max_hay=60):
    """Returns (engine, haystack str, threshold, description)."""
    words = UNI_WORDS + ASCII_WORDS[:8] if unicode_ else ASCII_WORDS
    fill = UNI_FILL if unicode_ else ASCII_FILL
    npat = r.randrange(1, 6)
    pats = [r.choice(words) for _ in range(npat)]
    ci = r.random() < 0.5
    b = FuzzyAhoCorasickBuilder.new(backend).case_insensitive(ci)
    mode = r.randrange(5)
    desc = {"pats": pats, "ci": ci, "mode": mode}
    plist = list(pats)
    if mode == 0:
        e = r.randrange(0, 4)
        if e:
            b = b.fuzzy(FuzzyLimits.new().edits(e))
        desc["edits"] = e
    elif mode == 1:
        b = b.fuzzy(rand_limits(r))
    elif mode == 2:  # per-pattern limits / weights
        plist = []
        for p in pats:
            pt = Pattern.from_(p)
            if r.random() < 0.6:
                pt = pt.fuzzy(rand_limits(r))
            if r.random() < 0.4:
                pt = pt.weight(r.choice([0.5, 1.0, 1.5, 2.0]))
            if r.random() < 0.3:
                pt = pt.custom_unique_id(r.randrange(3))
            plist.append(pt)
    elif mode == 3:
        b = b.fuzzy(FuzzyLimits.new().edits(r.randrange(1, 3)))
        if r.random() < 0.5:
            b = b.penalties(FuzzyPenalties.default().swap(0.6).insertion(0.5).deletion(0.8))
        if r.random() < 0.3:
            b = b.min_symbol_similarity(r.choice([0.3, 0.5]))
    else:
        b = b.fuzzy(FuzzyLimits.new().edits(r.randrange(1, 3)))
        if allow_mappings and r.random() < 0.8:
            for a, c in [("æ", "ae"), ("ß", "ss"), ("ks", "x")]:
                if r.random() < 0.7:
                    if r.random() < 0.5:
                        b = b.mapping(a, c)
                    else:
                        b = b.mapping_scored(a, c, r.choice([0.8, 0.5]))
            desc["mappings"] = True
    engine = b.build(plist)
    n = r.randrange(0, max_hay)
    hay = ""
    for _ in range(n):
        if r.randrange(7) == 0:
            w = r.choice(pats)
            if r.random() < 0.5 and len(w) > 1:  # mutate
                i = r.randrange(len(w))
                op = r.randrange(4)
                if op == 0:
                    w = w[:i] + w[i + 1:]
                elif op == 1:
                    w = w[:i] + r.choice(fill) + w[i:]
                elif op == 2:
                    w = w[:i] + r.choice(fill) + w[i + 1:]
                elif i + 1 < len(w):
                    w = w[:i] + w[i + 1] + w[i] + w[i + 2:]
            if r.random() < 0.3:
                w = w.upper()
            hay += w + r.choice(["", " "])
        else:
            hay += r.choice(fill)
    thr = r.choice([0.0, 0.3, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])
    desc["thr"] = thr
    desc["hay"] = hay
    return engine, hay, thr, desc


def rand_dense_case(r, backend):
    """Engines inside the fast kernel's domain (single-byte alphabet <= 31 symbols, edits(1..3)) with a few
    hundred random patterns over a small alphabet, so that the trie is dense, ties are common and outputs pile
    up -- plus random penalties / similarity / weights / case folding.  Returns (engine, haystack, thr, desc)."""
    from fac_b200 import SearchOptions  # noqa: F401
    alpha = r.choice(["abcde", "abcdefghij", "etaoinshrdlu", "abcdefghijklmnopqrstuvwxyz", "ab01-_ .",
                      "abcdefghijklmnopqrstuvwxyz0123456789-_ .",                    # 40 symbols: wide layout
                      "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456"])  # 59 symbols (case-sensitive engines)
    npat = r.choice([20, 60, 200, 500])
    lo, hi = r.choice([(2, 6), (3, 9), (5, 12)])
    pats, seen = [], set()
    while len(pats) < npat:
        w = "".join(r.choice(alpha) for _ in range(r.randrange(lo, hi + 1)))
        if w not in seen:
            seen.add(w)
            pats.append(w)
    edits = r.choice([1, 2, 2, 3])
    ci = r.random() < 0.4
    lim_mode = r.choice([0, 0, 0, 1, 2, 3])   # 0: edits(e) fast path; 1: per-type global; 2: per-pattern limits; 3: no limits
    b = FuzzyAhoCorasickBuilder.new(backend).case_insensitive(ci)
    if lim_mode == 0:
        b = b.fuzzy(FuzzyLimits.new().edits(edits))
    elif lim_mode == 1:
        b = b.fuzzy(r.choice([FuzzyLimits.new().substitutions(1).deletions(1), FuzzyLimits.new().insertions(1).swaps(1),
                              FuzzyLimits.new().edits(2).substitutions(1), FuzzyLimits.new().edits(2).swaps(0)]))
    desc = {"alpha": alpha, "npat": npat, "edits": edits, "ci": ci, "lim_mode": lim_mode}
    if r.random() < 0.4:
        b = b.penalties(FuzzyPenalties.default().swap(r.choice([0.3, 0.6, 0.52])).insertion(r.choice([0.5, 0.25]))
                        .deletion(r.choice([0.8, 0.91, 0.4])).substitution(r.choice([1.0, 1.43, 0.7])))
        desc["pen"] = True
    if r.random() < 0.3:
        b = b.min_symbol_similarity(r.choice([0.3, 0.5]))
        desc["minsym"] = True
    if r.random() < 0.3:
        pairs = {(a, c): r.choice([0.25, 0.5, 0.75]) for a in alpha[:6] for c in alpha[:6] if a != c and r.random() < 0.4}
        b = b.similarity(pairs)
        desc["sim"] = len(pairs)
    plist = pats
    if lim_mode == 2:
        kinds = [FuzzyLimits.new().edits(1), FuzzyLimits.new().edits(2), FuzzyLimits.new().edits(2).swaps(0),
                 FuzzyLimits.new().substitutions(1).deletions(1)]
        plist = [Pattern.from_(p).fuzzy(r.choice(kinds)) if r.random() < 0.8 else p for p in pats]
        if r.random() < 0.5:
            plist = [(q.weight(r.choice([0.75, 1.0, 1.5])) if isinstance(q, Pattern) and r.random() < 0.5 else q) for q in plist]
    elif r.random() < 0.3:
        plist = [Pattern.from_(p).weight(r.choice([0.75, 1.0, 1.25])) if r.random() < 0.5 else p for p in pats]
        desc["weights"] = True
    engine = b.build(plist)
    n = r.choice([200, 1000, 3000])
    hay = []
    while len(hay) < n:
        if r.random() < 0.15:
            w = list(r.choice(pats))
            for _ in range(r.randrange(0, 3)):
                i = r.randrange(len(w))
                op = r.randrange(4)
                if op == 0 and len(w) > 1:
                    del w[i]
                elif op == 1:
                    w.insert(i, r.choice(alpha))
                elif op == 2:
                    w[i] = r.choice(alpha)
                elif i + 1 < len(w):
                    w[i], w[i + 1] = w[i + 1], w[i]
            hay += w
        else:
            hay.append(r.choice(alpha))
    hay = "".join(hay[:n])
    if ci and r.random() < 0.5:
        hay = "".join(c.upper() if r.random() < 0.3 else c for c in hay)
    thr = r.choice([0.5, 0.6, 0.7, 0.8, 0.9])
    desc["thr"] = thr
    return engine, hay, thr, desc
