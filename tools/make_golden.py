#!/usr/bin/env python3
"""Generate tests/golden/*.json: small seeded workloads and the oracle's full match lists
(start, end, pattern, similarity bits, ins, del, sub, swap, edits) for them.

The reference is a Rust crate that cannot be built in this image, so these vectors come from the
CPU oracle (oracle/fac_oracle.cpp), which is itself pinned by the reference's known-answer tests
(tests/test_reference_kat.py, tests/test_oracle_pins.py).  They freeze the oracle: the CPU suite
checks the oracle against them, the GPU suite checks libfacgpu.so against them.

    python tools/make_golden.py        (rewrites tests/golden/)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, SearchOptions, workload  # noqa: E402
from oracle_backend import OracleBackend  # noqa: E402

CASES = {
    "cfg1_8KiB": lambda: ("workload", "cfg1", 8192, {}),
    "cfg2_3KiB_2000pat": lambda: ("workload", "cfg2", 3072, {"n_patterns": 2000}),
    "cfg4_64KiB_prefilter": lambda: ("workload", "cfg4", 65536, {"plant_every": 4096}),
}

UNICODE = {
    "patterns": ["straße", "cæsar", "Москва", "東京都", "école", "naïve", "xylophone", "Ωmega"],
    "mappings": [["æ", "ae"], ["ß", "ss"], ["ks", "x"]],
    "edits": 2, "case_insensitive": True, "threshold": 0.7,
    "text": "Die STRASSE nach Moskva: москва, МОСКВА! caesar kam nach 東京都 und 東京; école ÉCOLE naive "
            "xylophone ksylophone Ωmega ωmega strase\r\nstrasse cesaŕ " * 6,
}


def tuples(r):
    return [list(t) for t in r.tuples()]


def main():
    ob = OracleBackend()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, mk in CASES.items():
        _, fn, nbytes, kw = mk()
        cfg = getattr(workload, fn)(nbytes, **kw)
        eng = workload.build_engine(cfg, ob)
        text = bytes(cfg["text"])
        doc = {"generator": {"fn": fn, "nbytes": nbytes, "kwargs": kw}, "threshold": cfg["threshold"], "results": {}}
        for label, opts, pf in (("unsorted_keep", SearchOptions.new().threshold(cfg["threshold"]), False),
                                ("sorted_non_overlapping", SearchOptions.new().threshold(cfg["threshold"]).sorted().non_overlapping(), False),
                                ("prefilter_unsorted", SearchOptions.new().threshold(cfg["threshold"]), True)):
            r = (eng.with_prefilter() if pf else eng).search(text, opts)
            doc["results"][label] = tuples(r)
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, {k: len(v) for k, v in doc["results"].items()})
    b = FuzzyAhoCorasickBuilder.new(ob).fuzzy(FuzzyLimits.new().edits(UNICODE["edits"])).case_insensitive(True)
    for a, c in UNICODE["mappings"]:
        b = b.mapping(a, c)
    eng = b.build(UNICODE["patterns"])
    doc = dict(UNICODE)
    doc["results"] = {"unsorted_keep": tuples(eng.search(UNICODE["text"], SearchOptions.new().threshold(UNICODE["threshold"]))),
                      "greedy_unique": tuples(eng.search(UNICODE["text"], SearchOptions.new().threshold(UNICODE["threshold"]).greedy().non_overlapping_unique()))}
    with open(os.path.join(out_dir, "unicode_mappings.json"), "w") as f:
        json.dump(doc, f, ensure_ascii=True, separators=(",", ":"))
    print("unicode_mappings", {k: len(v) for k, v in doc["results"].items()})
    fast_domain_cases(ob, out_dir)
    headline_case(ob, out_dir)


def headline_case(ob, out_dir):
    """The headline engine at its full pattern count: cfg2 with 10 000 patterns over the first 64 KiB of the haystack,
    the oracle's complete match list (one row per match: start, end, pattern, similarity bits, ins, del, sub, swap,
    edits) as a compressed int64 array."""
    import numpy as np
    cfg = workload.cfg2(65536, 10000)
    eng = workload.build_engine(cfg, ob)
    r = eng.search(bytes(cfg["text"]), SearchOptions.new().threshold(cfg["threshold"]))
    rows = np.array([list(t) for t in r.tuples()], dtype=np.int64)
    np.savez_compressed(os.path.join(out_dir, "cfg2_64KiB_10000pat.npz"), matches=rows)
    print("cfg2_64KiB_10000pat", rows.shape)


def fast_domain_cases(ob, out_dir):
    """Engines inside the fast kernel's extended domain: wide alphabet (> 31 symbols), per-pattern limits, no limits
    at all, and a non-ASCII haystack under an ASCII-alphabet engine."""
    import random
    from fac_b200 import Pattern
    r = random.Random(4242)
    alpha = "abcdefghijklmnopqrstuvwxyz0123456789-_ ."
    pats = sorted({"".join(r.choice(alpha) for _ in range(r.randrange(3, 9))) for _ in range(150)})
    hay = "".join(r.choice(alpha) if r.random() > 0.1 else r.choice(pats) for _ in range(900))
    hay_u = "".join(ch + ("\u00e9" if i % 37 == 5 else "\r\n" if i % 53 == 7 else "") for i, ch in enumerate(hay[:500]))
    kinds = [FuzzyLimits.new().edits(1), FuzzyLimits.new().edits(2).swaps(0), FuzzyLimits.new().substitutions(1).deletions(1)]
    engines = {
        "wide_edits2": FuzzyAhoCorasickBuilder.new(ob).fuzzy(FuzzyLimits.new().edits(2)).build(pats),
        "wide_pattern_limits": FuzzyAhoCorasickBuilder.new(ob).build([Pattern.from_(p).fuzzy(kinds[i % 3]) for i, p in enumerate(pats)]),
        "wide_no_limits": FuzzyAhoCorasickBuilder.new(ob).build(pats),
    }
    doc = {"patterns": pats, "hay": hay, "hay_unicode": hay_u, "threshold": 0.7, "results": {}}
    for name, eng in engines.items():
        for hname, h in (("ascii", hay), ("unicode", hay_u)):
            doc["results"][name + "/" + hname] = tuples(eng.search(h, SearchOptions.new().threshold(0.7)))
    with open(os.path.join(out_dir, "fast_domain.json"), "w") as f:
        json.dump(doc, f, ensure_ascii=True, separators=(",", ":"))
    print("fast_domain", {k: len(v) for k, v in doc["results"].items()})


if __name__ == "__main__":
    main()
