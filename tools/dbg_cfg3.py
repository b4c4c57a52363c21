import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fac_b200 import GpuBackend, SearchOptions, workload
from oracle_backend import OracleBackend
n = int(os.environ.get("N", 1 << 15))
cfg = workload.cfg3(n, n_patterns=1000)
eo, eg = workload.build_engine(cfg, OracleBackend()), workload.build_engine(cfg, GpuBackend())
text = bytes(cfg["text"])
for name, opts in (("raw", SearchOptions.new().threshold(0.8)), ("sorted_no", SearchOptions.new().threshold(0.8).sorted().non_overlapping()),
                   ("greedy_uni", SearchOptions.new().threshold(0.6).greedy().non_overlapping_unique())):
    o, g = eo.search(text, opts), eg.search(text, opts)
    so, sg = set(o.tuples()), set(g.tuples())
    print(name, len(o), len(g), "only oracle", len(so - sg), "only gpu", len(sg - so), "same order", o.tuples() == g.tuples(),
          "states", o.stats["states_pushed"], g.stats["states_pushed"])
    for t in sorted(so - sg)[:5]:
        print("  O", t, repr(text[t[0]:t[1]].decode("utf-8", "replace")), cfg["patterns"][t[2]])
    for t in sorted(sg - so)[:5]:
        print("  G", t, repr(text[t[0]:t[1]].decode("utf-8", "replace")), cfg["patterns"][t[2]])
