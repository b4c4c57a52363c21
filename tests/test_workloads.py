"""The seeded workload generators behind the parity tests and bench.py (fac_b200/workload.py)."""
import hashlib
import io

import numpy as np

from fac_b200 import SearchOptions, workload


def _h(b):
    return hashlib.sha256(bytes(b)).hexdigest()[:16]


def test_generators_are_deterministic():
    for fn, n in ((workload.cfg1, 1 << 14), (workload.cfg2, 1 << 14), (workload.cfg3, 1 << 13), (workload.cfg4, 1 << 14)):
        a, b = fn(n), fn(n)
        assert _h(a["text"]) == _h(b["text"]) and a["patterns"] == b["patterns"]
        assert abs(len(a["text"]) - n) < 8
    c = workload.cfg5(total=1 << 16, n_pairs=50, block=1 << 14)
    assert _h(c["block"]) == _h(workload.cfg5(total=1 << 16, n_pairs=50, block=1 << 14)["block"])


def test_cfg3_text_is_valid_utf8_and_mixed_script():
    cfg = workload.cfg3(1 << 14)
    s = bytes(cfg["text"]).decode("utf-8")
    assert any("\u0400" <= ch <= "\u04ff" for ch in s) and any("\u4e00" <= ch <= "\u9fff" for ch in s)
    assert any("\u0300" <= ch <= "\u036f" for ch in s)


def test_block_reader_short_reads():
    r = workload.BlockReader(b"abcdefghij", 25, max_read=4)
    got = []
    while True:
        b = r.read(64)
        if not b:
            break
        got.append(b)
    assert b"".join(got) == (b"abcdefghij" * 3)[:25]
    assert max(len(g) for g in got) <= 4 and any(len(g) < 4 for g in got)


def test_cfg3_and_cfg5_run_on_the_oracle(oracle):
    cfg = workload.cfg3(1 << 12, n_patterns=300)
    e = workload.build_engine(cfg, oracle)
    assert len(e.search(bytes(cfg["text"]), SearchOptions.new().threshold(0.8))) > 10
    c5 = workload.cfg5(total=300_000, n_pairs=100, block=1 << 17, auto_beam=(5000, 20))
    rep = workload.build_engine(c5, oracle)
    out = io.BytesIO()
    rep.replace_stream(workload.BlockReader(c5["block"], c5["total"]), out, 0.8)
    assert b"<" in out.getvalue()
