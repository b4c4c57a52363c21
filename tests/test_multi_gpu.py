"""N>1 host logic (shard plan on grapheme-cluster boundaries, halo, ownership, exact-size gather, global apply):
world sizes 2..7 over gloo on the CPU with the oracle standing in for the per-rank searcher, ASCII and non-ASCII
haystacks with patterns straddling every cut; and the same flow on real GPUs (marked gpu): NCCL when the box has
two devices, and a single-device replay of the rank loop through fac_search_ex / fac_matches_apply_device."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))


def _tuples(arr):
    return [(m.start, m.end, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value, m.insertions,
             m.deletions, m.substitutions, m.swaps, m.edits) for m in arr]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


UNI_PATTERNS = ["éàüöñçéàüö", "москва", "straße", "école", "東京都", "café au lait", "naïve", "abc\r\ndef", "🇩🇪🇫🇷"]


def _unicode_case(kind):
    """(builder kwargs, patterns, text bytes): non-ASCII haystacks whose pattern occurrences sit on every byte offset
    a cut can take, plus long pure-ASCII stretches (an ASCII slice of a non-ASCII haystack)."""
    if kind == "dense":
        text = ("xéàüöñçéàüöx москва straße écolé école 東京都 café au lait naïve 🇩🇪🇫🇷🇩🇪 " * 9)
        text += "abc\r\ndef plain ascii filler with abc\r\ndef and more abc\r\ndfe text " * 6
        text += "éàüöñçeàüö москвa strasse 東京 " * 5
    else:
        text = "éàüöñçéàüö" * 7 + "x" + "éàüöñçéàüö" * 5
    return UNI_PATTERNS, text.encode("utf-8")


def _worker(rank, world, port, backend_name, case, q):
    sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, sharding, workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend_name == "gpu":
        from fac_b200 import GpuBackend
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank % torch.cuda.device_count()))
        be, dev = GpuBackend(), "cuda"
    else:
        from oracle_backend import OracleBackend
        dist.init_process_group("gloo", rank=rank, world_size=world)
        be, dev = OracleBackend(), "cpu"
    device = (rank % torch.cuda.device_count()) if backend_name == "gpu" else None
    if case == "cfg1":
        cfg = workload.cfg1(1 << 15)
        eng = workload.build_engine(cfg, be, device=device)
        text, thr = cfg["text"], 0.8
    else:
        pats, text = _unicode_case(case)
        b = FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True)
        if device is not None:
            b = b.device(device)
        eng = b.build(pats)
        thr = 0.7
    res = {}
    for order, overlap in ((0, 0), (1, 1), (2, 2), (3, 0)):
        final = sharding.search_sharded(eng, be, text, thr, order, overlap, dist, dev)
        if rank == 0:
            whole, _ = be.search(eng._h, bytes(text), thr, order, overlap, False)
            res[(order, overlap)] = (_tuples(final), _tuples(whole))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, backend_name, case="cfg1"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend_name, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for key, (sharded, whole) in res.items():
        assert len(whole) > 5, key
        assert sharded == whole, (key, world, case)   # Unsorted is the canonical (start, end, pattern) order on both sides


def test_plan_cuts_and_halo():
    from fac_b200 import sharding
    assert sharding.plan_shards(4, None, 3, n_bytes=10) == [(0, 3, 10), (3, 6, 10), (6, 10, 10)]
    assert sharding.plan_shards(4, None, 2, n_bytes=100) == [(0, 50, 57), (50, 100, 100)]
    t = bytes([0x61, 0xC3, 0xA9, 0x62])  # a é b: the cut must not land inside é
    assert sharding.plan_shards(0, t, 2) == [(0, 3, 4), (3, 4, 4)]
    # cuts never split an extended grapheme cluster (combining mark, CRLF, regional-indicator pair, ZWJ sequence)
    text = ("é" * 3 + "\r\n" * 3 + "🇩🇪" * 3 + "👩‍👩‍👧" * 2 + "한국") .encode("utf-8")
    from oracle_backend import OracleBackend
    ob = OracleBackend()
    starts = (C.c_uint64 * (len(text) + 1))()
    n = ob.lib.orc_grapheme_starts(text, len(text), starts, len(text) + 1)
    bounds = set(starts[:n]) | {len(text)}
    for world in range(2, 12):
        plan = sharding.plan_shards(2, text, world)
        assert plan[0][0] == 0 and plan[-1][1] == len(text)
        for k, (a, b, e) in enumerate(plan):
            assert a in bounds and b in bounds and e in bounds and a <= b <= e
            if k:
                assert a == plan[k - 1][1]
            # halo = max_match_graphemes + 3 clusters (or the end of the text)
            after = sorted(x for x in bounds if x >= b)
            assert e == (after[5] if len(after) > 5 else len(text))


def test_sharded_search_world2_gloo():
    _run(2, "oracle")


def test_sharded_search_world3_gloo():
    _run(3, "oracle")   # uneven cuts


@pytest.mark.parametrize("world", [2, 3, 4, 5, 7])
def test_sharded_search_unicode_gloo(world):
    _run(world, "oracle", "dense")


def test_sharded_unicode_every_cut_position():
    """Pattern occurrences straddle every possible cut of a non-ASCII haystack: shard-by-shard search through
    sharding.search_shard (oracle) equals the whole search for 2..7 shards and every rotation of the text."""
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, sharding
    from oracle_backend import OracleBackend
    be = OracleBackend()
    eng = FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).build(["éàüöñçéàüö"])
    base = "éàüöñçéàüö" * 4 + "zz" + "éàüöñçéàüö" * 3
    for pad in range(0, 12):
        text = np.frombuffer(("y" * pad + base).encode("utf-8"), dtype=np.uint8)
        whole, _ = be.search(eng._h, bytes(text), 0.7, 0, 0, False)
        assert any(m.similarity == 1.0 for m in whole)
        for world in range(2, 8):
            got = []
            for sh in sharding.plan_shards(eng.max_match_graphemes(), text, world):
                arr, _ = sharding.search_shard(eng, be, text, sh, 0.7, 0, whole_is_ascii=False)
                got += _tuples(arr)
            assert got == _tuples(whole), (pad, world)


def test_sharded_auto_beam_is_refused():
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, SearchError, sharding
    from oracle_backend import OracleBackend

    class FakeDist:
        @staticmethod
        def get_rank():
            return 0

        @staticmethod
        def get_world_size():
            return 2
    be = OracleBackend()
    eng = FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).auto_beam(100, 10).build(["hello"])
    with pytest.raises(SearchError):
        sharding.search_sharded(eng, be, b"hello world", 0.8, 0, 0, FakeDist)


@pytest.mark.gpu
def test_shard_replay_single_device(gpu, oracle):
    """The rank loop of search_sharded replayed on ONE device: every shard through fac_search_ex (device-resident
    result), lists concatenated in device memory, fac_matches_apply_device globally == the whole search == oracle."""
    import torch
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, _abi, sharding, workload
    cases = []
    cfg = workload.cfg2(1 << 18, 2000)
    cases.append((workload.build_engine(cfg, gpu), workload.build_engine(cfg, oracle), cfg["text"], 0.8))
    pats, text = _unicode_case("dense")
    mk = lambda be: FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True).build(pats)
    cases.append((mk(gpu), mk(oracle), np.frombuffer(text, dtype=np.uint8), 0.7))
    for eng, oeng, text, thr in cases:
        ascii_ = sharding.is_ascii(text)
        for order, overlap in ((0, 0), (1, 1), (2, 2)):
            whole, _ = gpu.search(eng._h, bytes(text), thr, order, overlap, False)
            ref, _ = oracle.search(oeng._h, bytes(text), thr, order, overlap, False)
            assert _tuples(whole) == _tuples(ref) and len(ref) > 5
            for world in (2, 5, 8):
                parts = []
                for sh in sharding.plan_shards(eng.max_match_graphemes(), text, world):
                    dm, _ = sharding.search_shard(eng, gpu, text, sh, thr, order, ascii_, result_on_device=True)
                    parts.append(dm.as_tensor("cuda"))
                buf = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device="cuda")
                flags = _abi.FAC_APPLY_PRESORTED if order == 0 else 0
                final, _ = gpu.apply_device(eng._h, buf.data_ptr(), buf.numel() // 32, order, overlap, flags)
                assert _tuples(final) == _tuples(whole), (world, order, overlap)


@pytest.mark.gpu
def test_shard_of_auto_beam_engine_is_unsupported(gpu):
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, SearchError
    eng = FuzzyAhoCorasickBuilder.new(gpu).fuzzy(FuzzyLimits.new().edits(1)).auto_beam(100, 10).build(["hello"])
    data = np.frombuffer(b"hello world hello", dtype=np.uint8)
    with pytest.raises(SearchError):
        gpu.search_ex(eng._h, data.ctypes.data, len(data), 0, 8, 0, 0.8, 0, 0, 0)


@pytest.mark.gpu
def test_sharded_search_world2_nccl():
    # two ranks over NCCL need two devices (NCCL refuses two ranks on one GPU); on 1-GPU boxes the device path is
    # covered by test_shard_replay_single_device and the NCCL path by bench.py's step-0 check at N > 1
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "gpu")
    _run(2, "gpu", "dense")


@pytest.mark.gpu
def test_stream_windows_dealt_over_devices(gpu, oracle):
    """fac_engine_create_multi + the stream pipeline (search_stream_parallel / replace_stream_parallel,
    src/stream.rs:378-429, 533-638): window batches dealt over every visible device (a single-device list on 1-GPU
    boxes still runs the threaded pipeline), results reassembled in stream order == the oracle's sequential stream."""
    import io
    import torch
    from fac_b200 import FuzzyAhoCorasickBuilder, FuzzyLimits, workload
    ndev = torch.cuda.device_count()
    devices = list(range(min(ndev, 4)))
    cfg = workload.cfg5(total=3 << 20, n_pairs=200, block=1 << 19, auto_beam=None)
    mk = lambda be, dev: (FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True)
                          .device(dev) if dev is not None else
                          FuzzyAhoCorasickBuilder.new(be).fuzzy(FuzzyLimits.new().edits(1)).case_insensitive(True)).build_replacer(cfg["pairs"])
    rg, ro = mk(gpu, devices), mk(oracle, None)
    assert gpu.lib.fac_engine_num_devices(rg.engine()._h) == len(devices)
    outs = []
    for rep in (rg, ro):
        sink = io.BytesIO()
        rep.replace_stream(workload.BlockReader(cfg["block"], cfg["total"], max_read=50000), sink, 0.8)
        outs.append(sink.getvalue())
    assert outs[0] == outs[1]
    assert len(outs[0]) != cfg["total"]     # replacements happened
    got, want = [], []
    rg.engine().search_stream(workload.BlockReader(cfg["block"], cfg["total"]), 0.8, lambda m: got.append((m.start, m.end, m.pattern_index)))
    ro.engine().search_stream(workload.BlockReader(cfg["block"], cfg["total"]), 0.8, lambda m: want.append((m.start, m.end, m.pattern_index)))
    assert got == want and len(got) > 100
