"""N>1 host logic (shard plan, halo, ownership, gather, global apply) with world_size 2 over gloo on the
CPU, the oracle standing in for the per-rank searcher; and the same flow on real GPUs (marked gpu)."""
import ctypes as C
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tuples(arr):
    return [(m.start, m.end, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value, m.insertions,
             m.deletions, m.substitutions, m.swaps, m.edits) for m in arr]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend_name, q):
    sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fac_b200 import sharding, workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend_name == "gpu":
        from fac_b200 import GpuBackend
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group("nccl", rank=rank, world_size=world)
        be, dev = GpuBackend(), "cuda"
    else:
        from oracle_backend import OracleBackend
        dist.init_process_group("gloo", rank=rank, world_size=world)
        be, dev = OracleBackend(), "cpu"
    cfg = workload.cfg1(1 << 15)
    eng = workload.build_engine(cfg, be, device=(rank % torch.cuda.device_count()) if backend_name == "gpu" else None)
    text = cfg["text"]
    res = {}
    for order, overlap in ((0, 0), (1, 1), (2, 2)):
        final = sharding.search_sharded(eng, be, text, 0.8, order, overlap, dist, dev)
        if rank == 0:
            whole, _ = be.search(eng._h, bytes(text), 0.8, order, overlap, False)
            res[(order, overlap)] = (_tuples(final), _tuples(whole))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, backend_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for key, (sharded, whole) in res.items():
        assert len(whole) > 20
        if key[0] == 0:
            assert sorted(sharded) == sorted(whole), key
        else:
            assert sharded == whole, key


def test_plan_and_halo():
    sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
    from fac_b200 import sharding
    assert sharding.plan_shards(10, 3) == [(0, 3), (3, 6), (6, 10)]
    t = bytes([0x61, 0xC3, 0xA9, 0x62])  # a é b: the cut must not land inside é
    assert sharding.plan_shards(4, 2, t) == [(0, 3), (3, 4)] or sharding.plan_shards(4, 2, t) == [(0, 2), (2, 4)]
    assert sharding.shard_slice(100, (0, 50), 7) == (0, 58)
    assert sharding.shard_slice(100, (50, 100), 7) == (50, 100)


def test_sharded_search_world2_gloo():
    _run(2, "oracle")


def test_sharded_search_world3_gloo():
    _run(3, "oracle")   # uneven cuts


@pytest.mark.gpu
def test_sharded_search_world2_nccl_single_device():
    # two ranks over NCCL; with one visible GPU both ranks share it (NCCL refuses that), so this needs >= 2 GPUs
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "gpu")
