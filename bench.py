#!/usr/bin/env python3
"""bench.py -- haystack GB/s of the fuzzy search path (BASELINE.json metric) on N B200s of one node.

Headline workload (`--config cfg2`, the configuration BASELINE.json's metric is quoted on): ONE 1 GiB synthetic
haystack, 10k ASCII patterns, edits(2), default penalties, threshold 0.8, Order::Unsorted / Overlap::Keep
(fac_b200/workload.py, seed 0xFAC00002).  A step is one `engine.search(hay, &SearchOptions)` over that haystack:

  N = 1   one call over the whole haystack;
  N > 1   STRONG scaling: the haystack is cut by fac_plan_shards into N shards (+ right halo); rank g searches its
          shard with fac_search_ex (match list ranked locally, kept in device memory), the lists are gathered to
          rank 0 over NCCL (all_gather of the counts + point-to-point sends of exactly count * 32 bytes into rank 0's
          device buffer) and rank 0 finishes with the global fac_matches_apply_device -- all inside the timed region.

  value : haystack bytes / step time, haystack (shards) resident in HBM when the timed region starts and the final
          list resident in rank 0's HBM when it ends
  e2e   : same step from HOST buffers: every rank copies its shard H2D from pinned memory, rank 0 copies the final
          list D2H (both inside the timed region); `e2e_pageable` repeats it from ordinary (pageable) host memory
  roofline / cpu_baseline : DESIGN.md "Measurement"

On step 0 (untimed) the result is verified: N > 1: the gathered sharded result equals rank 0's own whole-haystack
search byte for byte; every N: sampled regions (incl. the shard cuts) equal the CPU oracle (tests/ infrastructure).

  python bench.py --gpus N --steps K --warmup W [--config cfg1|cfg2|cfg3|cfg4|cfg5]   (N>1: launched by torchrun)
  python bench.py --impl reference ...       (CPU arm: the C++ restatement of the reference on all host cores)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "haystack GB/s (fuzzy, edits=2)"
# dram bytes (read + written) of one 2^25-window k_expand_succinct launch from the ncu --set full capture under profiles/
# (r2f_k_expand_succinct_raw_selected.txt: 1.735 GB read -- mostly rows of the 200 MB deep tables that miss the 126 MB L2 --
# + 381 MB written: the raw candidates)
TRAFFIC_NOTE = 2116422584


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# workloads: BASELINE.json configs.  Each returns the engine recipe and a generator for a byte range of the haystack.
# ---------------------------------------------------------------------------------------------------------------
CONFIG_BYTES = {"cfg1": 64 << 20, "cfg2": 1 << 30, "cfg3": 256 << 20, "cfg4": 4_000_000_000, "cfg5": 16 << 30}


def config_text(name, args, a, b, procs):
    """Bytes [a, b) of the config's haystack + the engine recipe (patterns, limits ...)."""
    from fac_b200 import workload
    total = args.bytes
    if name == "cfg2":
        return workload.cfg2(total, args.patterns, procs=procs, text_range=(a, b))
    if name == "cfg1":
        cfg = workload.cfg1(min(b, total))   # a prefix of the stream is the same text (workload.make_text)
        cfg["text"] = cfg["text"][a:b]
        return cfg
    raise SystemExit("unknown config " + name)


def describe(name, args):
    if name == "cfg2":
        return ("cfg2: %d ASCII patterns (len 5-16), FuzzyLimits edits(2), default penalties, threshold 0.8, Order::Unsorted / "
                "Overlap::Keep, ONE synthetic English-like haystack with planted fuzzy hits (seed 0xFAC00002)" % args.patterns)
    if name == "cfg1":
        return ("cfg1: 100 ASCII patterns, FuzzyLimits edits(1), case-insensitive, threshold 0.8, Order::Unsorted / Overlap::Keep, "
                "mixed-case synthetic English-like text (seed 0xFAC00001)")
    return name


def workload_config(args, world, extra=None):
    c = {"workload": describe(args.config, args), "config": args.config, "haystack_bytes_total": int(args.bytes),
         "haystack_bytes_per_gpu": int(args.bytes // max(world, 1)), "patterns": args.patterns, "threshold": args.threshold,
         "sharding": ("one fac_search call over the whole haystack" if world == 1 else
                      "strong scaling: fac_plan_shards cuts the haystack into %d shards + halo; fac_search_ex per rank, match lists "
                      "gathered to rank 0 over NCCL (exact-size send/recv of device-resident records), global "
                      "fac_matches_apply_device on rank 0, all inside the timed region" % world),
         "cache": "the haystack shard (>= 128 MiB) plus the candidate / reduction buffers written every step (> 1 GB per GPU) "
                  "exceed the 126 MB L2: inputs larger than L2, no explicit flush"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (C++ restatement, oracle/) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(cfg, cores, seconds_target):
    """Calibrate on a small slice and size a prefix sample for ~seconds_target of wall time on all cores.
    orc_search_parallel cuts the sample into one contiguous shard of start positions per thread (+ halo) -- the
    same decomposition the GPU ranks use."""
    seconds_target = float(os.environ.get("FAC_BENCH_CPU_SECONDS", seconds_target))  # tests shorten the sample
    from oracle_backend import OracleBackend
    from fac_b200 import workload
    ob = OracleBackend()
    eng = workload.build_engine(cfg, ob)
    text = cfg["text"]
    probe = min(len(text), 4096 * cores)
    t0 = time.time()
    ob.search_parallel(eng._h, text.ctypes.data, probe, cfg["threshold"], cores)
    rate = probe / max(time.time() - t0, 1e-6)
    sample = int(min(len(text), max(probe, rate * seconds_target)))
    return ob, eng, sample


def cpu_baseline_block(cfg, cores, seconds_target=15.0):
    ob, oeng, sample = cpu_arm(cfg, cores, seconds_target)
    text = cfg["text"]
    t2 = time.time()
    ob.search_parallel(oeng._h, text.ctypes.data, sample, cfg["threshold"], cores)
    dt = time.time() - t2
    # mode (i) of SURVEY 8d: single-thread whole-input search, on a smaller prefix (~1/4 of the time budget)
    one = max(4096, int(sample / cores / 4))
    t3 = time.time()
    ob.search_parallel(oeng._h, text.ctypes.data, min(one, len(text)), cfg["threshold"], 1)
    dt1 = time.time() - t3
    return {"value": sample / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "single_thread_value": min(one, len(text)) / dt1 / 1e9,
            "sample": "%d-byte prefix of the haystack on %d threads (one contiguous shard of start positions per thread); single-thread "
                      "figure on a %d-byte prefix; C++ restatement of the reference (the Rust crate cannot be built in this image)"
                      % (sample, cores, min(one, len(text)))}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cfg = config_text(args.config, args, 0, min(args.bytes, 8 << 20), 1)
    ob, eng, sample = cpu_arm(cfg, cores, seconds_target=15.0)
    text = cfg["text"]
    for _ in range(args.warmup):
        ob.search_parallel(eng._h, text.ctypes.data, min(sample, 4096 * cores), cfg["threshold"], cores)
    t0 = time.time()
    for _ in range(args.steps):
        ob.search_parallel(eng._h, text.ctypes.data, sample, cfg["threshold"], cores)
    dt = (time.time() - t0) / args.steps
    val = sample / dt / 1e9
    world = int(os.environ.get("WORLD_SIZE", 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, {"sample_bytes_per_step": sample}),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "each step searches a %d-byte prefix of the same haystack on %d threads (one contiguous shard of "
                                       "start positions per thread); C++ restatement of the reference" % (sample, cores)},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# verification helpers (step 0, untimed)
# ---------------------------------------------------------------------------------------------------------------
def oracle_sample_check(cfg_small, text, final_np, regions, region_len, halo):
    """Compare the final list with the CPU oracle on `regions` (start offsets) of `text` (numpy uint8, whole haystack or
    None-safe slices).  final_np: structured numpy view of the final fac_match list, ascending (start, end, pattern)."""
    import numpy as np
    from oracle_backend import OracleBackend
    from fac_b200 import workload
    ob = OracleBackend()
    oeng = workload.build_engine(cfg_small, ob)
    starts = final_np["start"]
    checked = 0
    for p in regions:
        sl = bytes(text[p:p + region_len + halo])
        arr, _ = ob.search(oeng._h, sl, cfg_small["threshold"], 0, 0, False)
        want = sorted((m.start + p, m.end + p, m.pattern_index, C.c_uint32.from_buffer(C.c_float(m.similarity)).value,
                       m.insertions, m.deletions, m.substitutions, m.swaps) for m in arr if m.start < region_len)
        lo, hi = np.searchsorted(starts, p, "left"), np.searchsorted(starts, p + region_len, "left")
        got = sorted((int(r["start"]), int(r["end"]), int(r["pat"]), int(r["simbits"]), int(r["ins"]), int(r["del"]), int(r["sub"]),
                      int(r["swap"])) for r in final_np[lo:hi])
        if got != want:
            raise SystemExit("PARITY FAILURE vs oracle in region [%d, %d): %d GPU records vs %d oracle records" % (p, p + region_len, len(got), len(want)))
        checked += len(want)
    return checked


MATCH_DTYPE = [("start", "<u8"), ("end", "<u8"), ("pat", "<u4"), ("simbits", "<u4"), ("ins", "u1"), ("del", "u1"), ("sub", "u1"),
               ("swap", "u1"), ("edits", "u1"), ("pad", "u1", (3,))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--bytes", type=int, default=0, help="total haystack bytes (default: the size BASELINE.md names for the config)")
    ap.add_argument("--patterns", type=int, default=int(os.environ.get("FAC_BENCH_PATTERNS", 10000)))
    ap.add_argument("--threshold", type=float, default=0.8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    if not args.bytes:
        args.bytes = int(os.environ.get("FAC_BENCH_BYTES", CONFIG_BYTES[args.config]))
    if args.config == "cfg1":
        args.patterns = 100
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.config in ("cfg3", "cfg4", "cfg5"):
        import bench_configs
        return bench_configs.run(args, sys.modules[__name__])

    import numpy as np
    import torch
    import torch.distributed as dist
    from fac_b200 import GpuBackend, _abi, sharding, workload

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    thr = args.threshold
    ORDER, OVERLAP = 0, 0
    gpu = GpuBackend()
    total = args.bytes
    # generator processes: rank 0 also produces the WHOLE haystack for the step-0 check, so it gets the cores the other ranks
    # (one shard each, 2 processes) leave free
    ncpu = os.cpu_count() or 1
    procs = max(1, min(16, ncpu // max(world, 1)))
    if world > 1:
        procs = max(2, min(16, ncpu - 2 * (world - 1))) if rank == 0 and not args.no_verify else max(1, min(2, ncpu // world))
    verify = not args.no_verify
    # engine first (the shard plan needs max_match_graphemes), then this rank's slice of the haystack
    recipe = config_text(args.config, args, 0, 0, 1)
    eng = workload.build_engine(recipe, gpu, device=local)
    plan = sharding.plan_shards(eng.max_match_graphemes(), None, world, n_bytes=total)
    own_a, own_b, read_end = plan[rank]
    whole_text = None
    if rank == 0 and (world == 1 or verify):
        whole_text = config_text(args.config, args, 0, total, procs)["text"]   # rank 0 keeps the whole haystack for the step-0 check
        text = whole_text[own_a:read_end]
    else:
        text = config_text(args.config, args, own_a, read_end, procs)["text"]
    n = len(text)
    own_len = own_b - own_a
    host = torch.from_numpy(np.ascontiguousarray(text)).pin_memory()
    pageable = np.ascontiguousarray(text).copy()
    dev = host.cuda(non_blocking=False)
    gather_buf = [None]
    F_DEV, F_RES = _abi.FAC_HAYSTACK_ON_DEVICE, _abi.FAC_RESULT_ON_DEVICE

    def step(src_ptr, on_device, final_on_device):
        """One engine.search over the whole haystack.  Returns (final list or None, per-step stats)."""
        t0 = time.perf_counter()
        if world == 1:
            flags = (F_DEV if on_device else 0) | (F_RES if final_on_device else 0)
            final, st = gpu.search_ex(eng._h, src_ptr, n, 0, n, 0, thr, ORDER, OVERLAP, flags)
            t1 = time.perf_counter()
            st.update(search_ms=(t1 - t0) * 1e3, gather_ms=0.0, apply_ms=0.0, local_matches=len(final))
            return final, st
        dm, st = gpu.search_ex(eng._h, src_ptr, n, 0, own_len, own_a, thr, ORDER, 0, (F_DEV if on_device else 0) | F_RES)
        t1 = time.perf_counter()
        buf, counts = sharding.gather_records(dm.as_tensor("cuda"), dist, "cuda", out=gather_buf[0])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        final = None
        if rank == 0:
            gather_buf[0] = buf
            flags = _abi.FAC_APPLY_PRESORTED | (F_RES if final_on_device else 0)
            final, st2 = gpu.apply_device(eng._h, buf.data_ptr(), sum(counts), ORDER, OVERLAP, flags)
            st["kernel_launches"] += st2["kernel_launches"]
        t3 = time.perf_counter()
        st.update(search_ms=(t1 - t0) * 1e3, gather_ms=(t2 - t1) * 1e3, apply_ms=(t3 - t2) * 1e3, local_matches=len(dm))
        return final, st

    # ---- step 0 (untimed): verification ----
    verified = {}
    final, st0 = step(host.data_ptr(), False, False)       # host flavour: rank 0 gets the final list on the host
    if verify and rank == 0:
        final_np = np.frombuffer(final, dtype=MATCH_DTYPE, count=len(final)) if len(final) else np.zeros(0, dtype=MATCH_DTYPE)
        if world > 1:
            whole_dev = torch.from_numpy(whole_text).cuda()
            whole, _ = gpu.search_ex(eng._h, whole_dev.data_ptr(), total, 0, total, 0, thr, ORDER, OVERLAP, F_DEV)
            same = len(whole) == len(final) and (len(final) == 0 or
                                                 np.array_equal(np.frombuffer(whole, dtype=np.uint8, count=len(whole) * 32),
                                                                np.frombuffer(final, dtype=np.uint8, count=len(final) * 32)))
            if not same:
                raise SystemExit("PARITY FAILURE: sharded result (%d matches) != whole-haystack search on rank 0 (%d matches)" % (len(final), len(whole)))
            verified["sharded_equals_whole"] = True
            del whole, whole_dev
        rng = np.random.default_rng(12345)
        region_len, halo = 192, eng.max_match_graphemes() + 3
        regions = [int(x) for x in rng.integers(0, max(1, total - region_len - halo), 40)]
        regions += [max(0, c[0] - region_len // 2) for c in plan[1:]] + [0, max(0, total - region_len)]
        verified["oracle_sample_records"] = oracle_sample_check(recipe, whole_text, final_np, regions, region_len, halo)
        verified["oracle_sample_regions"] = len(regions)
        del final_np
    n_final = len(final) if final is not None else 0
    del final
    barrier()

    # ---- resident: W warm-up + K timed steps ----
    for _ in range(args.warmup - 1):
        step(dev.data_ptr(), True, True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    per_step = []
    for _ in range(args.steps):
        f, st = step(dev.data_ptr(), True, True)
        per_step.append(st)
        del f
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    ev_ms = ev0.elapsed_time(ev1)

    # ---- end to end: pinned host shards in, final list on rank 0's host out ----
    step(host.data_ptr(), False, False)
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    e2e_steps = []
    for _ in range(args.steps):
        f, st = step(host.data_ptr(), False, False)
        e2e_steps.append(st)
        if f is not None:
            d2h = len(f) * 32
        del f
    barrier()
    wall_e2e = time.perf_counter() - t1
    # pageable input: a few steps (the copy runs at a few GB/s)
    pg_steps = max(1, min(args.steps, 3))
    barrier()
    t2 = time.perf_counter()
    for _ in range(pg_steps):
        f, _ = step(pageable.ctypes.data, False, False)
        del f
    barrier()
    wall_pg = time.perf_counter() - t2
    sampler.stop_flag = True
    sampler.join(timeout=2)

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K = args.steps
    mean = lambda key, rows=per_step: sum(r[key] for r in rows) / len(rows)
    mine = {"rank": rank, "own_bytes": own_len, "device_ms": mean("device_ms"), "expand_ms": mean("expand_ms"), "search_ms": mean("search_ms"),
            "gather_ms": mean("gather_ms"), "apply_ms": mean("apply_ms"), "matches": per_step[-1]["local_matches"],
            "e2e_search_ms": mean("search_ms", e2e_steps), "e2e_gather_ms": mean("gather_ms", e2e_steps), "e2e_apply_d2h_ms": mean("apply_ms", e2e_steps)}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    ms_step = maxr(max(ev_ms, wall * 1e3) / K)        # CUDA events on the current stream bracket host-synchronous steps: == wall
    ms_e2e = maxr(wall_e2e * 1e3 / K)
    ms_pg = maxr(wall_pg * 1e3 / pg_steps)
    states = sumr(float(mean("states_pushed")))
    launches = int(sumr(float(sum(r["kernel_launches"] for r in per_step))))
    clocks = sampler.summary()

    if rank == 0:
        peak, which = peaks()
        exp_ms_max = max(r["expand_ms"] for r in per_rank)
        alg_rank0 = own_len + 32.0 * mine["matches"]
        alg_total = total + 32.0 * sum(r["matches"] for r in per_rank)
        achieved = alg_rank0 / mine["expand_ms"] / 1e6 if mine["expand_ms"] > 0 else 0.0
        achieved_all = alg_total / exp_ms_max / 1e6 if exp_ms_max > 0 else 0.0
        # the reference counts pushed states (queue.len(), src/search.rs:1099); the fast kernel visits fewer (it never
        # pushes children that cannot emit).  The ratio is measured on a 1 MiB sample with the order-faithful kernel.
        ref_states_per_byte = None
        if not args.no_verify:
            try:
                os.environ["FAC_FAITHFUL"] = "1"
                feng = workload.build_engine(recipe, gpu, device=local)
                os.environ.pop("FAC_FAITHFUL")
                sb = min(n, 1 << 20)
                _, fst = gpu.search_ex(feng._h, dev.data_ptr(), sb, 0, sb, 0, thr, 0, 0, F_DEV | F_RES)
                ref_states_per_byte = fst["states_pushed"] / sb
            except Exception as e:  # a statistic, never fatal
                sys.stderr.write("queue.len() sample failed: %r\n" % (e,))
                os.environ.pop("FAC_FAITHFUL", None)
        line = {"metric": METRIC, "value": total / ms_step / 1e6, "unit": "GB/s", "n_gpus": world, "steps": K,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world),
                "states_per_s": {"visited_by_kernel": states / (exp_ms_max / 1e3) if exp_ms_max > 0 else None,
                                 "reference_queue_len_equivalent": (ref_states_per_byte * total / (exp_ms_max / 1e3)) if ref_states_per_byte else None,
                                 "reference_queue_len_per_byte": ref_states_per_byte,
                                 "note": "visited = states the fast kernel popped or walked; reference equivalent = the reference's pushed-state "
                                         "count (sum of queue.len(), measured by the order-faithful kernel on a 1 MiB sample) per second of "
                                         "k_expand time"},
                "matches_per_step": n_final,
                "e2e": {"value": total / ms_e2e / 1e6, "unit": "GB/s", "h2d_bytes_per_step": int(sum(r["own_bytes"] for r in per_rank) + sum(p[2] - p[1] for p in plan)),
                        "d2h_bytes_per_step": int(d2h), "steps": K, "input": "pinned host memory"},
                "e2e_pageable": {"value": total / ms_pg / 1e6, "unit": "GB/s", "steps": pg_steps, "input": "pageable host memory"},
                "gpu_launches": launches,
                "per_rank": per_rank,
                "verified": verified,
                "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": TRAFFIC_NOTE, "peak_source": which, "kernel": "k_expand_succinct",
                             "achieved_all_gpus": achieved_all, "frac_of_%dx" % world: achieved_all / (peak * world),
                             "note": "algorithmic bytes = owned haystack bytes + 32 B x raw matches, divided by the CUDA-event time of the "
                                     "k_expand_succinct launches of the step (rank 0; all-GPU figure: total bytes / slowest rank); the kernel is "
                                     "instruction-issue bound (hundreds of trie states per input byte), see DESIGN.md and profiles/"}}
        if not args.no_cpu_baseline and world == 1:
            cfg_cpu = dict(recipe)
            cfg_cpu["text"] = whole_text[: min(total, 8 << 20)]
            line["cpu_baseline"] = cpu_baseline_block(cfg_cpu, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
