#!/usr/bin/env python3
"""bench.py -- haystack GB/s of the fuzzy search path (BASELINE.json metric) on N B200s of one node.

A step is ONE pass of the hot path (`engine.search(hay, &SearchOptions)` through the C ABI) over one
synthetic haystack shard per GPU: cfg2 of BASELINE.json -- 10k ASCII patterns, edits(2), default
penalties, threshold 0.8, English-like text with planted fuzzy hits (fac_b200/workload.py).  The
haystack shards naturally: every rank searches its own shard (+halo) with no data-path
collective; the only exchange is the final gather of the match counts / lists.

  value : whole-job throughput, shards already resident in HBM when the timed region starts
  e2e   : same metric through fac_search() on HOST (pinned) buffers -- H2D of the shard and D2H of
          the match list inside the timed region
  roofline / cpu_baseline : see DESIGN.md ("Measurement")

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU arm: the C++ restatement of the reference)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULT_BYTES = int(os.environ.get("FAC_BENCH_BYTES", 128 << 20))  # haystack bytes per GPU per step
DEFAULT_PATTERNS = int(os.environ.get("FAC_BENCH_PATTERNS", 10000))
THRESHOLD = 0.8
# dram bytes of one k_expand_succinct launch from the ncu --set full capture under profiles/ (None until measured)
TRAFFIC_NOTE = 404099584  # profiles/r1_k_expand_succinct_raw_selected.txt: 65.5 MB read + 338.6 MB written (raw candidates, ~0.37 per start window) per 2^25-window launch
METRIC = "haystack GB/s (fuzzy, edits=2)"


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


def make_shard(rank, nbytes, npat):
    from fac_b200 import workload
    cfg = workload.cfg2(nbytes, npat, seed=0xFAC00002)
    if rank:  # same patterns, a different text stream per rank (weak scaling: per-GPU work is fixed)
        import numpy as np
        vocab = workload.make_vocab(0xFAC00002)
        text = workload.make_text(0xFAC00002 + 7919 * rank, nbytes, vocab)
        cfg["text"] = workload.plant(text, [p.encode() for p in cfg["patterns"]], 0xFAC00002 + rank)
    return cfg


def cpu_arm(args, cfg, cores, seconds_target=20.0):
    seconds_target = float(os.environ.get("FAC_BENCH_CPU_SECONDS", seconds_target))  # tests shorten the sample
    """The reference's algorithm on the host cores: C++ restatement (oracle), all cores, bounded sample."""
    from oracle_backend import OracleBackend
    from fac_b200 import workload
    ob = OracleBackend()
    eng = workload.build_engine(cfg, ob)
    text = cfg["text"]
    # calibrate on a tiny slice, then size the sample for ~seconds_target of wall time on all cores
    probe = min(len(text), 4096 * cores)
    t0 = time.time()
    ob.search_parallel(eng._h, text.ctypes.data, probe, THRESHOLD, cores)
    rate = probe / max(time.time() - t0, 1e-6)
    sample = int(min(len(text), max(probe, rate * seconds_target)))
    return ob, eng, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cfg = make_shard(0, min(args.bytes, 8 << 20), args.patterns)
    ob, eng, sample = cpu_arm(args, cfg, cores, seconds_target=15.0)
    text = cfg["text"]
    for _ in range(args.warmup):
        ob.search_parallel(eng._h, text.ctypes.data, min(sample, 4096 * cores), THRESHOLD, cores)
    t0 = time.time()
    for _ in range(args.steps):
        ob.search_parallel(eng._h, text.ctypes.data, sample, THRESHOLD, cores)
    dt = (time.time() - t0) / args.steps
    val = sample / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, sample),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "%d-byte prefix of the rank-0 shard per step, %d threads" % (sample, cores)},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, nbytes):
    return {"workload": "cfg2: %d ASCII patterns (len 5-16), FuzzyLimits edits(2), default penalties, threshold 0.8, "
                        "Order::Unsorted / Overlap::Keep, synthetic English-like haystack with planted fuzzy hits" % args.patterns,
            "haystack_bytes_per_gpu": int(nbytes), "patterns": args.patterns, "threshold": THRESHOLD,
            "sharding": "one shard per GPU, no data-path collective; match lists gathered to rank 0",
            "cache": "haystack shard (128 MiB default) plus the candidate / reduction buffers written every step (> 1 GB) exceed the "
                     "126 MB L2, so no step finds its input cached"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--bytes", type=int, default=DEFAULT_BYTES, help="haystack bytes per GPU per step")
    ap.add_argument("--patterns", type=int, default=DEFAULT_PATTERNS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fac_b200 import GpuBackend, workload

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

    gpu = GpuBackend()
    cfg = make_shard(rank, args.bytes, args.patterns)
    eng = workload.build_engine(cfg, gpu, device=local)
    text = cfg["text"]
    n = len(text)
    host = torch.from_numpy(text).pin_memory()
    dev = host.cuda(non_blocking=False)

    def step_resident():
        arr, st = gpu.search_device(eng._h, dev.data_ptr(), n, THRESHOLD, 0, 0, False)
        return len(arr), st

    def step_e2e():
        arr, st = gpu.search_host_ptr(eng._h, host.data_ptr(), n, THRESHOLD, 0, 0, False)
        return len(arr), st

    def gather_counts(cnt):
        # the only exchange on the path: the final gather of the per-shard match lists (here their sizes;
        # bench.py does not need the records on rank 0, tests/test_multi_gpu.py gathers the records)
        if world == 1:
            return cnt
        t = torch.tensor([cnt], device="cuda", dtype=torch.int64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return int(sum(int(x.item()) for x in out))

    for _ in range(args.warmup):
        c, _ = step_resident()
        gather_counts(c)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = expand_ms = 0.0
    launches = states = 0
    matches = 0
    for _ in range(args.steps):
        c, st = step_resident()
        matches = gather_counts(c)
        dev_ms += st["device_ms"]
        expand_ms += st["expand_ms"]
        launches += st["kernel_launches"]
        states += st["states_pushed"]
    barrier()
    wall = time.perf_counter() - t0
    my_matches = c
    # end to end: host buffers, H2D + D2H inside the timed region
    step_e2e()
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    d2h = 0
    for _ in range(e2e_steps):
        c2, _ = step_e2e()
        d2h = c2 * 32
        gather_counts(c2)
    barrier()
    wall_e2e = time.perf_counter() - t1
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

    ms_step_dev = maxr(dev_ms / args.steps)          # CUDA events on the library's stream, max over ranks
    ms_step_wall = maxr(wall * 1e3 / args.steps)
    ms_step = max(ms_step_dev, 1e-9)
    ms_e2e = maxr(wall_e2e * 1e3 / e2e_steps)
    total_bytes = sumr(float(n))
    ms_expand = maxr(expand_ms / args.steps)
    total_states = sumr(float(states) / args.steps)
    total_launches = int(sumr(float(launches)))
    clocks = sampler.summary()

    if rank == 0:
        peak, which = peaks()
        # dominant kernel: k_expand.  algorithmic bytes per launch = haystack bytes + 32 B per raw match
        n_exp_launches = max(1, round((launches / args.steps - 8) / 1))  # informative only
        alg_bytes = n + 32.0 * my_matches
        achieved = alg_bytes / (expand_ms / args.steps) / 1e6 if expand_ms > 0 else 0.0
        line = {"metric": METRIC, "value": total_bytes / ms_step / 1e6, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": ms_step_wall, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, n),
                "states_per_s": total_states / (ms_expand / 1e3) if ms_expand > 0 else None,
                "matches_per_step": matches,
                "e2e": {"value": total_bytes / ms_e2e / 1e6, "unit": "GB/s", "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": total_launches,
                "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": TRAFFIC_NOTE, "peak_source": which, "kernel": "k_expand_succinct",
                             "note": "algorithmic bytes = haystack bytes + 32 B x raw matches per step, divided by the CUDA-event time of the "
                                     "k_expand_succinct launches of the step; the kernel is instruction-issue bound (~2400 trie states per "
                                     "input byte, ncu: IPC 3.1 of 4, DRAM < 0.1 %), see DESIGN.md and profiles/"}}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ob, oeng, sample = cpu_arm(args, cfg, cores, seconds_target=15.0)
            t2 = time.time()
            ob.search_parallel(oeng._h, text.ctypes.data, sample, THRESHOLD, cores)
            dt = time.time() - t2
            line["cpu_baseline"] = {"value": sample / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": "%d-byte prefix of the rank-0 shard, %d threads (C++ restatement of the reference)" % (sample, cores)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
