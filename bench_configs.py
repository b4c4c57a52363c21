"""bench_configs.py -- the secondary BASELINE.json configurations (`bench.py --config cfg3|cfg4|cfg5`), one JSON line each
in the same contract as the headline line (value / e2e / roofline / cpu_baseline / clocks / gpu_launches).

  cfg3  Unicode workload: 1k patterns (Cyrillic / CJK / NFD Latin / German-Nordic), case-insensitive, mappings
        ae<->ae-ligature, ss<->sharp-s, ks<->x, edits(2), threshold 0.8; 256 MiB UTF-8 haystack (an 8 MiB generated block
        repeated on a word boundary -- the Python generator runs at ~3 MB/s)
  cfg4  sparse 4 000 000 000-byte haystack, 100 weighted patterns with per-pattern limits, bitap pre-filter on,
        threshold 0.85, sorted().non_overlapping()
  cfg5  streaming Read source (a repeating 1 MiB block, short reads at block ends), auto_beam(200000, 100), edits(2),
        case-insensitive, 1000-pair FuzzyReplacer::replace_stream, absolute u64 offsets.  Reader, writer and the
        replacement table are native (tests/abi_c/facio.c), so the line measures the library, not Python callbacks.
        The stream length per step is the largest that fits ~20 s at the measured rate (stated in the line); 16 GiB at
        that rate is given as `projected_16GiB_s`.

cfg3 / cfg4 run on one GPU (rank 0; other ranks exit): the pre-filter and the auto_beam budget are whole-haystack
properties.  cfg5 uses every GPU of the job from rank 0's process: `--gpus N` builds ONE engine replicated on
devices 0..N-1 (fac_engine_create_multi) and the stream pipeline deals its window batches over them (the other ranks of a
torchrun launch exit without touching their device).
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def _cut_utf8(b):
    """Trim a byte slice to whole UTF-8 scalars: continuation bytes at its start (the slice began inside a scalar) and an
    incomplete scalar at its end."""
    k = 0
    while k < len(b) and (b[k] & 0xC0) == 0x80:
        k += 1
    b = b[k:]
    k = len(b)
    while k and (b[k - 1] & 0xC0) == 0x80:
        k -= 1
    if k and b[k - 1] >= 0xC0:      # the last lead byte: keep it only if its scalar is complete
        need = 2 if b[k - 1] < 0xE0 else 3 if b[k - 1] < 0xF0 else 4
        if len(b) - (k - 1) < need:
            b = b[:k - 1]
    return b


def cfg3_text(total):
    import numpy as np
    from fac_b200 import workload
    block_bytes = min(total, 8 << 20)
    cfg = workload.cfg3(block_bytes)
    blk = bytes(cfg["text"])
    blk = blk[: blk.rfind(b" ") + 1] if b" " in blk else blk
    reps = max(1, total // len(blk))
    cfg["text"] = np.frombuffer(blk * reps, dtype=np.uint8).copy()
    return cfg


def _cfg4_block(args):
    from fac_b200 import workload
    seed, k, size, n_patterns, every = args
    vocab = workload.make_vocab(seed)
    pats = workload.random_words(seed, n_patterns, 8, 20)
    return workload.plant(workload.make_text(seed, size, vocab, mixed_case=False, stream=k), pats, seed, every=every, stream=k)


def cfg4_text(total, procs):
    """cfg4 haystack generated block-wise like cfg2 (independent PCG64 streams per 32 MiB block, process pool)."""
    import multiprocessing as mp
    import numpy as np
    from fac_b200 import workload
    cfg = workload.cfg4(0)
    B = 32 << 20
    blocks = [(0xFAC00004, k, min(B, total - k * B), 100, 1 << 20) for k in range((total + B - 1) // B)]
    out = np.empty(total, dtype=np.uint8)
    if procs > 1 and len(blocks) > 1:
        with mp.get_context("fork").Pool(min(procs, len(blocks))) as pool:
            pos = 0
            for part in pool.imap(_cfg4_block, blocks):
                out[pos:pos + len(part)] = part
                pos += len(part)
    else:
        pos = 0
        for b in blocks:
            part = _cfg4_block(b)
            out[pos:pos + len(part)] = part
            pos += len(part)
    cfg["text"] = out
    return cfg


def cpu_slices(ob, eng, text, thr, order, overlap, prefilter, cores, seconds, cut=None):
    """The oracle on `cores` host threads over consecutive slices of a prefix sample (ctypes releases the GIL).  Returns
    (bytes searched, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    probe = bytes(text[: 1 << 14]) if cut is None else cut(bytes(text[: 1 << 14]))
    t0 = time.time()
    ob.search(eng._h, probe, thr, order, overlap, prefilter)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    per = int(max(1 << 14, min(len(text) // cores, rate * seconds)))
    slices = []
    for k in range(cores):
        s = bytes(text[k * per:(k + 1) * per])
        slices.append(s if cut is None else cut(s))
    slices = [s for s in slices if s]
    t1 = time.time()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda s: ob.search(eng._h, s, thr, order, overlap, prefilter), slices))
    return sum(len(s) for s in slices), time.time() - t1


def load_facio():
    import test_abi_c
    test_abi_c.build()
    lib = C.CDLL(os.path.join(ROOT, "tests", "abi_c", "_build", "libfacio.so"))
    return lib


class BlockReader(C.Structure):
    _fields_ = [("block", C.c_void_p), ("block_len", C.c_size_t), ("total", C.c_uint64), ("pos", C.c_uint64), ("max_read", C.c_size_t),
                ("fail_at", C.c_int64)]


class Sink(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("fnv", C.c_uint64), ("keep", C.c_void_p), ("keep_cap", C.c_size_t), ("keep_len", C.c_size_t),
                ("fail_at", C.c_int64)]


def run(args, B):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from fac_b200 import GpuBackend, _abi, workload
    from oracle_backend import OracleBackend
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    gpu = GpuBackend()
    cores = os.cpu_count() or 1
    procs = max(1, min(16, cores))
    name = args.config
    K, W = args.steps, args.warmup
    peak, which = B.peaks()
    sampler = B.ClockSampler(local)
    line = {"metric": B.METRIC, "unit": "GB/s", "n_gpus": 1, "steps": K, "warmup": W, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}

    if name in ("cfg3", "cfg4"):
        if name == "cfg3":
            cfg = cfg3_text(args.bytes)
            order, overlap, pf, kernel = 0, 0, False, "k_expand_stack<false,true> (stack-machine kernel over merged records: mappings + grapheme-id stream)"
            desc = ("cfg3: 1000 Unicode patterns (Cyrillic / CJK / NFD Latin / German-Nordic), case-insensitive, mappings ae<->ae-ligature, "
                    "ss<->sharp-s, ks<->x, edits(2), threshold 0.8, Order::Unsorted / Overlap::Keep; UTF-8 haystack = an 8 MiB generated block repeated")
            cut = _cut_utf8
        else:
            cfg = cfg4_text(args.bytes, procs)
            order, overlap, pf, kernel = 1, 1, True, "k_bitap_scan + k_expand_succinct (limits mode) on the candidate slices"
            desc = ("cfg4: sparse haystack, 100 weighted patterns (len 8-20) with per-pattern limits (edits(1) / edits(2) / edits(2).swaps(0) / "
                    "substitutions(1).deletions(1)), bitap pre-filter on, threshold 0.85, sorted().non_overlapping(), one planted hit per ~1 MiB")
            cut = None
        thr = cfg["threshold"]
        eng = workload.build_engine(cfg, gpu, device=local)
        text = cfg["text"]
        n = len(text)
        host = torch.from_numpy(text).pin_memory()
        dev = host.cuda()

        def resident():
            arr, st = gpu.search_device(eng._h, dev.data_ptr(), n, thr, order, overlap, pf)
            return len(arr), st

        def e2e():
            arr, st = gpu.search_host_ptr(eng._h, host.data_ptr(), n, thr, order, overlap, pf)
            return len(arr), st

        for _ in range(W):
            resident()
        sampler.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = []
        for _ in range(K):
            nm, st = resident()
            rows.append(st)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        e2e()
        t1 = time.perf_counter()
        for _ in range(K):
            nm2, _ = e2e()
        wall_e2e = time.perf_counter() - t1
        sampler.stop_flag = True
        sampler.join(timeout=2)
        ms = wall * 1e3 / K
        dev_ms = sum(r["device_ms"] for r in rows) / K
        exp_ms = sum(r["expand_ms"] for r in rows) / K
        alg = n + 32.0 * nm
        # dominant kernel: cfg3 = the frontier expansion; cfg4 = the bitap scan (device time minus the expansion of the slices)
        dom_ms = exp_ms if name == "cfg3" else max(dev_ms - exp_ms, 1e-9)
        alg_dom = alg if name == "cfg3" else float(n)
        line.update({"value": n / ms / 1e6, "ms_per_step": ms, "device_ms_per_step": dev_ms, "expand_ms_per_step": exp_ms,
                     "config": {"workload": desc, "config": name, "haystack_bytes_total": int(n), "patterns": len(cfg["patterns"]), "threshold": thr,
                                "cache": "haystack larger than the 126 MB L2, no explicit flush"},
                     "matches_per_step": nm,
                     "states_per_s": {"visited_by_kernel": sum(r["states_pushed"] for r in rows) / K / (exp_ms / 1e3) if exp_ms > 0 else None},
                     "e2e": {"value": n / (wall_e2e * 1e3 / K) / 1e6, "unit": "GB/s", "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": int(nm2 * 32),
                             "steps": K, "input": "pinned host memory"},
                     "gpu_launches": int(sum(r["kernel_launches"] for r in rows)),
                     "clocks": sampler.summary(),
                     "roofline": {"bound": "hbm", "achieved": alg_dom / dom_ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg_dom / dom_ms / 1e6 / peak,
                                  "traffic": None, "peak_source": which, "kernel": kernel,
                                  "note": "algorithmic bytes (haystack%s) / CUDA-event time of the dominant kernel launches"
                                          % (" + 32 B x raw matches" if name == "cfg3" else " read once by the bitap scan")}})
        if not args.no_cpu_baseline:
            ob = OracleBackend()
            oeng = workload.build_engine(cfg, ob)
            nb, dt = cpu_slices(ob, oeng, text, thr, order, overlap, pf, cores, float(os.environ.get("FAC_BENCH_CPU_SECONDS", 15.0)), cut)
            line["cpu_baseline"] = {"value": nb / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": "%d bytes of the haystack as %d consecutive slices, one host thread each (C++ restatement of the reference)" % (nb, cores)}
        print(json.dumps(line), flush=True)
        return 0

    # ---- cfg5: streaming find-and-replace ----
    cfg = workload.cfg5(total=args.bytes)
    thr = cfg["threshold"]
    world = max(1, int(os.environ.get("WORLD_SIZE", args.gpus)))
    ndev = min(world, torch.cuda.device_count())
    rep = workload.build_engine(cfg, gpu, device=list(range(ndev)) if ndev > 1 else None)          # FuzzyReplacer (one replica per device)
    eng = rep.engine()
    io = load_facio()
    lib = gpu.lib
    block = np.frombuffer(cfg["block"], dtype=np.uint8).copy()
    repl = [r.encode() for _, r in cfg["pairs"]]
    arr_p = (C.c_void_p * len(repl))(*[C.cast(C.c_char_p(r), C.c_void_p) for r in repl])
    arr_l = (C.c_size_t * len(repl))(*[len(r) for r in repl])

    def stream(total):
        rd = BlockReader(block.ctypes.data, len(block), total, 0, 65536, -1)
        sk = Sink()
        io.facio_sink_init(C.byref(sk), None, 0)
        ss = _abi.fac_stream_stats()
        t0 = time.perf_counter()
        st = lib.fac_replace_stream_table(eng._h, C.cast(io.facio_block_read, C.c_void_p), C.byref(rd), C.cast(io.facio_sink_write, C.c_void_p),
                                          C.byref(sk), thr, arr_p, arr_l, len(repl), C.byref(ss))
        dt = time.perf_counter() - t0
        if st != 0:
            raise SystemExit("fac_replace_stream_table failed: %s" % lib.fac_last_error_string().decode())
        return dt, ss, sk

    dt, ss, _ = stream(4 << 20)                       # probe (also warm-up)
    rate = (4 << 20) / dt
    budget = float(os.environ.get("FAC_BENCH_STREAM_SECONDS", 20.0))
    total = int(min(args.bytes, max(8 << 20, rate * budget)))
    total -= total % (1 << 20)
    for _ in range(max(0, W - 1)):
        stream(min(total, 8 << 20))
    sampler.start()
    t0 = time.perf_counter()
    rows = []
    for _ in range(K):
        rows.append(stream(total))
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = wall * 1e3 / K
    ss = rows[-1][1]
    sk = rows[-1][2]
    exp_ms = sum(r[1].expand_ms for r in rows) / K
    alg = total + 32.0 * ss.matches
    line["n_gpus"] = int(ss.devices)
    exp_ms = exp_ms / max(1, int(ss.devices) * 2)   # expand_ms sums over the batches in flight (two workers per device)
    line.update({"value": total / ms / 1e6, "ms_per_step": ms, "device_ms_per_step": sum(r[1].device_ms for r in rows) / K, "expand_ms_per_step": exp_ms,
                 "config": {"workload": "cfg5: Read source repeating a 1 MiB block (64 KiB reads, short read at every block end), auto_beam(200000, 100), edits(2), "
                                        "case-insensitive, 1000-pair FuzzyReplacer::replace_stream, threshold 0.8, absolute u64 offsets; native reader / writer / "
                                        "replacement table (tests/abi_c/facio.c)",
                            "config": "cfg5", "stream_bytes_per_step": int(total), "stream_bytes_named": int(args.bytes), "pairs": len(repl), "threshold": thr,
                            "projected_16GiB_s": (16 << 30) / (total / (ms / 1e3)),
                            "cache": "every 256 KiB window is a fresh host buffer copied to the device; the per-window queues exceed L2"},
                 "matches_per_step": int(ss.matches), "windows_per_step": int(ss.windows), "bytes_written_per_step": int(sk.bytes),
                 "states_per_s": {"reference_queue_len": ss.states / (exp_ms / 1e3) if exp_ms > 0 else None},
                 "e2e": {"value": total / ms / 1e6, "unit": "GB/s", "h2d_bytes_per_step": int(total), "d2h_bytes_per_step": int(ss.matches * 32), "steps": K,
                         "input": "host Read source through fac_replace_stream_table (the stream API has no device-resident flavour: value == e2e)"},
                 "gpu_launches": int(sum(r[1].kernel_launches for r in rows)),
                 "clocks": sampler.summary(),
                 "roofline": {"bound": "hbm", "achieved": alg / exp_ms / 1e6 if exp_ms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                              "frac": (alg / exp_ms / 1e6 / peak) if exp_ms > 0 else 0.0, "traffic": None, "peak_source": which,
                              "kernel": "k_beam_warp (order-faithful pop loop with the beam cut, one warp per start window, queue + visited map in shared memory)",
                              "note": "algorithmic bytes (stream bytes + 32 B x owned matches) / CUDA-event time of the beamed expansion launches; the launches of "
                                      "the batches in flight overlap (two pipeline workers per device), so their summed time is divided by 2 x devices"}})
    if not args.no_cpu_baseline:
        ob = OracleBackend()
        orep = workload.build_engine(cfg, ob)
        # the reference's own parallel driver is replace_stream_parallel (src/stream.rs:533-638): windows over threads
        import io as pyio
        from concurrent.futures import ThreadPoolExecutor
        per = 1 << 17
        secs = float(os.environ.get("FAC_BENCH_CPU_SECONDS", 15.0))
        t1 = time.time()
        orep.replace_stream(workload.BlockReader(cfg["block"], 1 << 15), pyio.BytesIO(), thr)
        r1 = (1 << 15) / max(time.time() - t1, 1e-6)
        per = int(max(1 << 15, min(1 << 20, r1 * secs)))
        t2 = time.time()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda k: orep.replace_stream(workload.BlockReader(cfg["block"], per), pyio.BytesIO(), thr), range(cores)))
        dt = time.time() - t2
        line["cpu_baseline"] = {"value": per * cores / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                "sample": "%d host threads, each a replace_stream over a %d-byte prefix of the stream (C++ restatement of the reference)" % (cores, per)}
    print(json.dumps(line), flush=True)
    return 0
