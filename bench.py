#!/usr/bin/env python3
"""
bench.py — BASELINE.json metric: corpus GB/s and merges/s through the train() merge loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size-mib M]

A "step" is one merge iteration (get_stats -> arg-max with the reference tie-break -> merge) over
the whole resident token stream.  Workload at N=1: BASELINE.json configs[2], RegexTokenizer.train
(GPT-4 split pattern) on 1 GiB of synthetic UTF-8 (seed 1337), steps W..W+K of its merge loop.
With N>1 every rank holds its own 1 GiB shard (weak scaling, contiguous byte ranges of one
N GiB corpus) and the per-merge statistics delta is all-reduced over NCCL.

value   = corpus_bytes * K / t           (device-resident stream, CUDA-event time, max over ranks)
e2e     = the same through the C ABI from a pinned HOST text buffer: bpe_load_text_gpt4 (H2D of the text,
          GPT-4 split on the device, marked stream) + bpe_train(W+K merges) + D2H of the merges, wall
          clock around the calls.  Its merges must equal those of the device-resident run, which is
          loaded from the host `regex` split: a 1 GiB cross-check of the device splitter.
roofline= fused merge kernel: (4*N_in + 4*N_out bytes per launch) / CUDA-event time per launch,
          against MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference = the CPU oracle port of the reference loop (oracle/bpe_oracle.c,
          base.py:13-41 + regex.py:49-63 restated in C) on a bounded sample of the same corpus.
The host regex pre-split (third-party `regex`) is outside every timed region; the e2e leg does its
own split on the device inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GPT4 = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+"""


def kernel_source_sha():
    """sha256 over the sources of the dominant kernel (k_merge_seg): a committed ncu capture is only quoted
    while it describes the kernel that is being timed."""
    import hashlib
    h = hashlib.sha256()
    for f in ("k_merge_seg.cuh", "k_merge.cuh", "common.cuh", "k_seg.cuh"):
        h.update(open(os.path.join(ROOT, "minbpe_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """DRAM bytes of one profiled launch of the dominant kernel (ncu --set full, committed under profiles/);
    null when no capture of the CURRENT kernel source is on record (the capture carries the source hash)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:  # noqa: BLE001
        return None, None
    sha = kernel_source_sha()
    if d.get("kernel_source_sha") != sha:
        return None, {"stale": True, "capture_sha": d.get("kernel_source_sha"), "current_sha": sha,
                      "note": "profiles/ncu_traffic.json was captured from another version of the kernel; re-run tools/ncu_traffic.sh"}
    return d["dram_bytes"], d


from minbpe_b200.presplit import host_cores  # noqa: E402  (affinity mask capped by the cgroup CPU quota)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pin_host(arr):
    """cudaHostRegister the numpy buffer (the contract's "pinned host memory"); False if refused."""
    import torch
    try:
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        return int(rc) == 0
    except Exception:  # noqa: BLE001
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.proc = gpu, [], None
        self.t_begin = self.t_end = None

    def start(self):
        """Start sampling (nvidia-smi needs ~0.2 s to produce its first line: call this well before the timed region)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t_end is None:
            self.t_end = time.perf_counter()
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        lo = self.t_begin if self.t_begin is not None else 0.0
        rows = [r[1:] for r in self.rows if lo <= r[0] <= self.t_end + 0.03]
        window = "timed region"
        if not rows and self.rows:   # region shorter than the sampling period: the sample nearest to it
            mid = 0.5 * (lo + self.t_end)
            rows = [min(self.rows, key=lambda r: abs(r[0] - mid))[1:]]
            window = "nearest sample to the timed region"
        self.rows = rows
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def make_corpus(size_bytes, seed):
    from minbpe_b200.presplit import chunk_offsets
    from minbpe_b200.synth import generate
    t0 = time.time()
    raw = generate(seed, size_bytes)
    t1 = time.time()
    offs = chunk_offsets(GPT4, raw)
    t2 = time.time()
    return raw, offs, {"generate_s": round(t1 - t0, 2), "presplit_s": round(t2 - t1, 2), "chunks": int(offs.size)}


# ---------------------------------------------------------------------------------------------
def cpu_port_run(raw, offs, sample_bytes, steps, warmup=0):
    """Time `steps` iterations of the oracle's C restatement of the reference loop (regex.py:49-63)
    on the first `sample_bytes` of the corpus (cut at a chunk start).  Single thread."""
    import ctypes

    import oracle
    L = oracle.lib()
    k = int(np.searchsorted(offs, sample_bytes, side="left"))
    cut = int(offs[k]) if k < offs.size else min(sample_bytes, raw.size)
    ids = raw[:cut].astype(np.int32)
    start = np.zeros(max(cut, 1), dtype=np.uint8)
    start[offs[:k].astype(np.int64)] = 1
    if cut:
        start[0] = 1
    n = ctypes.c_uint64(cut)
    pair = (ctypes.c_int32 * 2)()
    cnt = ctypes.c_int64()
    done = 0
    for i in range(warmup):
        L.orc_train_step(ids.ctypes.data, start.ctypes.data, ctypes.byref(n), 256 + i, pair, ctypes.byref(cnt))
    t0 = time.perf_counter()
    for i in range(steps):
        rc = L.orc_train_step(ids.ctypes.data, start.ctypes.data, ctypes.byref(n), 256 + warmup + i, pair, ctypes.byref(cnt))
        if rc != 0:
            break
        done += 1
    dt = time.perf_counter() - t0
    return cut, done, dt


def python_reference_run(raw, nbytes=1 << 20, merges=8):
    """The UNMODIFIED pure-Python reference (vendored to oracle/_ref by oracle/make_ref.py) on the first `nbytes`
    of the corpus: RegexTokenizer.train(text, 256 + merges) — regex.py:36-70 with its own regex split.  One core
    (the reference is single-threaded).  None when the vendored copy is absent."""
    from oracle import make_ref
    ref = make_ref.load()
    if ref is None:
        return None
    cut = nbytes
    while cut < raw.size and (raw[cut] & 0xC0) == 0x80:
        cut += 1
    text = raw[:cut].tobytes().decode("utf-8")
    tok = ref.RegexTokenizer()
    t0 = time.perf_counter()
    tok.train(text, 256 + merges)
    dt = time.perf_counter() - t0
    return {"value": cut * merges / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference", "merges_per_s": merges / dt,
            "seconds": dt, "first_pairs": [list(p) for p in list(tok.merges)[:4]],
            "sample": f"karpathy/minbpe RegexTokenizer.train (pure Python, incl. its regex split) on the first {cut} bytes, {merges} merges"}


def _ref_worker(args):
    seed, shard, nbytes, steps, warmup = args
    from minbpe_b200.presplit import chunk_offsets_1proc
    from minbpe_b200.synth import generate
    import regex
    raw = generate(seed + 1000 * (shard + 1), nbytes, threads=1)
    offs = chunk_offsets_1proc(regex.compile(GPT4), raw.tobytes())
    return cpu_port_run(raw, offs, nbytes, steps, warmup)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle C port of base.py:13-41 + regex.py:49-63) on the host
    cores this process may use (affinity mask capped by the cgroup quota): one independent replica of the merge
    loop per core, each on its own 16 MiB shard of synthetic text.  Three passes, the median is reported, with the
    parallel efficiency against one replica running alone; the pure-Python reference itself is timed beside it."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = min(host_cores(), 64)
    shard = 16 << 20
    steps = max(1, min(args.steps, 8))
    warm = min(args.warmup, 1)
    t0 = time.perf_counter()
    solo = _ref_worker((args.seed, 0, shard, steps, warm))        # one replica alone: the per-core rate
    solo_rate = solo[0] * solo[1] / solo[2] / 1e9
    passes = []
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(3):
            res = pool.map(_ref_worker, [(args.seed, s, shard, steps, warm) for s in range(cores)])
            tmax = max(r[2] for r in res)
            passes.append((sum(r[0] * r[1] for r in res) / tmax / 1e9, tmax, sum(r[1] for r in res) / tmax / cores))
    wall = time.perf_counter() - t0
    value, tmax, merges_per_s = sorted(passes)[1]
    from minbpe_b200.synth import generate
    pyref = python_reference_run(generate(args.seed, 2 << 20, threads=1))
    sample = (f"{cores} independent single-thread replicas of the oracle C port (bpe_oracle.c orc_train_step), each "
              f"{steps} merge steps on its own 16 MiB synthetic shard (seed {args.seed}+1000*(shard+1)); time = slowest replica; "
              f"median of 3 passes")
    line = {
        "impl": "reference", "metric": "train_loop_corpus_GBps", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": tmax / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "RegexTokenizer.train merge loop, GPT-4 split, synthetic UTF-8 (BASELINE configs[2] shape), "
                               "bounded 16 MiB-per-core sample", "host_cores": cores, "os_cpu_count": os.cpu_count(),
                   "wall_s": round(wall, 2)},
        "merges_per_s": merges_per_s,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                         "passes_GBps": [p[0] for p in passes], "one_replica_GBps": solo_rate,
                         "parallel_efficiency": value / (solo_rate * cores), "python_reference": pyref},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def full_run(eng, raw, offs, merges, check=True):
    """BASELINE configs[2] to completion: bpe_load_text_gpt4 + bpe_train(all merges) from the host text, then every
    merge and count compared with the oracle's weighted loop over the distinct chunks of the host `regex` split
    (oracle.c_dedup_chunks + c_train(weights): same dict as regex.py:51-54 builds, tests/test_oracle.py)."""
    import torch
    from minbpe_b200 import engine as E
    eng.set_option(E.OPT_KERNEL_TIMING, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.load_text_gpt4(raw)
    t_load = time.perf_counter() - t0
    pairs, counts, done = eng.train(merges)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    tm = eng.timing()
    peak, _ = measured_peak()
    t_loop = tm["loop_ms"] / 1e3
    out = {"merges": int(done), "requested": int(merges), "seconds": t_all, "load_seconds": t_load, "loop_seconds": t_loop,
           "init_ms": tm["init_ms"], "merges_per_s": done / t_loop, "corpus_GBps": raw.size * done / t_loop / 1e9,
           "stream_GBps": 4.0 * tm["tokens_in"] / t_loop / 1e9,
           "fused_bytes_frac_of_peak": (4.0 * tm["tokens_in"] + 4.0 * tm["tokens_out"]) / t_loop / 1e9 / peak,
           "survey_8d_read_frac_of_peak": 8.0 * tm["tokens_in"] / t_loop / 1e9 / peak,
           "table_slots": int(tm["table_slots"]), "table_used": int(tm["table_used"]),
           "same_pairs": int(sum(1 for a, b in pairs.tolist() if a == b)), "final_tokens": int(eng.stream_len()),
           "gpu_launches": int(tm["kernel_launches"]), "end_to_end_corpus_MBps": raw.size / t_all / 1e6}
    if check:
        import oracle
        t0 = time.perf_counter()
        ub, uo, uw = oracle.c_dedup_chunks(raw, offs)
        wp, wc, wn = oracle.c_train(ub.astype(np.int32), uo, merges, weights=uw)
        out["oracle_seconds"] = time.perf_counter() - t0
        out["distinct_chunks"] = int(uo.size)
        out["parity_all_merges"] = bool(wn == done and np.array_equal(pairs, wp) and np.array_equal(counts, wc))
        if not out["parity_all_merges"]:
            k = min(len(pairs), len(wp))
            bad = np.flatnonzero((pairs[:k] != wp[:k]).any(axis=1) | (counts[:k] != wc[:k]))
            out["first_mismatch"] = int(bad[0]) if bad.size else k
    eng.set_option(E.OPT_KERNEL_TIMING, 1)
    return out, pairs


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    size = args.size_mib << 20
    K, W = args.steps, args.warmup

    if world > 1 or os.environ.get("BPE_BENCH_FORCE_SHARDED"):   # the env switch runs the sharded loop on one rank (tests)
        from minbpe_b200.dist import bench_sharded
        return bench_sharded(args, rank, world, local)

    # host-side preparation, before CUDA is touched (the pre-split forks worker processes)
    raw, offs, prep = make_corpus(size, args.seed)
    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        oracle.build()
        sample = min(size, 64 << 20)
        cut, done, dt = cpu_port_run(raw, offs, sample, 8, 0)
        cpu = {"value": cut * done / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": f"first {cut} bytes of the same corpus, {done} merge steps of oracle/bpe_oracle.c orc_train_step "
                         f"(C restatement of base.py:13-41 + regex.py:49-63), {dt:.1f} s, single thread",
               "merges_per_s": done / dt, "python_reference": python_reference_run(raw)}

    from minbpe_b200 import engine as E
    torch.cuda.set_device(local)
    eng = E.Engine(local)
    eng.set_option(E.OPT_KERNEL_TIMING, 1)

    sampler = ClockSampler(local)
    sampler.start()   # sampling runs from here; only the rows inside the timed region are reported
    # ---- e2e: C-ABI calls from host buffers (upload + device split + W+K merges + merges back) ----
    pinned = pin_host(raw)
    # untimed warm-up of the same calls (class tables, first-touch of the big device allocations, clocks)
    eng.load_text_gpt4(raw)
    eng.train(W)
    torch.cuda.synchronize()
    e2e_runs = []
    for _ in range(3):   # the wall clock of a 0.2 s region is noisy (allocator, PCIe): report the median run
        t0 = time.perf_counter()
        eng.load_text_gpt4(raw)
        load_tm = eng.timing()
        t_load = time.perf_counter() - t0
        pairs_e2e, _, done = eng.train(W + K)
        torch.cuda.synchronize()
        e2e_runs.append((time.perf_counter() - t0, t_load))
        assert done == W + K, "corpus ran out of pairs"
    tm_e2e = eng.timing()
    t_e2e, t_load = sorted(e2e_runs)[1]
    h2d = load_tm["h2d_bytes"]
    d2h = tm_e2e["d2h_bytes"]

    # ---- device-resident: W warm-up steps, then exactly K timed steps ----
    eng.load_stream(raw, offs)
    eng.train(W)
    torch.cuda.synchronize()
    sampler.begin()
    t0 = time.perf_counter()
    pairs, counts, done = eng.train(K, first_idx=256 + W)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sampler.end()
    clocks = sampler.stop()
    tm = eng.timing()
    assert done == K
    assert np.array_equal(pairs, pairs_e2e[W:W + K]), "timed run and e2e run disagree"
    t_loop = tm["loop_ms"] / 1e3          # CUDA events on the library's stream, around the K iterations
    value = size * K / t_loop / 1e9
    n_in, n_out = tm["tokens_in"], tm["tokens_out"]
    peak, peak_src = measured_peak()
    k_ms = tm["merge_kernel_ms"] / K
    achieved = (4.0 * n_in + 4.0 * n_out) / K / (k_ms / 1e3) / 1e9
    full = None
    if args.full_merges > 0:
        full, _ = full_run(eng, raw, offs, args.full_merges, check=not args.no_cpu_baseline)
    line = {
        "metric": "train_loop_corpus_GBps", "value": value, "unit": "GB/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": t_loop / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[2]: RegexTokenizer.train merge loop (GPT-4 split) on {args.size_mib} MiB synthetic "
                               f"UTF-8 seed {args.seed}, merge steps {W}..{W + K - 1} of 32512", "tokens_start": int(raw.size),
                   "chunks": prep["chunks"], "l2": "stream (>= 4 bytes/token, far larger than the 126 MB L2) is re-read from HBM every step",
                   "timing": "CUDA events on the library stream around the K enqueued iterations; wall-clock check in wall_ms_per_step",
                   "prep": prep},
        "merges_per_s": K / t_loop,
        "stream_GBps": 4.0 * n_in / t_loop / 1e9,
        "algorithmic_GBps_survey_8d": (8.0 * n_in + 4.0 * n_out) / t_loop / 1e9,
        "wall_ms_per_step": wall / K * 1e3,
        "gpu_launches": int(tm["kernel_launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_merge_seg (fused merge + in-place segment compaction + stats delta)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic()[0],
                     "traffic_capture": ncu_traffic()[1],
                     "peak_source": peak_src, "bytes_per_launch": (4.0 * n_in + 4.0 * n_out) / K, "ms_per_launch": k_ms,
                     "loop_frac_in_kernel": tm["merge_kernel_ms"] / tm["loop_ms"]},
        "cpu_baseline": cpu,
        "full_run": full,
        "e2e": {"value": size * (W + K) / t_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d / (W + K),
                "d2h_bytes_per_step": d2h / (W + K), "seconds": t_e2e, "merges": W + K,
                "load_seconds": t_load, "runs_seconds": [r[0] for r in e2e_runs], "host_buffer": "pinned (cudaHostRegister)" if pinned else "pageable",
                "what": "bpe_load_text_gpt4(host text: H2D + GPT-4 split on the device) + bpe_train(W+K) + merges D2H, wall clock, median of 3 runs"},
        "first_pairs": pairs[:4].tolist(),
    }
    eng.close()
    print(json.dumps(line), flush=True)


def run_extras(args):
    """--extras: side measurements recorded in profiles/ (not the contract line): cfg2 wall times and
    chunk-parallel encode throughput (BASELINE configs[1] and configs[4] shapes)."""
    import torch
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    from minbpe_b200 import engine as E
    from minbpe_b200.presplit import chunk_offsets
    from minbpe_b200.synth import generate
    out = {}
    text = open(os.path.join(ROOT, "tests", "golden", "taylorswift.txt"), encoding="utf-8").read()
    for name, cls in (("basic", BasicTokenizer), ("regex", RegexTokenizer)):
        tok = cls()
        tok.train(text, 300)  # warm-up (context, allocations)
        t0 = time.perf_counter(); tok.train(text, 512); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t1 = time.perf_counter(); ids = tok.encode(text); torch.cuda.synchronize(); de = time.perf_counter() - t1
        out[f"cfg2_{name}"] = {"train_wall_s": dt, "merges_per_s": 256 / dt, "loop_ms": tok.last_timing["loop_ms"],
                               "encode_wall_s": de, "n_ids": len(ids)}
    size = args.size_mib << 20
    raw = generate(1339, size)
    offs = chunk_offsets(GPT4, raw)
    eng = E.Engine(0)
    # a longer stretch of the cfg3 loop: dense early merges, sparse later ones, pairs (a,a), re-packing
    eng.load_stream(raw, offs)
    t0 = time.perf_counter(); mp, mc, md = eng.train(1024); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tm = eng.timing()
    out["train_1024"] = {"bytes": size, "merges": int(md), "wall_s": dt, "loop_ms": tm["loop_ms"], "init_ms": tm["init_ms"],
                         "merges_per_s": md / (tm["loop_ms"] / 1e3), "tokens_in_sum": tm["tokens_in"], "tokens_out_sum": tm["tokens_out"],
                         "stream_GBps": 4.0 * tm["tokens_in"] / (tm["loop_ms"] / 1e3) / 1e9, "same_pairs": int(sum(1 for a, b in mp.tolist() if a == b)),
                         "table_slots": tm["table_slots"], "final_tokens": int(eng.stream_len())}
    # the literal two-pass loop of the north star (full pair histogram every iteration: BPE_OPT_RESCAN)
    sub = 256 << 20
    eng.load_stream(raw[:sub], offs[: int(np.searchsorted(offs, sub))])
    eng.set_option(E.OPT_RESCAN, 1); eng.set_option(E.OPT_KERNEL_TIMING, 1)
    eng.train(3)
    _, _, rd = eng.train(8, first_idx=259)
    tr = eng.timing()
    eng.set_option(E.OPT_RESCAN, 0); eng.set_option(E.OPT_KERNEL_TIMING, 0)
    hist_ms = (tr["loop_ms"] - tr["merge_kernel_ms"]) / max(rd, 1)
    out["rescan_256MiB"] = {"merges": int(rd), "loop_ms_per_merge": tr["loop_ms"] / max(rd, 1), "merge_ms": tr["merge_kernel_ms"] / max(rd, 1),
                            "hist_argmax_ms": hist_ms, "hist_GBps": 4.0 * tr["tokens_in"] / max(rd, 1) / (hist_ms / 1e3) / 1e9}
    # device-side GPT-4 splitter: text bytes in, chunk offsets / marked stream out
    eng.split_gpt4(raw[: 16 << 20])  # warm-up (tables, allocations)
    eng.set_option(E.OPT_KERNEL_TIMING, 1)
    t0 = time.perf_counter(); got = eng.split_gpt4(raw); dt = time.perf_counter() - t0
    k_ms = eng.timing()["init_ms"]
    t0 = time.perf_counter(); eng.load_text_gpt4(raw); torch.cuda.synchronize(); dl = time.perf_counter() - t0
    eng.set_option(E.OPT_KERNEL_TIMING, 0)
    t0 = time.perf_counter(); eng.load_stream(raw, offs); torch.cuda.synchronize(); dh = time.perf_counter() - t0
    out["split_gpt4"] = {"bytes": size, "chunks": int(got.size), "equal_host_regex": bool(np.array_equal(got, offs)),
                         "offsets_wall_s": dt, "kernels_ms": k_ms, "kernels_GBps_text": size / (k_ms / 1e3) / 1e9,
                         "load_text_wall_s": dl, "load_stream_from_host_offsets_wall_s": dh}
    eng.load_stream(raw[: 64 << 20], offs[: int(np.searchsorted(offs, 64 << 20))])
    merges, _, done = eng.train(2048)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ids = eng.encode(raw, offs, merges); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tm = eng.timing()
    out["encode"] = {"bytes": size, "merges": int(done), "chunks": int(offs.size), "ids": int(ids.size), "wall_s": dt,
                     "GBps_e2e": size / dt / 1e9, "h2d_bytes": tm["h2d_bytes"], "d2h_bytes": tm["d2h_bytes"]}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="corpus bytes per GPU (MiB)")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-merges", type=int, default=32512,
                    help="N=1: also run the whole train() loop (this many merges, configs[2] = 32512) from the host text "
                         "and compare every merge with the oracle; 0 = skip")
    ap.add_argument("--extras", action="store_true", help="side measurements (cfg2 wall time, encode throughput)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.extras:
        run_extras(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
