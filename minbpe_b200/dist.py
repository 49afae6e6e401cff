"""
Sharded BPE training: one process per GPU, torch.distributed for the plumbing (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).

The corpus is cut into contiguous byte ranges at chunk starts (regex chunks never interact,
regex.py:51-54,60), rank r holding the r-th range, so rank order = text order.  Every rank keeps an
identical copy of the global pair-count table; per merge iteration the ranks exchange

  * one int64 reduced with MIN  — the arg-max candidate.  The max count is the same everywhere
    (same table); a tie is broken like the reference (first occurrence in the stream, basic.py:35)
    = the candidate of the lowest rank that sees one of the tied pairs, each rank offering the
    tied pair that occurs first in its own shard.  word = rank << 58 | p0 << 29 | p1.
  * the statistics delta vector (2V+1 uint64: L[x], R[x], ZZ — DESIGN.md) reduced with SUM.

`ShardedTrainer` only sequences these steps; the work is in the engine's step API
(include/b200bpe.h "step-wise training").  Tests drive the same class over gloo with a CPU
stand-in engine built on the oracle (tests/test_dist_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist

CAND_NONE = (1 << 63) - 1


def shard_chunks(n_bytes, offsets, rank, world):
    """Contiguous chunk range of rank `rank`: (byte_lo, byte_hi, chunk_lo, chunk_hi), balanced by
    bytes, cut only at chunk starts."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    k = len(offsets)
    cuts = [int(np.searchsorted(offsets, n_bytes * r // world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, k
    lo, hi = cuts[rank], cuts[rank + 1]
    byte_lo = int(offsets[lo]) if lo < k else n_bytes
    byte_hi = int(offsets[hi]) if hi < k else n_bytes
    return byte_lo, byte_hi, lo, hi


class GpuStepEngine:
    """Adapter: minbpe_b200.engine.Engine step API over torch CUDA tensors on the current stream."""

    def __init__(self, engine, device):
        self.e = engine
        self.device = torch.device("cuda", device)
        # a dedicated torch stream shared by our kernels and (through torch's stream semantics) the
        # NCCL collectives: the legacy default stream has handle 0, which bpe_set_stream reads as
        # "use the handle's own stream" and which would leave kernels and collectives unordered
        self.stream = torch.cuda.Stream(device=self.device)
        self.e.set_stream(self.stream.cuda_stream)

    def stream_ctx(self):
        return torch.cuda.stream(self.stream)

    def new_i64(self, n):
        return torch.zeros(n, dtype=torch.int64, device=self.device)

    def begin(self, dense):
        self.e.step_begin(dense.data_ptr())

    def table(self, dense, num_merges, first_idx, poll_every):
        self.e.step_table(dense.data_ptr(), num_merges, first_idx, poll_every)
        return self.e.step_delta_len()

    def select(self, cand, rank):
        self.e.step_select(cand.data_ptr(), rank)

    def merge(self, cand, delta):
        self.e.step_merge(cand.data_ptr(), delta.data_ptr())

    def apply(self, delta):
        self.e.step_apply(delta.data_ptr())

    def poll(self):
        return self.e.step_poll()

    def result(self, cap):
        return self.e.step_result(cap)


class ShardedTrainer:
    """regex.py:49-63 over `world` shards.  `eng` is a step engine whose stream already holds this
    rank's shard (Engine.load_stream)."""

    def __init__(self, eng, rank=None, world=None, group=None, poll_every=16):
        self.eng, self.group = eng, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.poll_every = poll_every

    def _allreduce(self, t, op):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)

    def _ctx(self):
        import contextlib
        return self.eng.stream_ctx() if hasattr(self.eng, "stream_ctx") else contextlib.nullcontext()

    def prepare(self, num_merges, first_idx=256):
        """Iteration-0 statistics: local histograms, SUM across ranks, identical tables."""
        self.num_merges, self.first_idx = num_merges, first_idx
        with self._ctx():
            dense = self.eng.new_i64(65536)
            self.eng.begin(dense)
            self._allreduce(dense, dist.ReduceOp.SUM)
            n_delta = self.eng.table(dense, num_merges, first_idx, self.poll_every)
            self.cand = self.eng.new_i64(2)
            self.delta = self.eng.new_i64(n_delta)
        self.done = 0

    def run(self, num_steps=None):
        """Enqueue merge iterations (all of them by default); returns (iterations done, exhausted)."""
        target = self.num_merges if num_steps is None else min(self.num_merges, self.done + num_steps)
        exhausted = False
        with self._ctx():
            while self.done < target and not exhausted:
                k = min(self.poll_every, target - self.done)
                for _ in range(k):
                    self.eng.select(self.cand, self.rank)
                    self._allreduce(self.cand[:1], dist.ReduceOp.MIN)
                    self.eng.merge(self.cand, self.delta)
                    self._allreduce(self.delta, dist.ReduceOp.SUM)
                    self.eng.apply(self.delta)
                self.done, exhausted = self.eng.poll()
        return self.done, exhausted

    def result(self):
        return self.eng.result(self.num_merges)


def train_sharded(engine, device, data, offsets, num_merges, first_idx=256, group=None, poll_every=16):
    """Convenience: every rank passes the FULL corpus description (bytes + chunk offsets); the rank's
    shard is cut out, uploaded and trained.  Returns (pairs, counts, n_done)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    blo, bhi, clo, chi = shard_chunks(raw.size, offsets, rank, world)
    local_offs = (np.asarray(offsets[clo:chi], dtype=np.uint64) - np.uint64(blo))
    engine.load_stream(raw[blo:bhi], local_offs if len(local_offs) else None)
    tr = ShardedTrainer(GpuStepEngine(engine, device), rank, world, group, poll_every)
    tr.prepare(num_merges, first_idx)
    tr.run()
    return tr.result()


# -------------------------------------------------------------------------------------------------
def bench_sharded(args, rank, world, local):
    """bench.py --gpus N>1: weak scaling, every rank trains on its own `size_mib` shard of one
    N*size_mib corpus (rank r = r-th contiguous range; synthetic text, seed + r), with the per-merge
    NCCL exchanges described above.  Rank 0 prints the JSON line."""
    import json
    import os
    import time

    from . import engine as E
    from .presplit import chunk_offsets
    from .synth import generate
    from .tokenizer import GPT4_SPLIT_PATTERN

    size = args.size_mib << 20
    K, W = args.steps, args.warmup
    t0 = time.time()
    # rank r holds MiB [r*size_mib, (r+1)*size_mib) of ONE corpus (one lexicon); rank 0's shard is the N=1 workload
    raw = generate(args.seed, size, threads=max(1, (os.cpu_count() or 8) // world), first_block=rank * args.size_mib)
    offs = chunk_offsets(GPT4_SPLIT_PATTERN, raw, workers=max(1, min(64, (os.cpu_count() or 8) // world)))
    prep_s = time.time() - t0

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = E.Engine(local)
    eng.set_option(E.OPT_KERNEL_TIMING, 1)
    step = GpuStepEngine(eng, local)
    sampler = None
    if rank == 0:
        from bench import ClockSampler
        sampler = ClockSampler(local)
        sampler.start()   # sampling runs from here; only the rows inside the timed region are reported

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: host buffers -> merges, through the C ABI + collectives, wall clock ----
    try:
        pinned = int(torch.cuda.cudart().cudaHostRegister(raw.ctypes.data, raw.nbytes, 0)) == 0
    except Exception:  # noqa: BLE001
        pinned = False
    eng.load_text_gpt4(raw)          # untimed warm-up of the load path (class tables, first touch of the allocations)
    sync_all()
    t0 = time.perf_counter()
    eng.load_text_gpt4(raw)          # H2D of the shard's text + GPT-4 split on the device
    h2d = eng.timing()["h2d_bytes"]
    tr = ShardedTrainer(step, rank, world, poll_every=16)
    tr.prepare(W + K)
    tr.run()
    pairs_e2e, _, n_e2e = tr.result()
    sync_all()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)

    # ---- device-resident: W warm-up merges, then exactly K timed ----
    eng.load_stream(raw, offs)
    P = 8   # extra merges after the timed ones, with CUDA events between the phases of every step
    tr = ShardedTrainer(step, rank, world, poll_every=16)
    tr.prepare(W + K + P)
    tr.run(W)
    sync_all()
    if sampler:
        sampler.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(step.stream)
    t0 = time.perf_counter()
    done, exhausted = tr.run(K)
    ev1.record(step.stream)
    sync_all()
    wall = time.perf_counter() - t0
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    # where a step's time goes: select | all-reduce MIN | merge pass | all-reduce SUM | apply (rank 0's view;
    # a collective's share includes waiting for the slower rank)
    sync_all()   # rank 0 has just spent 0.1 s stopping its clock sampler: do not bill that to rank 1's first collective
    marks = []
    with tr._ctx():
        for _ in range(P):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            e[0].record(); tr.eng.select(tr.cand, tr.rank)
            e[1].record(); tr._allreduce(tr.cand[:1], dist.ReduceOp.MIN)
            e[2].record(); tr.eng.merge(tr.cand, tr.delta)
            e[3].record(); tr._allreduce(tr.delta, dist.ReduceOp.SUM)
            e[4].record(); tr.eng.apply(tr.delta)
            e[5].record(); marks.append(e)
        tr.done, _ = tr.eng.poll()
    sync_all()
    names = ["select", "allreduce_min", "merge", "allreduce_sum", "apply"]
    phases = {nm: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in marks[1:]])) for i, nm in enumerate(names)}
    all_phases = [None] * world
    dist.all_gather_object(all_phases, phases)
    t_loop = torch.tensor([ev0.elapsed_time(ev1) / 1e3], device="cuda")   # kernels + collectives share torch's stream
    dist.all_reduce(t_loop, op=dist.ReduceOp.MAX)
    pairs, counts, n = tr.result()
    tm = eng.timing()
    ok = (n == W + K + P) and np.array_equal(pairs[: W + K], pairs_e2e)
    if rank == 0:
        from bench import measured_peak
        t = float(t_loop.item())
        peak, peak_src = measured_peak()
        # token counters cover the whole W+K run of rank 0.  The per-launch time is taken from the step time
        # (merge pass + arg-max + both collectives), i.e. a lower bound on the kernel's own rate: CUDA events
        # around single launches are not meaningful on a stream that NCCL work is interleaved with.
        k_ms = t / K * 1e3
        bytes_per_launch = 4.0 * (tm["tokens_in"] + tm["tokens_out"]) / max(n, 1)
        line = {
            "metric": "train_loop_corpus_GBps", "value": size * world * K / t / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": t / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"RegexTokenizer.train merge loop (GPT-4 split), {args.size_mib} MiB synthetic UTF-8 per GPU "
                                   f"(seed {args.seed}), contiguous shards of one {args.size_mib * world} MiB corpus, "
                                   f"merge steps {W}..{W + K - 1}; per merge: NCCL all-reduce MIN (8 B) + SUM (delta vector)",
                       "parallelism": f"shard{world}", "prep_s": round(prep_s, 1), "consistent": bool(ok),
                       "l2": "per-GPU stream >> 126 MB L2, re-read from HBM every step",
                       "timing": "CUDA events on the shared torch stream, max over ranks, barrier + synchronize on both sides"},
            "merges_per_s": K / t, "wall_ms_per_step": wall / K * 1e3,
            "gpu_launches": int(tm["kernel_launches"]),
            "clocks": clocks, "phases_ms": {f"rank{r}": ph for r, ph in enumerate(all_phases)},
            "roofline": {"bound": "hbm", "kernel": "k_merge_seg (rank 0; rate over the whole step incl. collectives)", "achieved": bytes_per_launch / (k_ms / 1e3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": bytes_per_launch / (k_ms / 1e3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "ms_per_launch": k_ms},
            "cpu_baseline": None,
            "e2e": {"value": size * world * (W + K) / float(t_e2e.item()) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": h2d / (W + K), "d2h_bytes_per_step": 16.0, "seconds": float(t_e2e.item()),
                    "host_buffer": "pinned (cudaHostRegister)" if pinned else "pageable",
                    "what": "per rank: bpe_load_text_gpt4(host shard text: H2D + device split) + sharded loop of W+K merges + merges D2H, wall clock, max over ranks"},
            "first_pairs": pairs[W:W + 4].tolist(),
        }
        print(json.dumps(line), flush=True)
    eng.close()
    dist.destroy_process_group()


def encode_sharded(engine, data, offsets, merges, byte_perm=None, group=None, gather=False):
    """regex.py:111-121 over `world` GPUs: chunks are independent, so encode needs no exchange at all —
    every rank encodes its contiguous chunk range (replicas over byte-range shards).  Returns this
    rank's ids; with gather=True rank 0 also gets the concatenation in text order (others: None)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    blo, bhi, clo, chi = shard_chunks(raw.size, offsets, rank, world)
    local_offs = np.asarray(offsets[clo:chi], dtype=np.uint64) - np.uint64(blo)
    ids = engine.encode(raw[blo:bhi], local_offs if len(local_offs) else None, merges, byte_perm) if bhi > blo else np.zeros(0, np.int32)
    if not gather:
        return ids
    parts = [None] * world if rank == 0 else None
    dist.gather_object(ids, parts, dst=0, group=group)
    return np.concatenate(parts) if rank == 0 else None
