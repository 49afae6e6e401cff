#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (run by tests/test_emu.py in a subprocess with BPE_LIB_PATH = the emulator build).

The sharded training loop on N emulated GPUs: one OS thread per rank, each with its own bpe handle (the emulator
reports 8 devices), driving minbpe_b200.dist.ShardedTrainer through the real C ABI step functions.

  exchange = "collective"   bpe_step_select / merge / apply with the two per-merge all-reduces done here between the
                            threads (what NCCL does on the GPUs)
  exchange = "p2p"          bpe_xchg_create / attach / probe + bpe_step_fused: the NVLink peer-memory kernels
                            (k_xchg_cand, k_xchg_apply) really wait for each other across the threads; "peer memory" is
                            the other thread's allocation, "CUDA IPC handles" carry the pointer

Every rank's merges and counts must equal oracle.c_train on the whole corpus.  This pins the LOGIC of the exchange
protocol (flags, double-buffered delta vectors, tie-break across shards, exchange-block reuse by a second run); the
system-scope memory model and CUDA IPC themselves can only be validated on NVLink hardware.
"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import regex  # noqa: E402

import oracle  # noqa: E402
from minbpe_b200.dist import ShardedTrainer, shard_chunks  # noqa: E402
from minbpe_b200.engine import Engine, EngineError  # noqa: E402

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


class Team:
    """What torch.distributed provides on the GPUs, between threads."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def allreduce(self, rank, arr, op):
        self.slots[rank] = arr.copy()
        self.barrier.wait()
        parts = np.stack(self.slots)
        res = parts.sum(axis=0, dtype=arr.dtype) if op == "sum" else parts.min(axis=0)
        self.barrier.wait()
        arr[...] = res

    def gather(self, rank, obj):
        self.slots[rank] = obj
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()
        return out


class EmuStepEngine:
    """Counterpart of minbpe_b200.dist.GpuStepEngine for the emulator: "device" buffers are numpy arrays."""

    def __init__(self, engine, team, rank):
        self.e, self.team, self.rank = engine, team, rank

    def new_i64(self, n):
        return np.zeros(n, dtype=np.int64)

    def begin(self, dense):
        self.e.step_begin(dense.ctypes.data)

    def table(self, dense, num_merges, first_idx, poll_every):
        self.e.step_table(dense.ctypes.data, num_merges, first_idx, poll_every)
        return self.e.step_delta_len()

    def select(self, cand, rank):
        self.e.step_select(cand.ctypes.data, rank)

    def merge(self, cand, delta):
        self.e.step_merge(cand.ctypes.data, delta.ctypes.data)

    def apply(self, delta):
        self.e.step_apply(delta.ctypes.data)

    def poll(self):
        return self.e.step_poll()

    def result(self, cap):
        return self.e.step_result(cap)

    def sync_ranks(self, group=None):
        self.team.barrier.wait()

    def fused(self, n_iters):
        self.e.step_fused(n_iters)

    def xchg_setup(self, world, rank, vocab_cap, group=None):
        key = (world, rank, vocab_cap)
        if getattr(self, "_key", None) == key:
            return True
        self._key = None
        if world > 1:
            self.e.xchg_detach()
            self.team.barrier.wait()
        mine = self.e.xchg_create(world, rank, vocab_cap)
        if world > 1:
            handles = np.stack(self.team.gather(rank, mine))
            self.e.xchg_attach(handles)
            self.team.barrier.wait()
            if not self.e.xchg_probe(3000):
                return False
            self.team.barrier.wait()
        self._key = key
        return True


class EmuTrainer(ShardedTrainer):
    def _allreduce(self, t, op):
        import torch.distributed as dist
        if self.world > 1:
            self.eng.team.allreduce(self.rank, t, "sum" if op == dist.ReduceOp.SUM else "min")


def corpus(nbytes):
    if nbytes < 0:
        # a small text whose merge counts fall to 2..3 quickly: nearly every arg-max is tied, and the tied pairs' first
        # occurrences are spread over all shards (the cross-shard tie-break decides almost every merge)
        rng = np.random.default_rng(5)
        words = ["".join(rng.choice(list("abcdeé"), size=int(rng.integers(1, 7)))) for _ in range(-nbytes // 4)]
        return oracle.split_to_stream(" ".join(words) + " zz zz qq qq qq", GPT4)
    from minbpe_b200.synth import generate
    text = generate(1337, 1 << 20)[:nbytes].tobytes().decode("utf-8", errors="ignore")
    text += " the the the a a a zz zz qq qq qq"     # ties between pairs that live in the last shard only
    return oracle.split_to_stream(text, GPT4)


def rank_main(rank, world, team, exchange, data, offs, merges, out, errs):
    try:
        blo, bhi, clo, chi = shard_chunks(data.size, offs, rank, world)
        local = np.asarray(offs[clo:chi], dtype=np.uint64) - np.uint64(blo)
        eng = Engine(rank)
        step = EmuStepEngine(eng, team, rank)
        res = []
        # second run: the exchange block is reused (sequence numbers keep growing); third: re-created for another vocabulary
        for m, poll in ((merges, 7), (merges, 16), (merges // 2, 16)):
            eng.load_stream(data[blo:bhi], local if len(local) else None)
            tr = EmuTrainer(step, rank, world, None, poll, exchange)
            tr.prepare(m)
            tr.run()
            pairs, counts, n = tr.result()
            res.append((pairs.copy(), counts.copy(), n, tr.exchange))
        team.barrier.wait()
        if exchange == "p2p" and world > 1:
            eng.xchg_detach()
            team.barrier.wait()
        eng.close()
        out[rank] = res
    except Exception as ex:  # noqa: BLE001
        errs.append((rank, repr(ex)))
        team.barrier.abort()


def run(world, exchange, nbytes, merges):
    data, offs = corpus(nbytes)
    team = Team(world)
    out, errs = [None] * world, []
    threads = [threading.Thread(target=rank_main, args=(r, world, team, exchange, data, offs, merges, out, errs)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    wp, wc, wn = oracle.c_train(data.astype(np.int32), offs, merges)     # wn < merges: the text ran out of pairs
    for r in range(world):
        for i, (p, c, n, ex) in enumerate(out[r]):
            want = min(wn, merges if i < 2 else merges // 2)
            assert ex == exchange, (ex, exchange)
            assert n == want and np.array_equal(p, wp[:want]) and np.array_equal(c, wc[:want]), f"rank {r}, run {i}: differs from the oracle"
    ties = int(np.sum(wc[1:] == wc[:-1]))
    print(f"world {world} {exchange}: {wn} of {merges} merges x {data.size} bytes bit-exact on every rank ({ties} equal-count neighbours in the run)")


if __name__ == "__main__":
    nbytes = int(os.environ.get("EMU_SHARD_BYTES", 192 << 10))
    merges = int(os.environ.get("EMU_SHARD_MERGES", 60))
    cases = [(2, "collective"), (2, "p2p"), (3, "p2p"), (4, "p2p")]
    if len(sys.argv) > 1:
        cases = [(int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[1:]]
    for world, exchange in cases:
        run(world, exchange, nbytes, merges)
        run(world, exchange, -6000, 260)      # tie-heavy: counts of 2..3
        run(world, exchange, -600, 400)       # ... and to exhaustion: every chunk ends as one token, "no pair left" on all ranks
    print("emu sharded ok")
