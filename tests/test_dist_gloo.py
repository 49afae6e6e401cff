"""CPU, world_size 2 and 3 over gloo: the sharded training loop (minbpe_b200/dist.py) — shard cuts
at chunk starts, MIN-reduced candidate word for the first-occurrence tie-break across ranks,
SUM-reduced statistics delta, identical tables everywhere — must reproduce the single-stream
oracle merges and counts bit for bit.  Local work is done by tests/cpu_step_engine.py."""
import os
import sys

import numpy as np
import pytest
import regex
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import GOLDEN, ROOT

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def _worker(rank, world, port, text, num_merges, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_step_engine import CpuStepEngine
        from minbpe_b200.dist import ShardedTrainer, shard_chunks
        data, offs = oracle.split_to_stream(text, GPT4)
        blo, bhi, clo, chi = shard_chunks(data.size, offs, rank, world)
        local = offs[clo:chi] - np.uint64(blo)
        eng = CpuStepEngine(data[blo:bhi].tobytes(), local)
        tr = ShardedTrainer(eng, rank, world, poll_every=5)
        tr.prepare(num_merges)
        done, exhausted = tr.run()
        pairs, counts, n = tr.result()
        q.put((rank, pairs.tolist(), counts.tolist(), n, exhausted))
    finally:
        dist.destroy_process_group()


def run_world(world, text, num_merges, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, text, num_merges, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29612)])
def test_sharded_equals_single_stream(world, port):
    text = open(os.path.join(GOLDEN, "taylorswift.txt"), encoding="utf-8").read()[:6000]
    text += " the the the a a a zz zz qq qq qq"  # ties whose first occurrence sits in the last shard
    data, offs = oracle.split_to_stream(text, GPT4)
    num_merges = 40
    want_p, want_c, want_n = oracle.c_train(data.astype(np.int32), offs, num_merges)
    res = run_world(world, text, num_merges, port)
    for rank, pairs, counts, n, exhausted in res:
        assert n == want_n == num_merges and not exhausted
        assert pairs == want_p.tolist(), f"rank {rank}"
        assert counts == want_c.tolist(), f"rank {rank}"


def test_sharded_exhaustion():
    # the corpus runs out of pairs: every rank must stop at the same iteration (reference: ValueError)
    text = "ab cd ab cd ef"
    data, offs = oracle.split_to_stream(text, GPT4)
    want_p, want_c, want_n = oracle.c_train(data.astype(np.int32), offs, 30)
    res = run_world(2, text, 30, 29613)
    for rank, pairs, counts, n, exhausted in res:
        assert exhausted and n == want_n < 30
        assert pairs == want_p.tolist() and counts == want_c.tolist()


def test_shard_chunks_cover_and_respect_chunks():
    from minbpe_b200.dist import shard_chunks
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(1, 5000))
        k = int(rng.integers(1, min(n, 300) + 1))
        offs = np.unique(np.concatenate([[0], rng.integers(0, n, size=k)])).astype(np.uint64)
        for world in (1, 2, 3, 8):
            spans = [shard_chunks(n, offs, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1, c0, c1), (b0, b1, d0, d1) in zip(spans, spans[1:]):
                assert a1 == b0 and c1 == d0
            for lo, hi, c0, c1 in spans:
                assert lo <= hi and (lo == n or lo in offs)


def test_shard_byte_range_cuts_are_chunk_boundaries(taylorswift):
    """Byte-range shards (files, generated corpora): every interior cut must be a boundary of the regex split of the
    WHOLE text, for any world size — then the ranks' chunk lists concatenate to RegexTokenizer's (regex.py:41)."""
    import regex
    from minbpe_b200.dist import first_safe_cut, shard_byte_range
    from minbpe_b200.synth import generate
    from minbpe_b200.tokenizer import GPT2_SPLIT_PATTERN, GPT4_SPLIT_PATTERN
    texts = [taylorswift, generate(1337, 3 << 20).tobytes().decode("utf-8"),
             "word " * 10 + "   \n  spaces   and ends " * 2000 + "tail"]
    for pat in (GPT4_SPLIT_PATTERN, GPT2_SPLIT_PATTERN):
        cp = regex.compile(pat)
        for text in texts:
            raw = np.frombuffer(text.encode("utf-8"), dtype=np.uint8)
            whole = cp.findall(text)
            for world in (1, 2, 3, 8):
                parts, prev_hi = [], 0
                for r in range(world):
                    lo, hi = shard_byte_range(raw.size, r, world, lambda a, b: raw[a:b], window=1 << 16)
                    assert lo == prev_hi and lo <= hi
                    prev_hi = hi
                    parts += cp.findall(raw[lo:hi].tobytes().decode("utf-8"))
                assert prev_hi == raw.size
                assert parts == whole, (pat[:12], world)
    assert first_safe_cut(np.frombuffer(b"12 34 a b", dtype=np.uint8)) == 7
    assert first_safe_cut(np.frombuffer(b"  \n12", dtype=np.uint8)) == -1
    with pytest.raises(ValueError):
        shard_byte_range(100, 1, 2, lambda a, b: np.frombuffer(b"1" * (b - a), dtype=np.uint8))


def test_bench_unique_chunk_merge_equals_plain_oracle():
    """bench.py's cfg4 parity check merges per-piece / per-rank tables of distinct chunks: the weighted oracle loop
    over the merged table must equal the plain loop over the whole text (regex.py:51-54: one dict across chunks)."""
    import regex
    import bench
    from minbpe_b200.synth import generate
    from minbpe_b200.tokenizer import GPT4_SPLIT_PATTERN
    text = generate(1337, 1 << 20).tobytes().decode("utf-8")
    data, offs = oracle.split_to_stream(text, regex.compile(GPT4_SPLIT_PATTERN))
    want = oracle.c_train(data.astype(np.int32), offs, 80)
    k = len(offs)
    parts = []
    for a, b in ((0, k // 3), (k // 3, k // 2), (k // 2, k)):
        lo, hi = int(offs[a]), (int(offs[b]) if b < k else data.size)
        ub, uo, uw = oracle.c_dedup_chunks(data[lo:hi], offs[a:b] - offs[a])
        ends = np.append(uo[1:], ub.size).astype(np.int64)
        blob = ub.tobytes()
        parts.append(([blob[int(x):int(y)] for x, y in zip(uo.astype(np.int64), ends)], uw))
    chunks, weights = bench.merge_unique(parts)
    got = bench.oracle_train_unique(chunks, weights, 80)
    assert got[2] == want[2] == 80 and np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
