"""GPU: the step-wise C ABI (bpe_step_*) behind minbpe_b200.dist.ShardedTrainer.
world = 1 runs in-process on cuda:0 (every kernel of the sharded loop, no collective needed);
world = 2 spawns two ranks over NCCL when the box has two GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest
import regex

import oracle
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def corpus():
    from minbpe_b200.synth import generate
    text = generate(1337, 2 << 20).tobytes().decode("utf-8")
    text += " the the the a a a zz zz qq qq qq"
    return oracle.split_to_stream(text, GPT4)


@pytest.mark.parametrize("exchange", ["p2p", "collective"])
def test_step_api_world1(exchange):
    """p2p = the NVLink exchange kernels (k_xchg_cand / k_xchg_apply) with this rank as its only peer;
    collective = the host-sequenced step API (no collective needed at world 1)."""
    import torch
    from minbpe_b200.dist import GpuStepEngine, ShardedTrainer
    from minbpe_b200.engine import Engine
    data, offs = corpus()
    eng = Engine(0)
    eng.load_stream(data, offs)
    tr = ShardedTrainer(GpuStepEngine(eng, 0), rank=0, world=1, poll_every=7, exchange=exchange)
    tr.prepare(60)
    done, exhausted = tr.run()
    pairs, counts, n = tr.result()
    torch.cuda.synchronize()
    wp, wc, wn = oracle.c_train(data.astype(np.int32), offs, 60)
    assert n == wn == 60 and not exhausted
    assert np.array_equal(pairs, wp) and np.array_equal(counts, wc)
    eng.close()


def _rank_main(rank, world, port, q, exchange="p2p"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from minbpe_b200.dist import train_sharded
        from minbpe_b200.engine import Engine
        data, offs = corpus()
        eng = Engine(rank)
        pairs, counts, n = train_sharded(eng, rank, data, offs, 60, poll_every=7, exchange=exchange)
        # a second run on the same engine: the exchange block is reused (sequence numbers keep growing)
        p2, c2, n2 = train_sharded(eng, rank, data, offs, 30, poll_every=16, exchange=exchange)
        assert n2 == 30 and np.array_equal(p2, pairs[:30]) and np.array_equal(c2, counts[:30])
        q.put((rank, pairs.tolist(), counts.tolist(), n))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "collective"])
def test_step_api_world2_nccl(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29621 + (exchange == "p2p"), q, exchange)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data, offs = corpus()
    wp, wc, wn = oracle.c_train(data.astype(np.int32), offs, 60)
    for rank, pairs, counts, n in out:
        assert n == wn == 60 and pairs == wp.tolist() and counts == wc.tolist(), rank
