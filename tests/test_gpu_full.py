"""GPU: BASELINE configs[2] (cfg3) END TO END through the Python class, every merge checked.

    RegexTokenizer().train(<1 GiB synthetic UTF-8, seed 1337>, vocab_size=32768)     (regex.py:36-70)

The oracle side: host `regex` split of the same text -> distinct chunks in first-occurrence order with
multiplicities (oracle.c_dedup_chunks) -> oracle.c_train(weights=...), which
tests/test_oracle.py::test_dedup_weights_equal_plain pins to the plain loop and to the reference.  ALL
32,512 merges and their counts must agree (table growth, re-packing, the (a,a) path and the tie-breaks
of a complete run are all inside).  A second test encodes a 64 MiB slice of another corpus with the resulting
32k-entry table and compares it with oracle.c_encode (regex.py:92-121).

Sizes can be reduced for a quick run: BPE_FULL_MIB (default 1024), BPE_FULL_MERGES (default 32512)."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

SIZE_MIB = int(os.environ.get("BPE_FULL_MIB", "1024"))
MERGES = int(os.environ.get("BPE_FULL_MERGES", "32512"))


_RUN = {}     # the trained tokenizer of the first test, reused by the second (training 1 GiB twice would double the run time)


def _train_cfg3():
    if "tok" not in _RUN:
        from minbpe_b200 import RegexTokenizer
        from minbpe_b200.synth import generate
        raw = generate(1337, SIZE_MIB << 20)
        tok = RegexTokenizer()
        tok.train(raw.tobytes().decode("utf-8"), 256 + MERGES)                      # device split + device loop
        _RUN.update(tok=tok, raw=raw, timing=tok.last_timing)
    return _RUN["tok"], _RUN["raw"], _RUN["timing"]


def test_cfg3_full_run_all_merges_vs_oracle():
    from minbpe_b200.presplit import chunk_offsets
    from minbpe_b200.tokenizer import GPT4_SPLIT_PATTERN
    tok, raw, tm = _train_cfg3()
    got = np.array(list(tok.merges.keys()), dtype=np.int32)
    assert got.shape == (MERGES, 2)
    assert list(tok.merges.values()) == list(range(256, 256 + MERGES))
    # oracle: host regex split (third-party `regex`, as the reference), dedup, weighted loop
    offs = chunk_offsets(GPT4_SPLIT_PATTERN, raw, workers=min(16, len(os.sched_getaffinity(0))))
    ub, uo, uw = oracle.c_dedup_chunks(raw, offs)
    assert int(uw.sum()) == offs.size
    wp, wc, wn = oracle.c_train(ub.astype(np.int32), uo, MERGES, weights=uw)
    assert wn == MERGES
    _RUN["wp"] = wp
    bad = np.flatnonzero((got != wp).any(axis=1))
    assert bad.size == 0, f"first differing merge {int(bad[0])}: got {got[bad[0]].tolist()} want {wp[bad[0]].tolist()}"
    # counts of every merge (stats[pair] before the merge, basic.py:45) through the engine API
    eng = tok.engine
    eng.load_text_gpt4(raw)
    p2, c2, d2 = eng.train(MERGES)
    assert d2 == MERGES and np.array_equal(p2, wp) and np.array_equal(c2, wc)
    # token accounting of the whole run: sum over merges of replacements == n0 - final length
    tm2 = eng.timing()
    assert tm2["tokens_in"] - tm2["tokens_out"] == raw.size - eng.stream_len()
    assert tm["tokens_in"] == tm2["tokens_in"]
    # vocab of the class: vocab[idx] = vocab[p0] + vocab[p1] (basic.py:42)
    for i in (0, MERGES // 2, MERGES - 1):
        a, b = wp[i].tolist()
        assert tok.vocab[256 + i] == tok.vocab[a] + tok.vocab[b]


def test_cfg3_table_encodes_another_corpus():
    """A 64 MiB slice of a DIFFERENT corpus (seed 1339 = cfg5's) encoded with the full cfg3 table, bit-exact ids
    (regex.py:111-121 through the class: fused split + memoised encode on the device), and decoded back."""
    from minbpe_b200.presplit import chunk_offsets
    from minbpe_b200.synth import generate
    from minbpe_b200.tokenizer import GPT4_SPLIT_PATTERN
    tok, _, _ = _train_cfg3()
    merges = _RUN.get("wp")
    if merges is None:
        merges = np.array(list(tok.merges.keys()), dtype=np.int32)
    other = generate(1339, min(64, 32 * SIZE_MIB) << 20)
    o2 = chunk_offsets(GPT4_SPLIT_PATTERN, other, workers=min(16, len(os.sched_getaffinity(0))))
    text = other.tobytes().decode("utf-8")
    ids = np.asarray(tok.encode_ordinary(text), dtype=np.int32)
    want = oracle.c_encode(other, o2, merges)
    assert np.array_equal(ids, want)
    head = tok.decode(ids[:100000].tolist())
    assert head == text[: len(head)]


def test_load_ids_above_255_then_train():
    """bpe_load_ids accepts any id; the dense delta vector of bpe_train must cover them."""
    from minbpe_b200.engine import Engine
    rng = np.random.default_rng(11)
    ids = rng.choice(np.array([5, 300, 70000, 123456, 70001], dtype=np.int32), size=50000).astype(np.int32)
    offs = np.unique(np.concatenate([[0], rng.integers(1, ids.size, 4000)])).astype(np.uint64)
    eng = Engine(0)
    eng.load_ids(ids, offs)
    first = 200000
    gp, gc, gd = eng.train(20, first_idx=first)
    wp, wc, wd = oracle.c_train(ids, offs, 20, first_idx=first)
    assert gd == wd and np.array_equal(gp, wp) and np.array_equal(gc, wc)
    table = eng.debug_table()
    sp, sc = eng.get_stats()
    assert table == {(int(a), int(b)): int(c) for (a, b), c in zip(sp, sc)}
    eng.close()


def test_encode_rejects_bad_offsets():
    from minbpe_b200.engine import Engine, EngineError
    eng = Engine(0)
    merges = np.array([[97, 98]], dtype=np.int32)
    with pytest.raises(EngineError):
        eng.encode(b"abababab", np.array([2, 4], dtype=np.uint64), merges)      # offsets[0] != 0
    with pytest.raises(EngineError):
        eng.encode(b"abababab", np.array([0, 4, 4], dtype=np.uint64), merges)   # not strictly increasing
    with pytest.raises(EngineError):
        eng.encode(b"abababab", np.array([0, 9], dtype=np.uint64), merges)      # beyond the text
    assert eng.encode(b"abababab", np.array([0, 4], dtype=np.uint64), merges).tolist() == [256, 256, 256, 256]
    eng.close()
