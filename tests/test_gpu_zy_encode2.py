"""GPU: the memoised chunk encode (k_encode2.cuh) — bpe_encode_text_gpt4 (split + encode on the device) and
bpe_encode with host offsets — against the oracle's restatement of regex.py:92-121 (oracle.c_encode), bit-exact ids."""
import os

import numpy as np
import pytest
import regex

import oracle

pytestmark = pytest.mark.gpu
SMALL = bool(os.environ.get("BPE_TEST_SMALL"))    # set by tests/test_emu.py: the same tests on the CPU SIMT emulator, smaller table

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


@pytest.fixture(scope="module")
def eng():
    from minbpe_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def trained(eng):
    """merges from 4 MiB of the synthetic corpus (600 merges) + its text"""
    from minbpe_b200.synth import generate
    n_merges = 250 if SMALL else 600
    text = generate(1337, (1 if SMALL else 4) << 20).tobytes().decode("utf-8")
    data, offs = oracle.split_to_stream(text, GPT4)
    eng.load_stream(data, offs)
    merges, _, done = eng.train(n_merges)
    assert done == n_merges
    return text, merges


def want(text, merges, perm=None):
    data, offs = oracle.split_to_stream(text, GPT4)
    return data, offs, oracle.c_encode(data, offs, merges, perm)


def test_fused_text_encode_vs_oracle(eng, trained, taylorswift):
    text, merges = trained
    long_bits = ("x" * 70 + " " + "ab" * 500 + " " + "=" * 3000 + "\n" + "lyiltumdya" * 300 + " " + "z" * 8000 + " end " +
                 " " * 100 + "\n" * 50 + "word")
    for t in (text, taylorswift, text[:300000] + long_bits + text[300000:600000], long_bits, "a", "ab", "hello world", " ", "日本語のテキスト " * 1000):
        data, offs, w = want(t, merges)
        got = eng.encode_text_gpt4(data.tobytes(), merges)
        assert np.array_equal(got, w), (len(t), len(got), len(w))
        got2 = eng.encode(data, offs, merges)             # same kernels, chunk starts from host offsets
        assert np.array_equal(got2, w)
    st = eng.encode_stats()
    assert st["memo_chunks"] > 1000 and st["fallback_pieces"] == 0
    assert eng.encode_text_gpt4(b"", merges).size == 0


def test_memo_is_warm_and_follows_the_merges(eng, trained):
    text, merges = trained
    data, offs, w = want(text[:1 << 20], merges)
    a = eng.encode_text_gpt4(data.tobytes(), merges)
    n1 = eng.encode_stats()
    b = eng.encode_text_gpt4(data.tobytes(), merges)       # second call: every chunk is already in the table
    n2 = eng.encode_stats()
    assert np.array_equal(a, w) and np.array_equal(b, w)
    assert n2["new_chunks"] == 0 and n2["memo_chunks"] == n1["memo_chunks"]
    assert n2["pool_ids"] == n1["pool_ids"]                  # repeated calls do not grow the persistent state
    # fewer merges -> different ids: the table must not serve entries of the old merges
    m2 = merges[:100]
    assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), m2), oracle.c_encode(data, offs, m2))
    assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges), w)
    # byte permutation (gpt4.py:76-77 style) is part of the key as well
    perm = np.random.default_rng(3).permutation(256).astype(np.uint8)
    assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges, perm), oracle.c_encode(data, offs, merges, perm))
    assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges), w)


def test_tiny_memo_table_overflows_to_the_direct_path(eng, trained):
    """A 64-slot table holds 32 chunks: nearly every chunk takes the direct list (thread per chunk + position map).
    Ids must not change; neither when the id pool / lists run out and a piece takes the general path."""
    from minbpe_b200 import engine as E
    text, merges = trained
    data, offs, w = want(text[:600000], merges)
    try:
        eng.set_option(E.OPT_ENC_MEMO_LOG2, 6)
        got = eng.encode_text_gpt4(data.tobytes(), merges)
        st = eng.encode_stats()
        assert np.array_equal(got, w)
        assert st["direct_chunks"] > 10000 and st["memo_chunks"] <= 64   # the fill limit is checked without a lock: approximate
        # the ids of directly encoded chunks live in a per-piece area (not in the memo's pool): it started too small for a
        # text that is all direct chunks, was grown, and the piece was done again
        assert st["repeated_pieces"] == 1 and st["direct_ids"] > 10000 and st["fallback_pieces"] == 0
        eng.set_option(E.OPT_SPLIT_PIECE, 1 << 16)         # several pieces per call share the (tiny) table
        assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges), w)
        assert eng.encode_stats()["pieces"] > 5
    finally:
        eng.set_option(E.OPT_ENC_MEMO_LOG2, 0)
        eng.set_option(E.OPT_SPLIT_PIECE, 0)
    eng.set_option(E.OPT_SPLIT_PIECE, 1 << 17)
    try:
        assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges), w)
    finally:
        eng.set_option(E.OPT_SPLIT_PIECE, 0)


def test_oversize_chunk_takes_the_general_path(eng, trained):
    text, merges = trained
    t = text[:100000] + " " + "lyiltumdya" * 1200 + " " + text[100000:200000]     # one chunk of 12,000 letters > 8192
    data, offs, w = want(t, merges)
    assert np.array_equal(eng.encode_text_gpt4(data.tobytes(), merges), w)
    assert eng.encode_stats()["fallback_pieces"] == 1
    # and the table works again afterwards
    d2, o2, w2 = want(text[:200000], merges)
    assert np.array_equal(eng.encode_text_gpt4(d2.tobytes(), merges), w2)
    assert eng.encode_stats()["fallback_pieces"] == 0


def test_tokenizer_encode_paths(trained):
    """RegexTokenizer.encode_ordinary / encode (specials) and BasicTokenizer.encode through the classes."""
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    text, merges = trained
    tok = RegexTokenizer()
    tok.merges = {(int(a), int(b)): 256 + i for i, (a, b) in enumerate(merges.tolist())}
    tok.vocab = tok._build_vocab()
    data, offs, w = want(text[:500000], merges)
    ids = tok.encode_ordinary(text[:500000])
    assert ids == w.tolist()
    assert tok.decode(ids) == text[:500000]
    tok.register_special_tokens({"<|endoftext|>": 100257})
    parts = text[:200000] + "<|endoftext|>" + text[200000:400000]
    got = tok.encode(parts, allowed_special="all")
    d1, o1, w1 = want(text[:200000], merges)
    d2, o2, w2 = want(text[200000:400000], merges)
    assert got == w1.tolist() + [100257] + w2.tolist()
    short = "hello world, this is short"
    ds, os_, ws = want(short, merges)
    assert tok.encode(short) == ws.tolist()
    b = BasicTokenizer()
    b.merges = dict(tok.merges)
    b.vocab = b._build_vocab()
    raw = text[:20000].encode("utf-8")
    assert b.encode(text[:20000]) == oracle.c_encode(np.frombuffer(raw, dtype=np.uint8), None, merges).tolist()
