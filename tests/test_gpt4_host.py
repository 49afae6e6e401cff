"""CPU: GPT4Tokenizer's host side (gpt4.py:11-46 recover_merges, :57-80 construction) on a synthetic tiktoken-style rank
table — built from merges the REFERENCE restatement (oracle.pyref) trains, with shuffled single-byte ranks — since
cl100k_base itself needs a network download.  The device side is tests/test_gpu_zz_gpt4.py."""
import numpy as np
import pytest


def synthetic_ranks(text, n_merges, seed=3):
    """(mergeable_ranks, merges in permuted id space, perm): what tiktoken would ship for a tokenizer with these merges."""
    from oracle import pyref
    rng = np.random.default_rng(seed)
    perm = rng.permutation(256)                       # rank of byte b = perm[b]
    ids = [int(perm[b]) for b in text.encode("utf-8")]
    merges = [tuple(pair) for pair, _count in pyref.train([ids], n_merges)[0]]   # [(p0, p1)] in the permuted byte space, rank order
    inv = {int(perm[b]): b for b in range(256)}
    tok_bytes = {i: bytes((inv[i],)) for i in range(256)}
    ranks = {bytes((b,)): int(perm[b]) for b in range(256)}
    ranks = dict(sorted(ranks.items(), key=lambda kv: kv[1]))
    for r, (a, b) in enumerate(merges):
        tok_bytes[256 + r] = tok_bytes[a] + tok_bytes[b]
        ranks[tok_bytes[256 + r]] = 256 + r
    return ranks, merges, perm


def test_recover_merges_inverts_a_rank_table(taylorswift):
    from minbpe_b200.gpt4 import recover_merges
    ranks, merges, _ = synthetic_ranks(taylorswift[:20000], 150)
    if len(ranks) != 256 + len(merges):
        pytest.skip("two merges produced the same byte string: not a tiktoken-style table")
    got = recover_merges(ranks)
    assert list(got.keys()) == [tuple(m) for m in merges]
    assert list(got.values()) == list(range(256, 256 + len(merges)))


def test_construction_and_refusals(taylorswift, tmp_path):
    from minbpe_b200 import GPT4Tokenizer
    from minbpe_b200.gpt4 import GPT4_SPECIAL_TOKENS
    ranks, merges, perm = synthetic_ranks(taylorswift[:20000], 60)
    tok = GPT4Tokenizer(mergeable_ranks=ranks)
    assert tok.special_tokens == GPT4_SPECIAL_TOKENS and tok.inverse_special_tokens[100257] == "<|endoftext|>"
    assert tok.byte_shuffle == {b: int(perm[b]) for b in range(256)}
    assert all(tok.inverse_byte_shuffle[tok.byte_shuffle[b]] == b for b in range(256))
    assert tok.vocab[300] == tok.vocab[merges[44][0]] + tok.vocab[merges[44][1]]
    for call in (lambda: tok.train("abc", 300), lambda: tok.save("x"), lambda: tok.load("x.model")):
        with pytest.raises(NotImplementedError):
            call()
    # decode of short id lists stays on the host: no GPU needed
    word = "hello"
    assert tok.decode([tok.byte_shuffle[b] for b in word.encode()]) == word
    with pytest.raises(KeyError):
        tok.decode([10 ** 6])
    tok.save_vocab(str(tmp_path / "g.vocab"))
    lines = open(tmp_path / "g.vocab", encoding="utf-8").read().splitlines()
    assert len(lines) == 256 + 60 and lines[-1].endswith(f"{256 + 59}") and " -> " in lines[-1]
    with pytest.raises(ValueError):
        GPT4Tokenizer(mergeable_ranks={b"a": 0})
    bad = dict(ranks)
    bad[b"\xff\xfe\xfd"] = 999999                                   # not the merge of two known tokens
    with pytest.raises(ValueError):
        GPT4Tokenizer(mergeable_ranks=bad)


def test_tiktoken_file_reader(tmp_path):
    import base64
    from minbpe_b200.gpt4 import load_tiktoken_file
    p = tmp_path / "t.tiktoken"
    p.write_bytes(b"".join(base64.b64encode(bytes((b,))) + b" " + str(b).encode() + b"\n" for b in range(256)) +
                  base64.b64encode(b"ab") + b" 256\n")
    r = load_tiktoken_file(str(p))
    assert len(r) == 257 and r[b"ab"] == 256 and r[b"\x00"] == 0
