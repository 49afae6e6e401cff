"""GPU: GPT4Tokenizer (gpt4.py:57-130) on the device encode path — byte shuffle + recovered merges — with a synthetic
tiktoken-style rank table (cl100k_base itself needs a download; when tiktoken can load it, ids are compared with
tiktoken's as the reference's tests/test_tokenizer.py:61-77 does)."""
import numpy as np
import pytest
import regex

import oracle
from test_gpt4_host import synthetic_ranks

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def expect(text, merges, perm, special=None):
    out = []
    pat = "(" + "|".join(regex.escape(k) for k in special) + ")" if special else None
    for part in (regex.split(pat, text) if pat else [text]):
        if special and part in special:
            out.append(special[part])
        elif part:
            d, o = oracle.split_to_stream(part, GPT4)
            out.extend(oracle.c_encode(d, o, merges, perm).tolist())
    return out


def test_gpt4_tokenizer_on_a_synthetic_rank_table(taylorswift):
    from minbpe_b200 import GPT4Tokenizer
    ranks, merges, perm = synthetic_ranks(taylorswift[:30000], 120)
    if len(ranks) != 256 + len(merges):
        pytest.skip("degenerate synthetic table")
    tok = GPT4Tokenizer(mergeable_ranks=ranks)
    m = np.asarray(merges, dtype=np.int32)
    p8 = perm.astype(np.uint8)
    for text in ("hello world", "a", "", "hello world!!!? (안녕하세요!) lol123 😉", taylorswift[:5000], taylorswift):
        ids = tok.encode_ordinary(text)                      # short: host split + bpe_encode; long: fused on the device
        assert ids == expect(text, m, p8), len(text)
        assert tok.decode(ids) == text
        assert tok.encode(text, allowed_special="all") == ids
    sp_text = "<|endoftext|>" + taylorswift[:80000] + "<|fim_prefix|>x<|endofprompt|>" + taylorswift[80000:] + "<|endoftext|>"
    want = expect(sp_text, m, p8, tok.special_tokens)
    assert tok.encode(sp_text, allowed_special="all") == want
    short = "<|endoftext|>Hello world this is one document<|fim_suffix|> tail"
    assert tok.encode(short, allowed_special="all") == expect(short, m, p8, tok.special_tokens)
    with pytest.raises(AssertionError):
        tok.encode(short)                                    # none_raise
    assert tok.encode(short, allowed_special="none") == expect(short, m, p8)
    # per-chunk entry point (regex.py:92-109 through gpt4.py:81-86): single bytes are permuted too
    assert tok._encode_chunk(b"a") == [int(perm[ord("a")])]
    assert tok._encode_chunk(b"hello") == oracle.c_encode(np.frombuffer(b"hello", dtype=np.uint8), None, m, p8).tolist()


def test_gpt4_tiktoken_equality_when_the_vocabulary_is_available():
    """tests/test_tokenizer.py:61-77 of the reference: needs tiktoken's cl100k_base (a download)."""
    import os
    if not os.environ.get("BPE_TEST_TIKTOKEN"):
        pytest.skip("set BPE_TEST_TIKTOKEN=1 on a machine where tiktoken can load cl100k_base (network or cache)")
    import tiktoken
    enc = tiktoken.get_encoding("cl100k_base")
    from minbpe_b200 import GPT4Tokenizer
    tok = GPT4Tokenizer()
    for text in ("", "?", "hello world!!!? (안녕하세요!) lol123 😉"):
        assert tok.encode(text) == enc.encode(text)
        assert tok.decode(enc.encode(text)) == text
    s = "<|endoftext|>Hello world this is one document\n<|endoftext|>And this is another document\n<|fim_prefix|>x<|fim_suffix|>"
    assert tok.encode(s, allowed_special="all") == enc.encode(s, allowed_special="all")
