"""GPU: the GPT-2 split pattern (regex.py:18) on the device — BPE_OPT_SPLIT_PATTERN = 1 selects spl_chunk_start_gpt2 in the
split kernels — through the C ABI and through RegexTokenizer(GPT2_SPLIT_PATTERN), against `regex` and the oracle."""
import numpy as np
import pytest
import regex

import oracle

pytestmark = pytest.mark.gpu

GPT2_PAT = r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""
GPT2 = regex.compile(GPT2_PAT)
GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def synth(n):
    from minbpe_b200.synth import generate
    return generate(1337, 1 << 20).tobytes().decode("utf-8")[:n]


def test_split_offsets_gpt2_vs_regex(taylorswift):
    from minbpe_b200 import engine as E
    eng = E.Engine(0)
    eng.set_option(E.OPT_SPLIT_PATTERN, 1)
    for text in (taylorswift, synth(400000), "it's we'll I'VE  x\n\n  y 123 4567 !!'s 's\t\n", "a", " ", "日本語 テキスト 12"):
        data, offs = oracle.split_to_stream(text, GPT2)
        assert np.array_equal(eng.split_gpt4(data.tobytes()), offs)
    eng.set_option(E.OPT_SPLIT_PIECE, 1 << 15)           # several pieces, cut at letter+space
    try:
        data, offs = oracle.split_to_stream(taylorswift, GPT2)
        assert np.array_equal(eng.split_gpt4(data.tobytes()), offs)
    finally:
        eng.set_option(E.OPT_SPLIT_PIECE, 0)
    eng.set_option(E.OPT_SPLIT_PATTERN, 0)
    data, offs = oracle.split_to_stream(taylorswift, GPT4)
    assert np.array_equal(eng.split_gpt4(data.tobytes()), offs)
    with pytest.raises(E.EngineError):
        eng.set_option(E.OPT_SPLIT_PATTERN, 2)
    eng.close()


def test_regex_tokenizer_with_the_gpt2_pattern(taylorswift):
    from minbpe_b200 import RegexTokenizer
    text = taylorswift                                   # 185 KB: above the device-split threshold
    tok = RegexTokenizer(GPT2_PAT)
    tok.train(text, 256 + 80)
    data, offs = oracle.split_to_stream(text, GPT2)
    wp, wc, wn = oracle.c_train(data.astype(np.int32), offs, 80)
    assert wn == 80 and [list(p) for p in tok.merges] == wp.tolist()
    assert tok.encode_ordinary(text) == oracle.c_encode(data, offs, wp).tolist()
    # specials under the GPT-2 pattern (regex.py:152-163)
    tok.register_special_tokens({"<|endoftext|>": 50256})
    t2 = text[:90000] + "<|endoftext|>" + text[90000:]
    want = []
    for part in regex.split("(" + regex.escape("<|endoftext|>") + ")", t2):
        if part == "<|endoftext|>":
            want.append(50256)
        elif part:
            d, o = oracle.split_to_stream(part, GPT2)
            want.extend(oracle.c_encode(d, o, wp).tolist())
    assert tok.encode(t2, allowed_special="all") == want
    # a GPT-4 tokenizer on the same (shared) engine afterwards still splits with its own pattern
    t4 = RegexTokenizer()
    t4.train(text, 256 + 40)
    d4, o4 = oracle.split_to_stream(text, GPT4)
    w4, _, _ = oracle.c_train(d4.astype(np.int32), o4, 40)
    assert [list(p) for p in t4.merges] == w4.tolist()
    assert tok.encode_ordinary(text) == oracle.c_encode(data, offs, wp).tolist()
