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


def test_split_gpt2_random_adversarial():
    """Random strings over an alphabet of the characters the two patterns treat differently (apostrophe forms in both cases,
    digit runs, Unicode spaces and line separators, case-folding specials) — device split vs regex.findall, GPT-2 pattern."""
    import random
    from minbpe_b200 import engine as E
    rnd = random.Random(2424)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,'") + [chr(c) for c in (0x3000, 0xe9, 0x65e5, 0x17f, 0x212a, 0xbd, 0x2028, 0x85, 0xa0, 0x1f600)]
    alphabet += [" ", " ", "  ", "\n\n", "'ll", "'ve", "'LL", " '", "123456", "x" * 7]
    eng = E.Engine(0)
    eng.set_option(E.OPT_SPLIT_PATTERN, 1)
    try:
        for size in (1, 2, 3, 17, 2047, 2048, 2049, 4097, 60000, 300000):
            text = "".join(rnd.choice(alphabet) for _ in range(size))
            data, offs = oracle.split_to_stream(text, GPT2)
            assert np.array_equal(eng.split_gpt4(data.tobytes()), offs), (size, text[:80])
        text = " " * 5000 + "a" * 7000 + "1" * 9001 + "!" * 4099 + "\n" * 3000 + " \n" * 2500 + "x" + "'s" * 3000
        data, offs = oracle.split_to_stream(text, GPT2)
        assert np.array_equal(eng.split_gpt4(data.tobytes()), offs)
    finally:
        eng.set_option(E.OPT_SPLIT_PATTERN, 0)
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
