"""GPU: the device-side GPT-4 splitter (bpe_split_gpt4 / bpe_load_text_gpt4, k_split.cuh) against
`regex.findall` — the reference's own pre-split (regex.py:41,114) — on the reference corpus, the
synthetic corpus, edge cases and a large random adversarial text."""
import random

import numpy as np
import pytest
import regex

import oracle

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


@pytest.fixture(scope="module")
def eng():
    from minbpe_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def want_offsets(text):
    data, offs = oracle.split_to_stream(text, GPT4)
    return data, offs


def check(eng, text):
    data, offs = want_offsets(text)
    got = eng.split_gpt4(data.tobytes())
    assert np.array_equal(got, offs), (len(text), got[:10], offs[:10], text[:60])


def test_edge_cases(eng):
    from test_split_rules import CASES
    for t in CASES:
        if t:
            check(eng, t)
    assert eng.split_gpt4(b"").size == 0


def test_corpora(eng, taylorswift):
    check(eng, taylorswift)
    from minbpe_b200.synth import generate
    check(eng, generate(1337, 8 << 20).tobytes().decode("utf-8"))


def test_random_adversarial(eng):
    rnd = random.Random(4242)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,'") + [chr(c) for c in (0x3000, 0xe9, 0x65e5, 0x17f, 0x212a, 0xbd, 0x2028, 0x85, 0xa0, 0x1f600)]
    alphabet += [" ", " ", "  ", "\n\n", "'ll", "'ve", " '", "123456", "x" * 7]
    for size in (1, 2, 3, 17, 2047, 2048, 2049, 4097, 100000, 1500000):
        check(eng, "".join(rnd.choice(alphabet) for _ in range(size)))
    # long runs crossing many 2048-byte scan tiles
    check(eng, " " * 5000 + "a" * 7000 + "1" * 9001 + "!" * 4099 + "\n" * 3000 + " \n" * 2500 + "x")
    check(eng, "é" * 5000 + " 日本" * 3000 + "٣" * 4001)


def test_load_text_equals_load_stream(eng, taylorswift):
    """Training from bpe_load_text_gpt4 == training from host offsets (and both == golden in test_gpu_parity)."""
    data, offs = want_offsets(taylorswift)
    eng.load_stream(data, offs)
    p1, c1, d1 = eng.train(100)
    n_chunks = eng.load_text_gpt4(data.tobytes(), count_chunks=True)
    assert n_chunks == len(offs)
    p2, c2, d2 = eng.train(100)
    assert d1 == d2 == 100 and np.array_equal(p1, p2) and np.array_equal(c1, c2)


def test_piecewise_split_equals_whole(eng, taylorswift):
    """Texts beyond one split call (1 GiB pieces by default) are cut where an ASCII letter is followed by U+0020
    (a provable chunk boundary of the GPT-4 pattern).  With the piece size forced down to 64 KiB the
    same code path runs on small texts: offsets, chunk counts and the marked training stream must not change."""
    from minbpe_b200 import engine as E
    from minbpe_b200.synth import generate
    texts = [taylorswift, generate(1337, 8 << 20).tobytes().decode("utf-8")]
    try:
        for text in texts:
            data, offs = want_offsets(text)
            raw = data.tobytes()
            for piece in (1 << 16, (1 << 20) + 4096):
                eng.set_option(E.OPT_SPLIT_PIECE, piece)
                assert np.array_equal(eng.split_gpt4(raw), offs)
                assert eng.load_text_gpt4(raw, count_chunks=True) == len(offs)
                p1, c1, d1 = eng.train(24)
                eng.load_text_gpt4(raw)          # the marks-straight-from-the-rule-kernel path
                p2, c2, d2 = eng.train(24)
                eng.set_option(E.OPT_SPLIT_PIECE, 0)
                eng.load_stream(data, offs)
                p0, c0, d0 = eng.train(24)
                assert d0 == d1 == d2 == 24
                assert np.array_equal(p0, p1) and np.array_equal(c0, c1) and np.array_equal(p0, p2) and np.array_equal(c0, c2)
        # a text without any letter+space inside a piece cannot be cut: a clean error, not a wrong split
        eng.set_option(E.OPT_SPLIT_PIECE, 4096)
        with pytest.raises(E.EngineError):
            eng.split_gpt4(("12345 " * 4000).encode())
    finally:
        eng.set_option(E.OPT_SPLIT_PIECE, 0)
