"""GPU: bpe_decode (k_decode.cuh) against the reference's decode (basic.py:51-55, regex.py:78-90)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_decode(vocab, ids):
    return b"".join(vocab[i] for i in ids)


def test_decode_matches_join(golden_train, taylorswift):
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    for kind, cls in (("basic", BasicTokenizer), ("regex", RegexTokenizer)):
        tok = cls()
        tok.train(taylorswift, 512)
        ids = tok.encode(taylorswift)
        assert len(ids) >= tok.DEVICE_DECODE_MIN_IDS
        assert tok.decode(ids) == taylorswift                       # device path
        tok.DEVICE_DECODE_MIN_IDS = 1 << 60
        assert tok.decode(ids) == taylorswift                       # host path, same text


def test_decode_random_ids_and_errors():
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    from minbpe_b200.engine import Engine
    rng = np.random.default_rng(7)
    eng = Engine(0)
    vocab = {i: bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8)) for i in range(3000)}
    top = 3000
    lens = np.array([len(vocab[i]) for i in range(top)], dtype=np.uint32)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    blob = np.frombuffer(b"".join(vocab[i] for i in range(top)), dtype=np.uint8)
    for n in (0, 1, 7, 2047, 2048, 2049, 300001):
        ids = rng.integers(0, top, size=n).astype(np.int32)
        data, bad = eng.decode(ids, blob, starts, lens)
        assert bad == -1 and data == ref_decode(vocab, ids.tolist())
    ids = rng.integers(0, top, size=100000).astype(np.int32)
    ids[77777] = top + 5
    ids[88888] = -1
    data, bad = eng.decode(ids, blob, starts, lens)
    assert data is None and bad == 77777
    lens2 = lens.copy(); lens2[123] = 0xFFFFFFFF
    ids = np.full(70000, 5, dtype=np.int32); ids[69999] = 123
    assert eng.decode(ids, blob, starts, lens2) == (None, 69999)
    eng.close()
    # reference error types through the classes
    b = BasicTokenizer(); r = RegexTokenizer()
    b.engine, r.engine   # decode goes to the device only for tokenizers of a process that already holds an engine
    long_ids = [65] * (b.DEVICE_DECODE_MIN_IDS + 5)
    assert b.decode(long_ids) == "A" * len(long_ids) and r.decode(long_ids) == "A" * len(long_ids)
    with pytest.raises(KeyError):
        b.decode(long_ids + [100000])
    with pytest.raises(ValueError):
        r.decode(long_ids + [100000])
    r.register_special_tokens({"<|endoftext|>": 100257})
    assert r.decode(long_ids + [100257]) == "A" * len(long_ids) + "<|endoftext|>"
