"""GPU: the file / shard entry points beside the reference API — RegexTokenizer.train_from_file (regex.py:36-70 for a
text file of any size, split on the device in pieces) and dist.encode_sharded (regex.py:111-121 over byte-range shards)."""
import os

import numpy as np
import pytest
import regex

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def test_train_from_file_equals_train(golden_train):
    """One process: the whole file goes to this GPU.  Same merges as the reference's RegexTokenizer.train on the text."""
    from minbpe_b200 import RegexTokenizer
    from minbpe_b200 import engine as E
    g = golden_train["taylorswift_regex_512"]
    path = os.path.join(GOLDEN, "taylorswift.txt")
    tok = RegexTokenizer()
    tok.train_from_file(path, 256 + 64)
    assert [list(p) for p in tok.merges] == g["merges"][:64]
    # the same with the text split on the device in small pieces (cut at letter+space)
    tok2 = RegexTokenizer()
    tok2.engine.set_option(E.OPT_SPLIT_PIECE, 1 << 15)
    try:
        tok2.train_from_file(path, 256 + 64)
    finally:
        tok2.engine.set_option(E.OPT_SPLIT_PIECE, 0)
    assert tok2.merges == tok.merges and tok2.vocab == tok.vocab
    with pytest.raises(ValueError):
        RegexTokenizer(r"\w+|\s+").train_from_file(path, 300)


def test_encode_sharded_concatenates_to_the_whole(taylorswift):
    from minbpe_b200.dist import encode_sharded
    from minbpe_b200.engine import Engine
    data, offs = oracle.split_to_stream(taylorswift, GPT4)
    eng = Engine(0)
    eng.load_stream(data, offs)
    merges, _, done = eng.train(120)
    want = oracle.c_encode(data, offs, merges)
    for world in (1, 2, 3):
        parts = [encode_sharded(eng, data, offs, merges, rank=r, world=world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), want), world
    eng.close()


def test_encode_file_shards_and_specials(tmp_path, taylorswift):
    """RegexTokenizer.encode_file / dist.encode_file: byte-range shards cut at letter+space and never inside a special."""
    from minbpe_b200 import RegexTokenizer
    from minbpe_b200.dist import encode_file
    tok = RegexTokenizer()
    tok.train(taylorswift, 256 + 100)
    tok.register_special_tokens({"<|endoftext|>": 100257, "<|im start|>": 100264})
    # specials every ~3 KB, one kind containing letter+space, so that a naive cut could land inside one
    body = taylorswift[:150000]
    text = "".join(body[i:i + 3000] + ("<|im start|>" if (i // 3000) % 2 else "<|endoftext|>") for i in range(0, len(body), 3000))
    p = tmp_path / "doc.txt"
    p.write_bytes(text.encode("utf-8"))
    want = tok.encode(text, allowed_special="all")
    assert tok.encode_file(str(p), "all").tolist() == want
    spec = [(k.encode("utf-8"), v) for k, v in tok.special_tokens.items()]
    m = tok._merge_array()
    for world in (2, 3, 5):
        parts = [encode_file(tok.engine, str(p), m, None, spec, rank=r, world=world) for r in range(world)]
        assert np.concatenate(parts).tolist() == want, world
    assert tok.encode_file(str(p), "none").tolist() == tok.encode(text, allowed_special="none")
    with pytest.raises(ValueError):
        tok.encode_file(str(p), "none_raise")
