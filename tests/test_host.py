"""CPU: host-side logic that mirrors the reference API without touching the device —
save/load file format (base.py:97-165) byte-exact against files written by the reference,
vocab construction, decode, special-token splitting, regex split offsets, error conventions."""
import hashlib
import os

import numpy as np
import pytest
import regex

import oracle
from conftest import GOLDEN
from minbpe_b200 import GPT2_SPLIT_PATTERN, GPT4_SPLIT_PATTERN, BasicTokenizer, RegexTokenizer, Tokenizer
from minbpe_b200.tokenizer import render_token, split_text

test_strings = ["", "?", "hello world!!!? (안녕하세요!) lol123 😉"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("kind", ["basic", "regex"])
def test_save_matches_reference_files(golden_train, tmp_path, kind):
    """Install the reference's merges and save(): .model and .vocab must be byte-identical to the
    files the reference wrote (SURVEY.md §8c fingerprints)."""
    g = golden_train[f"taylorswift_{kind}_512"]
    tok = BasicTokenizer() if kind == "basic" else RegexTokenizer()
    tok.merges = {tuple(p): 256 + i for i, p in enumerate(g["merges"])}
    tok.vocab = tok._build_vocab()
    prefix = str(tmp_path / "m")
    tok.save(prefix)
    assert sha(prefix + ".model") == g["model_sha256"] == sha(os.path.join(GOLDEN, f"ref_{kind}512.model"))
    assert sha(prefix + ".vocab") == g["vocab_sha256"] == sha(os.path.join(GOLDEN, f"ref_{kind}512.vocab"))


@pytest.mark.parametrize("kind", ["basic", "regex"])
def test_load_reference_model(golden_train, taylorswift, kind):
    g = golden_train[f"taylorswift_{kind}_512"]
    tok = BasicTokenizer() if kind == "basic" else RegexTokenizer()
    tok.load(os.path.join(GOLDEN, f"ref_{kind}512.model"))
    assert [list(p) for p in tok.merges] == g["merges"]
    assert list(tok.merges.values()) == list(range(256, 512))
    assert tok.pattern == ("" if kind == "basic" else GPT4_SPLIT_PATTERN)
    # decode the reference's ids back to the text with the loaded vocab
    ids = oracle.c_encode(*oracle.split_to_stream(taylorswift, None if kind == "basic" else regex.compile(GPT4_SPLIT_PATTERN)),
                          np.array(g["merges"], dtype=np.int32))
    assert tok.decode(ids.tolist()) == taylorswift


def test_save_load_with_specials(golden_train, tmp_path):
    g = golden_train["llama_regex_320_specials"]
    tok = RegexTokenizer()
    tok.merges = {tuple(p): 256 + i for i, p in enumerate(g["merges"])}
    tok.vocab = tok._build_vocab()  # train() builds the vocab; registering specials does not touch it
    tok.register_special_tokens(g["specials"])
    prefix = str(tmp_path / "t")
    tok.save(prefix)
    assert open(prefix + ".model", encoding="utf-8").read() == g["model_text"]
    assert open(prefix + ".vocab", encoding="utf-8").read() == g["vocab_text"]
    t2 = RegexTokenizer()
    t2.load(prefix + ".model")
    assert t2.merges == tok.merges and t2.special_tokens == g["specials"]
    assert {k: v for k, v in t2.vocab.items() if k < 100000} == tok.vocab  # load() adds the specials (base.py:93-94)
    assert t2.vocab[100257] == b"<|endoftext|>"
    llama = open(os.path.join(GOLDEN, "llama_text.txt"), encoding="utf-8").read()
    assert tok.decode(g["ids_all"]) == llama


def test_load_asserts():
    with pytest.raises(AssertionError):
        Tokenizer().load("foo.txt")


@pytest.mark.parametrize("factory", [BasicTokenizer, RegexTokenizer])
@pytest.mark.parametrize("text", test_strings)
def test_untrained_identity(factory, text):
    # reference tests/test_tokenizer.py:52-59 (untrained tokenizers emit raw bytes)
    tok = factory()
    ids = tok.encode(text)
    assert ids == list(text.encode("utf-8"))
    assert tok.decode(ids) == text


def test_untrained_specials_split():
    tok = RegexTokenizer()
    sp = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258}
    tok.register_special_tokens(sp)
    s = open(os.path.join(GOLDEN, "specials_string.txt"), encoding="utf-8").read()
    ids = tok.encode(s, allowed_special="all")
    assert ids.count(100257) == s.count("<|endoftext|>") and ids.count(100258) == 1
    assert tok.decode(ids) == s
    assert tok.encode(s, allowed_special="none") == list(s.encode())
    assert 100258 not in tok.encode(s, allowed_special={"<|endoftext|>"})
    with pytest.raises(AssertionError):
        tok.encode(s)  # none_raise, regex.py:139
    with pytest.raises(ValueError):
        tok.encode(s, allowed_special="bogus")


def test_decode_errors():
    with pytest.raises(ValueError, match="invalid token id"):
        RegexTokenizer().decode([99999])  # regex.py:87
    with pytest.raises(KeyError):
        BasicTokenizer().decode([99999])  # basic.py:53
    assert BasicTokenizer().decode([0xE2, 0x82]) == "�"  # errors="replace"


def test_vocab_size_assert():
    for f in (BasicTokenizer, RegexTokenizer):
        with pytest.raises(AssertionError):
            f().train("abc", 255)


def test_surrogate_raises():
    with pytest.raises(UnicodeEncodeError):
        BasicTokenizer().encode("\ud800")


def test_render_token():
    assert render_token(b"\n") == "\\u000a"
    assert render_token(b"\xff") == "�"
    assert render_token(" hello".encode()) == " hello"


@pytest.mark.parametrize("pattern", [GPT4_SPLIT_PATTERN, GPT2_SPLIT_PATTERN, r"\w+", r"\s*"])
def test_split_text_matches_findall(taylorswift, pattern):
    cp = regex.compile(pattern)
    texts = test_strings + [taylorswift[:20000], "a\r\n\r\n  b\t\n 123456 x's Y'LL  éè \U0001f600\U0001f600!!!\n\n"]
    for t in texts:
        data, offs = split_text(cp, t)
        want_b, want_o = oracle.split_to_stream(t, cp)
        chunks = [c.encode() for c in cp.findall(t) if c]
        assert bytes(data) == b"".join(chunks)
        got = [bytes(data)[int(a):int(b)] for a, b in zip(offs, list(offs[1:]) + [len(data)])] if len(offs) else []
        assert got == chunks


def test_synthetic_shards_are_ranges_of_one_corpus():
    """bench.py --gpus N gives rank r the r-th contiguous range of ONE corpus (minbpe_b200/synth.py first_block):
    the shards concatenate to the text a single generation produces, and each is valid UTF-8 on its own."""
    from minbpe_b200.synth import generate
    full = generate(1337, 6 << 20, threads=3)
    parts = [generate(1337, 2 << 20, threads=2, first_block=2 * r) for r in range(3)]
    assert np.array_equal(np.concatenate(parts), full)
    for p in parts:
        p.tobytes().decode("utf-8")
    assert not np.array_equal(parts[0], parts[1])


def test_encode_file_cuts_never_fall_inside_a_special_token(tmp_path):
    """dist.encode_file: a rank's byte range ends at a letter+space point that is not inside an occurrence of a special
    token — also when the occurrence straddles the nominal cut (size * r // world), i.e. starts before the search window."""
    import numpy as np
    from minbpe_b200.dist import encode_file

    class Ranges:   # stands in for the engine: "ids" = the bytes it was handed
        def encode_text_gpt4(self, raw, merges, perm, specials=None):
            return np.asarray(raw, dtype=np.uint8).astype(np.int32)

    sp = b"<|end a of b text c|>"          # letter+space points inside the token
    for shift in range(len(sp) + 2):
        tail = b" tail word" * 120
        head = b"xy" * ((len(tail) + len(sp)) // 2 + 8)
        text = head[: len(tail) + shift] + sp + tail          # nominal cut of world 2 walks through the token as shift grows
        p = tmp_path / f"t{shift}.txt"
        p.write_bytes(text)
        for world in (2, 3):
            parts = [encode_file(Ranges(), str(p), None, None, [(sp, 1000)], rank=r, world=world) for r in range(world)]
            assert np.concatenate(parts).astype(np.uint8).tobytes() == text
            at = text.find(sp)
            for c in np.cumsum([len(x) for x in parts])[:-1]:
                assert not (at < c < at + len(sp)), (shift, world, int(c), at)
                assert text[c - 1: c].isalpha() and text[c: c + 1] == b" "
