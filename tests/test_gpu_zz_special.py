"""GPU: the special-token front end of RegexTokenizer.encode on the device (regex.py:123-164):
bpe_encode_text_gpt4_special finds the specials, splits every part between them with the GPT-4 pattern on its own and
encodes it — against the reference's own procedure (re.split + encode_ordinary per part) restated with the oracle."""
import random

import numpy as np
import pytest
import regex

import oracle

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def ref_encode(text, special, merges):
    """regex.py:152-163 with oracle.c_encode as encode_ordinary"""
    pat = "(" + "|".join(regex.escape(k) for k in special) + ")"
    out = []
    for part in regex.split(pat, text):
        if part in special:
            out.append(special[part])
        elif part:
            d, o = oracle.split_to_stream(part, GPT4)
            out.extend(oracle.c_encode(d, o, merges).tolist())
    return out


@pytest.fixture(scope="module")
def tok():
    from minbpe_b200 import RegexTokenizer
    from minbpe_b200.synth import generate
    text = generate(1337, 1 << 20).tobytes().decode("utf-8")[:300000]
    t = RegexTokenizer()
    t.train(text, 256 + 200)
    t.text = text
    t.m = np.array(list(t.merges.keys()), dtype=np.int32)
    return t


def test_encode_with_specials_through_the_class(tok):
    a, b, c = tok.text[:100000], tok.text[100000:200000], tok.text[200000:300000]
    sp = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|x|>": 100259, " <sp> ": 100260, "<|" + "z" * 44 + "|>": 100261}
    tok.register_special_tokens(sp)
    cases = [a + "<|endoftext|>" + b + "<|fim_prefix|><|x|>" + c,
             "<|endoftext|>" + a + "  <|endoftext|>\n\n" + b + " <sp> 'll" + c[:70000] + "<|x|>",
             a + " <|x|> 123<|x|>456 <|x|>'s " + b + "<|" + "z" * 44 + "|>" + c[:1000]]
    for t in cases:
        assert tok.encode(t, allowed_special="all") == ref_encode(t, sp, tok.m)
        sub = {"<|x|>"}
        assert tok.encode(t, allowed_special=sub) == ref_encode(t, {k: v for k, v in sp.items() if k in sub}, tok.m)
        assert tok.encode(t, allowed_special="none") == ref_encode(t, {"\x00never": 0}, tok.m)      # specials as ordinary text
    with pytest.raises(AssertionError):
        tok.encode(cases[0])                                                                         # none_raise (regex.py:139)
    # a special the device does not take (longer than 48 bytes) keeps the host split: same answer
    long_sp = dict(sp)
    long_sp["<|" + "y" * 60 + "|>"] = 100300
    tok.register_special_tokens(long_sp)
    t = cases[0] + "<|" + "y" * 60 + "|>" + a[:70000]
    assert tok.encode(t, allowed_special="all") == ref_encode(t, long_sp, tok.m)
    ids = tok.encode(cases[0], allowed_special="all")
    assert tok.decode(ids) == cases[0]


def test_random_adversarial_specials_through_the_abi(tok):
    from minbpe_b200.engine import Engine
    eng = Engine(0)
    rnd = random.Random(99)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '", "the ", "ing", "\U0001f600"]
    special_sets = [["<|endoftext|>"], ["<|a|>", "<|ab|>", "|>"], [" <s> ", "'s", "12"], ["x", "日"], ["\n<eot>\n", "  "], ["a b", "b a"]]
    for it in range(300):
        names = rnd.choice(special_sets)
        special = {k: 100000 + i for i, k in enumerate(names)}
        parts = []
        for _ in range(rnd.randint(1, 30)):
            parts.append("".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 40))))
            if rnd.random() < 0.7:
                parts.append(rnd.choice(names))
        text = "".join(parts)
        if not text:
            continue
        spec = [(k.encode("utf-8"), v) for k, v in special.items()]
        got = eng.encode_text_gpt4(text.encode("utf-8"), tok.m, specials=spec).tolist()
        assert got == ref_encode(text, special, tok.m), (it, text, names)
    eng.close()


def test_special_sets_do_not_leak_through_the_memo(tok):
    """The encode memo holds the specials as single-id entries: another set (or none) must not see them."""
    from minbpe_b200.engine import Engine
    eng = Engine(0)
    text = tok.text[:80000] + "<|x|>" + tok.text[80000:160000] + "<|y|>" + tok.text[160000:200000]
    raw = text.encode("utf-8")
    s1, s2 = {"<|x|>": 7001}, {"<|y|>": 7002, "<|x|>": 7003}
    for special in (s1, s2, s1, None, s2):
        spec = None if special is None else [(k.encode("utf-8"), v) for k, v in special.items()]
        got = eng.encode_text_gpt4(raw, tok.m, specials=spec).tolist()
        assert got == ref_encode(text, special or {"\x00never": 0}, tok.m)
    eng.close()


def test_pieces_and_oversize_chunks_with_specials(tok):
    from minbpe_b200 import engine as E
    eng = E.Engine(0)
    sp = {"<|im start|>": 9001, "d e": 9002}          # specials that contain letter+space: a piece cut must not split them
    body = tok.text[:150000]
    text = "".join(body[i:i + 5000] + ("<|im start|>" if (i // 5000) % 2 else "d e") for i in range(0, 150000, 5000))
    spec = [(k.encode("utf-8"), v) for k, v in sp.items()]
    want = ref_encode(text, sp, tok.m)
    eng.set_option(E.OPT_SPLIT_PIECE, 1 << 14)
    try:
        assert eng.encode_text_gpt4(text.encode("utf-8"), tok.m, specials=spec).tolist() == want
        assert eng.encode_stats()["pieces"] > 5
    finally:
        eng.set_option(E.OPT_SPLIT_PIECE, 0)
    # a chunk of 12,000 letters does not fit the CTA path: the piece takes the general path, part by part
    t2 = body[:50000] + "<|im start|>" + "lyiltumdya" * 1200 + " d e" + body[50000:90000] + "<|im start|>"
    assert eng.encode_text_gpt4(t2.encode("utf-8"), tok.m, specials=spec).tolist() == ref_encode(t2, sp, tok.m)
    assert eng.encode_stats()["fallback_pieces"] == 1
    assert eng.encode_text_gpt4(text.encode("utf-8"), tok.m, specials=spec).tolist() == want     # and the memo works again
    # limits of the device front end are errors of the C ABI (the class keeps the host split for such sets)
    with pytest.raises(E.EngineError):
        eng.encode_text_gpt4(b"abc", tok.m, specials=[(b"x" * 49, 1)])
    with pytest.raises(E.EngineError):
        eng.encode_text_gpt4(b"abc", tok.m, specials=[(b"", 1)])
    eng.close()
