"""GPU: the round-2 paths against vectors produced by the UNMODIFIED reference (tests/golden/make_golden_r2.py ->
golden_r2.json): the GPT-2 split pattern, RegexTokenizer.encode with special tokens under both patterns
(regex.py:123-164), and GPT4Tokenizer on a synthetic tiktoken-style rank table (gpt4.py)."""
import base64
import hashlib

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def ids_sha(ids):
    return hashlib.sha256(np.asarray(ids, dtype="<i4").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g():
    return load_golden("golden_r2.json")


def bodies(text):
    a, b, c = text[:70000], text[70000:140000], text[140000:]
    return [
        a + "<|endoftext|>" + b + "<|fim_prefix|><|x|>" + c,
        "<|endoftext|>" + a + "  <|endoftext|>\n\n" + b + " <sp> 'll" + c[:30000] + "<|x|>",
        a[:50000] + "<|ab|>x<|a|>|>" + b[:30000] + " 12's <s> 123" + c[:20000] + "<|im start|> d e",
    ]


def test_gpt2_pattern_train_and_encode_equal_the_reference(g, taylorswift):
    from minbpe_b200 import GPT2_SPLIT_PATTERN, RegexTokenizer
    want = g["gpt2_train"]
    tok = RegexTokenizer(GPT2_SPLIT_PATTERN)
    tok.train(taylorswift, want["vocab_size"])
    assert [list(p) for p in tok.merges] == want["merges"]
    ids = tok.encode_ordinary(taylorswift)
    assert len(ids) == want["n_ids"] and ids[:64] == want["ids_head"] and ids_sha(ids) == want["ids_sha256"]


def test_encode_with_specials_equals_the_reference(g, taylorswift):
    from minbpe_b200 import GPT2_SPLIT_PATTERN, GPT4_SPLIT_PATTERN, RegexTokenizer
    sp = g["specials"]
    toks = {}
    for name, pat in (("gpt4", GPT4_SPLIT_PATTERN), ("gpt2", GPT2_SPLIT_PATTERN)):
        t = RegexTokenizer(pat)
        t.merges = {(a, b): 256 + i for i, (a, b) in enumerate(sp["merges_" + name])}
        t.vocab = t._build_vocab()
        toks[name] = t
    texts = bodies(taylorswift)
    for c in sp["cases"]:
        tok = toks[c["pattern"]]
        tok.register_special_tokens(c["special"])
        got = tok.encode(texts[c["body"]], allowed_special="all")
        assert len(got) == c["n_ids"] and got[:40] == c["ids_head"] and got[-40:] == c["ids_tail"], (c["pattern"], c["special"], c["body"])
        assert ids_sha(got) == c["ids_sha256"]


def test_gpt4_tokenizer_equals_the_reference_class_on_a_synthetic_table(g, taylorswift):
    from minbpe_b200 import GPT4Tokenizer
    want = g["gpt4_synthetic"]
    ranks = {base64.b64decode(k): v for k, v in want["ranks_b64"]}
    tok = GPT4Tokenizer(mergeable_ranks=ranks)
    assert [list(p) for p in tok.merges] == want["merges"]
    ids = tok.encode_ordinary(taylorswift[:want["ordinary"]["text_chars"]])
    assert ids[:48] == want["ordinary"]["ids_head"] and ids_sha(ids) == want["ordinary"]["ids_sha256"]
    assert tok.decode(ids) == taylorswift[:40000] and want["roundtrip_ok"]
    s = "<|endoftext|>" + taylorswift[:20000] + "<|fim_prefix|>x<|endofprompt|>" + taylorswift[20000:40000]
    got = tok.encode(s, allowed_special="all")
    assert len(got) == want["special_all"]["n_ids"] and ids_sha(got) == want["special_all"]["ids_sha256"]
