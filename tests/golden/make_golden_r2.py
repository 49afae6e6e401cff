#!/usr/bin/env python3
"""
Round-2 golden vectors from the UNMODIFIED reference (karpathy/minbpe at /root/reference), for the paths added in round 2:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py        (build container only)

golden_r2.json:
  gpt2_train      RegexTokenizer(GPT2_SPLIT_PATTERN).train(taylorswift, 256 + 64): merges, ids of the text (sha256 + head)
  specials        RegexTokenizer.encode(text, allowed_special=...) for the GPT-4 and the GPT-2 pattern, several special sets
                  (regex.py:123-164): the ids the reference returns, with the text and the set so that the device front end
                  (bpe_encode_text_gpt4_special) can be run on the same input
  gpt4_synthetic  GPT4Tokenizer built on a synthetic tiktoken-style rank table (the reference class with its tiktoken loading
                  replaced by that table): recovered merges, ids, decode — cl100k_base itself is a download
Nothing here is imported by the product; tests read the JSON.
"""
import hashlib
import json
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from minbpe import RegexTokenizer  # noqa: E402  (the reference)
from minbpe.regex import GPT2_SPLIT_PATTERN, GPT4_SPLIT_PATTERN  # noqa: E402


def ids_sha(ids):
    import numpy as np
    return hashlib.sha256(np.asarray(ids, dtype="<i4").tobytes()).hexdigest()


def main():
    text = open(os.path.join(HERE, "taylorswift.txt"), encoding="utf-8").read()
    out = {}
    # ---- GPT-2 pattern
    t2 = RegexTokenizer(GPT2_SPLIT_PATTERN)
    t2.train(text, 256 + 64)
    ids = t2.encode_ordinary(text)
    out["gpt2_train"] = {"vocab_size": 320, "merges": [list(p) for p in t2.merges], "ids_sha256": ids_sha(ids), "n_ids": len(ids),
                         "ids_head": ids[:64]}
    # ---- specials under both patterns (tokenizers trained on the same text, 64 merges)
    t4 = RegexTokenizer(GPT4_SPLIT_PATTERN)
    t4.train(text, 256 + 64)
    a, b, c = text[:70000], text[70000:140000], text[140000:]
    cases = []
    sets = [
        {"<|endoftext|>": 100257},
        {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|x|>": 100259, " <sp> ": 100260},
        {"<|a|>": 7, "<|ab|>": 8, "|>": 9},
        {"'s": 900001, "12": 900002, " <s> ": 900003},
        {"<|im start|>": 9001, "d e": 9002},
    ]
    bodies = [
        a + "<|endoftext|>" + b + "<|fim_prefix|><|x|>" + c,
        "<|endoftext|>" + a + "  <|endoftext|>\n\n" + b + " <sp> 'll" + c[:30000] + "<|x|>",
        a[:50000] + "<|ab|>x<|a|>|>" + b[:30000] + " 12's <s> 123" + c[:20000] + "<|im start|> d e",
    ]
    for pattern_name, tok in (("gpt4", t4), ("gpt2", t2)):
        for si, special in enumerate(sets):
            tok.register_special_tokens(special)
            for bi, body in enumerate(bodies):
                got = tok.encode(body, allowed_special="all")
                cases.append({"pattern": pattern_name, "special": special, "body": bi, "ids_sha256": ids_sha(got), "n_ids": len(got),
                              "ids_head": got[:40], "ids_tail": got[-40:]})
    out["specials"] = {"merges_gpt4": [list(p) for p in t4.merges], "merges_gpt2": [list(p) for p in t2.merges],
                       "bodies": [[0, 70000, 140000], "see tests/test_gpu_zz_golden_r2.py: bodies are rebuilt from taylorswift.txt"],
                       "cases": cases}
    # ---- GPT4Tokenizer on a synthetic rank table: the reference class, tiktoken replaced by a stub that returns the table
    import numpy as np
    perm = np.random.default_rng(3).permutation(256)
    base = RegexTokenizer(GPT4_SPLIT_PATTERN)
    shuffled = bytes(int(perm[x]) for x in text[:30000].encode("utf-8"))
    # train on the permuted bytes as ONE chunk stream per regex chunk: use the reference's own helpers
    from minbpe.base import get_stats, merge
    import regex as re
    chunks = [list(bytes(int(perm[x]) for x in ch.encode("utf-8"))) for ch in re.findall(base.compiled_pattern, text[:30000])]
    merges = []
    for i in range(120):
        stats = {}
        for ch in chunks:
            get_stats(ch, stats)
        pair = max(stats, key=stats.get)
        chunks = [merge(ch, pair, 256 + i) for ch in chunks]
        merges.append(pair)
    inv = {int(perm[x]): x for x in range(256)}
    tb = {i: bytes((inv[i],)) for i in range(256)}
    ranks = {bytes((x,)): int(perm[x]) for x in range(256)}
    ranks = dict(sorted(ranks.items(), key=lambda kv: kv[1]))
    for r, (p0, p1) in enumerate(merges):
        tb[256 + r] = tb[p0] + tb[p1]
        ranks[tb[256 + r]] = 256 + r
    assert len(ranks) == 256 + len(merges)
    stub = types.ModuleType("tiktoken")
    stub.get_encoding = lambda name: types.SimpleNamespace(_mergeable_ranks=ranks)
    real = sys.modules.get("tiktoken")
    sys.modules["tiktoken"] = stub
    try:
        import importlib
        import minbpe.gpt4 as g4
        importlib.reload(g4)
        tok = g4.GPT4Tokenizer()
    finally:
        if real is not None:
            sys.modules["tiktoken"] = real
        else:
            del sys.modules["tiktoken"]
    s = "<|endoftext|>" + text[:20000] + "<|fim_prefix|>x<|endofprompt|>" + text[20000:40000]
    out["gpt4_synthetic"] = {
        "perm_seed": 3, "train_chars": 30000, "n_merges": 120,
        "ranks_b64": [[__import__("base64").b64encode(k).decode(), v] for k, v in ranks.items()],
        "merges": [[int(a_), int(b_)] for (a_, b_) in tok.merges],
        "ordinary": {"text_chars": 40000, "ids_sha256": ids_sha(tok.encode_ordinary(text[:40000])), "ids_head": tok.encode_ordinary(text[:40000])[:48]},
        "special_all": {"ids_sha256": ids_sha(tok.encode(s, allowed_special="all")), "n_ids": len(tok.encode(s, allowed_special="all"))},
        # (the reference's decode indexes self.vocab only: ids of special tokens raise KeyError there — gpt4.py:88-93)
        "roundtrip_ok": tok.decode(tok.encode_ordinary(text[:40000])) == text[:40000],
    }
    with open(os.path.join(HERE, "golden_r2.json"), "w", encoding="utf-8") as f:
        json.dump(out, f)
    print("wrote golden_r2.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
