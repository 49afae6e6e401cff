#!/usr/bin/env python3
"""
Generates tests/golden/* from the UNMODIFIED reference (karpathy/minbpe mounted read-only at
/root/reference).  Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (all committed):
  taylorswift.txt          the reference's own test corpus (tests/taylorswift.txt), fixture data
  ref_basic512.model/.vocab, ref_regex512.model/.vocab   files written by the reference's save()
  golden_train.json        merges + verbose counts + encode fingerprints for the named configs
  golden_cases.json        small adversarial train/encode cases (runs, ties, exhaustion, unicode)
  golden_primitives.json   get_stats / merge input->output vectors
Nothing here is imported by the product; tests read the JSON.
"""
import contextlib
import hashlib
import io
import json
import os
import random
import re as stdre
import shutil
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from minbpe import BasicTokenizer, RegexTokenizer  # noqa: E402  (the reference)
from minbpe.base import get_stats, merge  # noqa: E402

from minbpe_b200.synth import generate as synth_generate  # noqa: E402  (our generator)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def ids_sha(ids):
    import numpy as np
    return sha(np.asarray(ids, dtype="<i4").tobytes())


def train_verbose(cls, text, vocab_size):
    """Train the reference with verbose=True; return (tokenizer, [counts]) or the exception name."""
    tok = cls()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tok.train(text, vocab_size, verbose=True)
    counts = [int(x) for x in stdre.findall(r"had (\d+) occurrences", buf.getvalue())]
    return tok, counts, buf.getvalue()


def case(cls_name, text, vocab_size, encode_texts=()):
    cls = {"basic": BasicTokenizer, "regex": RegexTokenizer}[cls_name]
    rec = {"tokenizer": cls_name, "text": text, "vocab_size": vocab_size}
    try:
        tok, counts, _ = train_verbose(cls, text, vocab_size)
    except Exception as e:  # noqa: BLE001
        rec["raises"] = type(e).__name__
        # how many merges the reference completed before running out of pairs
        done = 0
        for k in range(vocab_size - 256, -1, -1):
            try:
                cls().train(text, 256 + k)
                done = k
                break
            except ValueError:
                continue
        rec["n_done"] = done
        tok, counts, _ = train_verbose(cls, text, 256 + done)  # the merges it did complete
        rec["merges"] = [[a, b] for (a, b) in tok.merges]
        rec["counts"] = counts
        return rec
    rec["merges"] = [[a, b] for (a, b) in tok.merges]
    rec["counts"] = counts
    rec["encode"] = {t: tok.encode(t) for t in (text,) + tuple(encode_texts)}
    return rec


def main():
    random.seed(20260922)
    # ---- the reference's fixture corpus -------------------------------------------------
    shutil.copyfile(os.path.join(REF, "tests", "taylorswift.txt"), os.path.join(HERE, "taylorswift.txt"))
    text = open(os.path.join(HERE, "taylorswift.txt"), "r", encoding="utf-8").read()

    train = {}
    for name, cls in (("basic", BasicTokenizer), ("regex", RegexTokenizer)):
        tok, counts, verbose_out = train_verbose(cls, text, 512)
        ids = tok.encode(text)
        with tempfile.TemporaryDirectory() as d:
            tok.save(os.path.join(d, "m"))
            for ext in ("model", "vocab"):
                shutil.copyfile(os.path.join(d, "m." + ext), os.path.join(HERE, f"ref_{name}512.{ext}"))
            model_sha = sha(open(os.path.join(d, "m.model"), "rb").read())
            vocab_sha = sha(open(os.path.join(d, "m.vocab"), "rb").read())
        train[f"taylorswift_{name}_512"] = {
            "merges": [[a, b] for (a, b) in tok.merges],
            "counts": counts,
            "merges_sha256": sha("\n".join(f"{a} {b}" for a, b in tok.merges).encode()),
            "n_ids": len(ids),
            "ids_sha256": ids_sha(ids),
            "ids_head": ids[:64],
            "model_sha256": model_sha,
            "vocab_sha256": vocab_sha,
            "verbose_sha256": sha(verbose_out.encode()),
            "verbose_head": verbose_out.splitlines()[:3],
            "decode_roundtrip": tok.decode(ids) == text,
        }

    # ---- Wikipedia known-answer test (tests/test_tokenizer.py:80-107) --------------------
    for name, cls in (("basic", BasicTokenizer), ("regex", RegexTokenizer)):
        tok = cls()
        tok.train("aaabdaaabac", 259)
        train[f"wikipedia_{name}"] = {
            "merges": [[a, b] for (a, b) in tok.merges],
            "vocab_tail": [list(tok.vocab[i]) for i in (256, 257, 258)],
            "ids": tok.encode("aaabdaaabac"),
        }

    # ---- save/load + special tokens (tests/test_tokenizer.py:109-132) --------------------
    # the reference test module holds the fixture strings (tests/test_tokenizer.py:28-46)
    from tests.test_tokenizer import llama_text, specials_string  # noqa: E402
    open(os.path.join(HERE, "llama_text.txt"), "w", encoding="utf-8").write(llama_text)
    open(os.path.join(HERE, "specials_string.txt"), "w", encoding="utf-8").write(specials_string)
    specials = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|fim_middle|>": 100259,
                "<|fim_suffix|>": 100260, "<|endofprompt|>": 100276}
    tok = RegexTokenizer()
    tok.train(llama_text, 256 + 64)
    tok.register_special_tokens(specials)
    with tempfile.TemporaryDirectory() as d:
        tok.save(os.path.join(d, "m"))
        model_text = open(os.path.join(d, "m.model"), "r", encoding="utf-8").read()
        vocab_text = open(os.path.join(d, "m.vocab"), "r", encoding="utf-8").read()
    train["llama_regex_320_specials"] = {
        "merges": [[a, b] for (a, b) in tok.merges],
        "specials": specials,
        "ids_all": tok.encode(llama_text, "all"),
        "ids_none": tok.encode(llama_text, "none"),
        "ids_subset": tok.encode(llama_text, {"<|endoftext|>", "<|fim_suffix|>"}),
        "model_text": model_text,
        "vocab_text": vocab_text,
    }

    # ---- synthetic corpus prefix (seed 1337), both tokenizers ----------------------------
    syn = bytes(synth_generate(1337, 256 * 1024)).decode("utf-8")
    for name, cls, V in (("basic", BasicTokenizer, 300), ("regex", RegexTokenizer, 320)):
        tok, counts, _ = train_verbose(cls, syn, V)
        ids = tok.encode(syn) if name == "regex" else None
        train[f"synth1337_256k_{name}_{V}"] = {
            "corpus_sha256": sha(syn.encode("utf-8")),
            "merges": [[a, b] for (a, b) in tok.merges],
            "counts": counts,
            "n_ids": None if ids is None else len(ids),
            "ids_sha256": None if ids is None else ids_sha(ids),
        }
    json.dump(train, open(os.path.join(HERE, "golden_train.json"), "w"), indent=0, ensure_ascii=True)

    # ---- small adversarial cases ---------------------------------------------------------
    cases = []
    texts = [
        ("aaabdaaabac", 259), ("aaaaaaa", 259), ("aaaaaaaa", 260), ("abababab", 259),
        ("aaaa bbbb aaaa bbbb", 262), ("a" * 33 + "b" + "a" * 32, 262), ("====----====", 260),
        ("hello world!!!? (안녕하세요!) lol123 😉", 270), ("the the the then then", 266),
        ("abcd", 259), ("ab cd", 262), ("", 257), ("a", 257), ("ab", 258), ("ab", 257),
        ("xyxyzxyxyz" * 7, 262), ("aXbXaXbXaa", 261), ("  \n\n  \n\t\t  x  ", 262),
        ("I'll we're DON'T it's 12345 6789 x1y22z333", 272), ("\r\n\r\n a\r\nb \r\n", 262),
        ("zzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzzz", 262),
    ]
    rnd = random.Random(7)
    for alpha, n, V in (("ab", 60, 262), ("abc", 200, 270), ("ab \n", 150, 266), ("aé😉 1", 80, 268),
                        ("ab", 500, 280), ("abcdefgh ", 2000, 300), ("a ", 64, 262)):
        texts.append(("".join(rnd.choice(alpha) for _ in range(n)), V))
    for t, V in texts:
        for name in ("basic", "regex"):
            extra = ("aaaa", "ab ab", t[: len(t) // 2]) if t else ()
            cases.append(case(name, t, V, extra))
    json.dump(cases, open(os.path.join(HERE, "golden_cases.json"), "w"), indent=0, ensure_ascii=True)

    # ---- primitives ----------------------------------------------------------------------
    prims = {"get_stats": [], "merge": []}
    lists = [[7, 7, 7, 7, 7], [7, 7, 7, 7], [1, 2, 3, 1, 2], [], [5], [5, 5], [1, 2, 1, 2, 1, 2, 1]]
    for k in range(40):
        n = rnd.randint(0, 60)
        V = rnd.choice([2, 3, 5, 300])
        lists.append([rnd.randrange(V) for _ in range(n)])
    for L in lists:
        st = get_stats(L)
        prims["get_stats"].append({"ids": L, "items": [[a, b, c] for (a, b), c in st.items()]})
        pairs = list(st)[:3] + [(0, 0), (1, 1), (7, 7)]
        for p in pairs:
            prims["merge"].append({"ids": L, "pair": list(p), "idx": 999, "out": merge(L, p, 999)})
    # get_stats with an existing counts dict (base.py:19: counts passed in and updated)
    acc = {}
    get_stats([1, 2, 3], acc)
    get_stats([2, 3, 4, 1, 2], acc)
    prims["get_stats_accumulate"] = {"lists": [[1, 2, 3], [2, 3, 4, 1, 2]],
                                     "items": [[a, b, c] for (a, b), c in acc.items()]}
    json.dump(prims, open(os.path.join(HERE, "golden_primitives.json"), "w"), indent=0)
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
