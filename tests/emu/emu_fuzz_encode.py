#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (BPE_LIB_PATH = the emulator build): random small inputs through both encode entry points
(bpe_encode with host offsets, bpe_encode_text_gpt4) against oracle.c_encode — sizes around the kernels' tile (2048 B),
halo (64 B), memo limit (48 B) and long-chunk (8192 tokens) boundaries, tiny memo tables, several pieces per call.
Allocations sit in front of guard pages, so an out-of-bounds access of a kernel ends the process."""
import os
import sys

import numpy as np
import regex

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from minbpe_b200 import engine as E  # noqa: E402

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def random_text(rng, n_words, max_word):
    alphabet = list("abcdeé ") + ["\n", "1", "!", "日"]
    words = []
    for _ in range(n_words):
        k = int(rng.integers(1, max_word + 1))
        if rng.random() < 0.02:
            k = int(rng.choice([31, 32, 33, 34, 46, 47, 48, 49, 50, 63, 64, 65, 66, 100, 2040, 2048, 2056, 8190, 8192]))
            words.append("".join(rng.choice(list("abc"), size=k)))
        else:
            words.append("".join(rng.choice(alphabet, size=k)))
    return " ".join(words)


def main(rounds, seed):
    rng = np.random.default_rng(seed)
    eng = E.Engine(0)
    total = 0
    for it in range(rounds):
        text = random_text(rng, int(rng.integers(1, 400)), int(rng.choice([3, 8, 40])))
        if rng.random() < 0.15:          # past the 64 KiB threshold of bpe_encode: the memoised kernels with flags from host offsets
            text = (text + " ") * (70000 // (len(text.encode("utf-8")) + 1) + 1)
        # cut to interesting byte lengths now and then
        raw = text.encode("utf-8")
        if rng.random() < 0.4:
            cut = int(rng.choice([1, 2, 7, 8, 9, 2047, 2048, 2049, 2048 + 63, 2048 + 64, 2048 + 65, 4095, 4096, 4097]))
            raw = raw[:cut]
            text = raw.decode("utf-8", errors="ignore")
        if not text:
            continue
        data, offs = oracle.split_to_stream(text, GPT4)
        # merges trained on the text itself (few), so that multi-level merges exist
        eng.load_stream(data, offs)
        want_m = int(rng.integers(0, 60))
        merges, _, done = eng.train(want_m)
        merges = merges[:done]
        perm = rng.permutation(256).astype(np.uint8) if rng.random() < 0.2 else None
        w = oracle.c_encode(data, offs, merges, perm)
        memo = int(rng.choice([0, 0, 6, 8]))
        piece = int(rng.choice([0, 0, 4096, 1 << 14]))
        eng.set_option(E.OPT_ENC_MEMO_LOG2, memo)
        eng.set_option(E.OPT_SPLIT_PIECE, piece)
        try:
            g1 = eng.encode(data, offs, merges, perm) if len(merges) else None
            try:
                g2 = eng.encode_text_gpt4(data.tobytes(), merges, perm)
            except E.EngineError as ex:          # a tiny piece size may find no letter+space cut: a clean error, not a wrong answer
                assert "cut point" in str(ex), ex
                g2 = None
        finally:
            eng.set_option(E.OPT_ENC_MEMO_LOG2, 0)
            eng.set_option(E.OPT_SPLIT_PIECE, 0)
        if g1 is not None:
            assert np.array_equal(g1, w), (it, "encode(offsets)", len(raw), memo, piece)
        if g2 is not None:
            assert np.array_equal(g2, w), (it, "encode_text_gpt4", len(raw), memo, piece)
        total += len(raw)
    eng.close()
    print(f"emu fuzz encode ok: {rounds} rounds, {total} bytes")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
