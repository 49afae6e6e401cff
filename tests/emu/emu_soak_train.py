#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (BPE_LIB_PATH = the emulator build): random small corpora (tiny alphabets: ties, long runs of one
pair, pairs (a,a); RegexTokenizer chunks or one BasicTokenizer chunk) trained with the segment filter forced on and in
automatic mode, with either byte-pair histogram kernel and several batch sizes, against oracle.c_train.
    python tests/emu/emu_soak_train.py <seed> <rounds>"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, regex
import oracle
from minbpe_b200 import engine as E
GPT4 = regex.compile(r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(seed)
eng = E.Engine(0)
t0 = time.time()
for it in range(rounds):
    alpha = list("abcde  \n") if rng.random() < 0.5 else list("abcdefghijklmnopqrstuvwxyz    \n1!")
    n = int(rng.choice([50, 600, 5000, 40000, 200000]))
    text = "".join(rng.choice(alpha, size=n))
    if rng.random() < 0.3:
        text = ("ab" * int(rng.integers(1, 3000)) + "aaaa" * int(rng.integers(1, 2000)) + " ") * int(rng.integers(1, 4)) + text
    data, offs = oracle.split_to_stream(text, GPT4)
    if rng.random() < 0.25:
        offs = None                                   # BasicTokenizer: one chunk
    M = int(rng.choice([5, 40, 200, 900]))
    w = oracle.c_train(data.astype(np.int32), offs, M)
    for mode in (2, 1):
        eng.set_option(E.OPT_SEG_FILTER, mode)
        eng.set_option(E.OPT_BATCH, int(rng.choice([16, 64, 256])))
        eng.set_option(E.OPT_HIST_KERNEL, int(rng.choice([1, 2])))
        eng.load_stream(data, offs)
        p, c, d = eng.train(M)
        assert d == w[2] and np.array_equal(p, w[0]) and np.array_equal(c, w[1]), (it, mode, n, M)
print("soak train ok", seed, rounds, round(time.time() - t0), "s")
