"""
oracle/split_rules.py — a data-parallel restatement of the GPT-4 split pattern (regex.py:19)

    '(?i:[sdmt]|ll|ve|re)|[^\\r\\n\\p{L}\\p{N}]?+\\p{L}+|\\p{N}{1,3}| ?[^\\s\\p{L}\\p{N}]++[\\r\\n]*|\\s*[\\r\\n]|\\s+(?!\\S)|\\s+

as a function "is a chunk start at character i?" of LOCAL run structure only (class runs, run
lengths, a few neighbour characters), so that it can run as scans + element-wise kernels on the
GPU.  TEST INFRASTRUCTURE: this numpy version exists to pin the rules against
the installed `regex` module (tests/test_split_rules.py); the product kernel is written from it.

Character classes (one per code point):  L = \\p{L}, N = \\p{N}, NL = \\r or \\n, SP = other \\s,
AP = apostrophe, O = everything else.  "Oish" = O or AP = [^\\s\\p{L}\\p{N}].
Derivation of the rules: DESIGN.md "GPT-4 splitter".
"""
import numpy as np
import regex

L, N, NL, SP, AP, O = 0, 1, 2, 3, 4, 5
_TABLE = None
_CONTR = None


def class_table():
    """uint8[0x110000]: class of every code point, enumerated from the installed regex module."""
    global _TABLE
    if _TABLE is None:
        t = np.full(0x110000, O, dtype=np.uint8)
        pl, pn, ps = regex.compile(r"\p{L}"), regex.compile(r"\p{N}"), regex.compile(r"\s")
        for cp in range(0x110000):
            if 0xD800 <= cp <= 0xDFFF:
                continue
            ch = chr(cp)
            if pl.match(ch):
                t[cp] = L
            elif pn.match(ch):
                t[cp] = N
            elif ps.match(ch):
                t[cp] = SP
        t[ord("\r")] = NL
        t[ord("\n")] = NL
        t[ord("'")] = AP
        _TABLE = t
    return _TABLE


def contraction_sets():
    """Code points matching (?i:s|d|m|t), and (?i:l|v|e|r) — for 's 'd 'm 't 'll 've 're."""
    global _CONTR
    if _CONTR is None:
        sets = {}
        for letter in "sdmtlver":
            p = regex.compile("(?i:%s)" % letter)
            sets[letter] = {cp for cp in range(0x3000) if not (0xD800 <= cp <= 0xDFFF) and p.fullmatch(chr(cp))}
        _CONTR = sets
    return _CONTR


def contraction_len(cps, i):
    """Length in characters (2 or 3) of a contraction starting at the apostrophe cps[i], or 0."""
    S = contraction_sets()
    n = len(cps)
    if i + 1 < n:
        c1 = int(cps[i + 1])
        if c1 in S["s"] or c1 in S["d"] or c1 in S["m"] or c1 in S["t"]:
            return 2
        if i + 2 < n:
            c2 = int(cps[i + 2])
            if (c1 in S["l"] and c2 in S["l"]) or (c1 in S["v"] and c2 in S["e"]) or (c1 in S["r"] and c2 in S["e"]):
                return 3
    return 0


def chunk_starts(text):
    """bool[len(text)]: True where a match of the GPT-4 pattern starts.  Vectorised over runs."""
    n = len(text)
    if n == 0:
        return np.zeros(0, dtype=bool)
    cps = np.fromiter((ord(c) for c in text), dtype=np.int64, count=n)
    cls = class_table()[cps]
    kind = cls.copy()                      # run type: L, N, WS (=SP), OI (=O)
    kind[cls == NL] = SP
    kind[cls == AP] = O
    idx = np.arange(n)
    run_start_flag = np.ones(n, dtype=bool)
    run_start_flag[1:] = kind[1:] != kind[:-1]
    rs = np.maximum.accumulate(np.where(run_start_flag, idx, 0))                     # start of the run containing i
    run_end_flag = np.ones(n, dtype=bool)
    run_end_flag[:-1] = kind[1:] != kind[:-1]
    re_ = np.minimum.accumulate(np.where(run_end_flag, idx, n)[::-1])[::-1] + 1      # one past the end of the run
    is_nl = cls == NL
    last_nl = np.maximum.accumulate(np.where(is_nl, idx, -1))                        # last NL at or before i
    next_non_nl = np.minimum.accumulate(np.where(~is_nl, idx, n)[::-1])[::-1]        # first non-NL at or after i
    prev_cls = np.concatenate([[255], cls[:-1]])
    prev_cp = np.concatenate([[-1], cps[:-1]])

    start = np.zeros(n, dtype=bool)
    start[0] = True

    # ---- digits: groups of three from the start of the run ----
    isN = kind == N
    start |= isN & ((idx - rs) % 3 == 0)

    # ---- Oish runs: one chunk, starting at the run start, or one earlier if a space precedes it ----
    isO = kind == O
    o_first = isO & run_start_flag
    start |= o_first & (prev_cp != 0x20)          # (with a space in front, the start is that space: whitespace rule C)

    # ---- whitespace runs ----
    isW = kind == SP
    ws = rs
    we = re_
    prev_of_run_is_oish = np.zeros(n, dtype=bool)
    has_prev = ws > 0
    pk = np.where(has_prev, kind[np.maximum(ws - 1, 0)], 255)
    prev_of_run_is_oish = pk == O
    w2s = np.where(prev_of_run_is_oish, np.minimum(next_non_nl[ws], we), ws)          # after the NLs taken by A4
    lnl = last_nl[we - 1]                                                             # last NL of the run (or before it)
    has_nl = lnl >= w2s
    w3s = np.where(has_nl, lnl + 1, w2s)
    k = we - w3s
    # A: \s*[\r\n] from w2s through the last NL
    start |= isW & (idx == w2s) & has_nl & (w2s < we)
    # B: \s+(?!\S): the spaces after the last NL, all of them at the end of the text, else all but the last
    start |= isW & (idx == w3s) & (k >= 1) & ((we == n) | (k >= 2))
    # C: the last space, when something non-space follows: it starts " word", " !!!" or a one-character chunk
    start |= isW & (idx == we - 1) & (we < n) & (k >= 1)

    # ---- letter runs ----
    isL = kind == L
    l_first = isL & run_start_flag
    # absorbed by "[^\r\n\p{L}\p{N}]?+\p{L}+": the previous character is a non-newline whitespace (it always
    # starts a match), or a single Oish character that itself starts a match (not preceded by a space)
    prev_is_sp = prev_cls == SP
    prev_single_oish = np.zeros(n, dtype=bool)
    i1 = np.maximum(idx - 1, 0)
    prev_is_oish = (idx > 0) & (kind[i1] == O)
    prev_run_len1 = prev_is_oish & (rs[i1] == i1)
    prev_prev_cp = np.concatenate([[-1, -1], cps])[:n]
    prev_single_oish = prev_run_len1 & (prev_prev_cp != 0x20)
    absorbed = l_first & (prev_is_sp | prev_single_oish)
    start |= l_first & ~absorbed
    # contractions: an apostrophe that starts a match (single Oish char, no space in front) followed by
    # s/d/m/t/ll/ve/re takes one or two letters of the run; the rest of the run starts a new chunk
    for i in np.flatnonzero(l_first & prev_single_oish & (prev_cp == 0x27)):
        clen = contraction_len(cps, i - 1)
        if clen:
            start[i] = False
            j = i - 1 + clen
            if j < re_[i]:
                start[j] = True
    return start


def split(text):
    st = chunk_starts(text)
    cuts = np.flatnonzero(st).tolist() + [len(text)]
    return [text[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
