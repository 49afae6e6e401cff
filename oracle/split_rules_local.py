"""
oracle/split_rules_local.py — the GPT-4 split rules of oracle/split_rules.py restated so that every
rule reads only values AT a character and at its one or two predecessors, given six SEGMENTED scan
results (segments = class runs).  TEST INFRASTRUCTURE: this is the specification of the next version
of the device splitter: with these quantities no rule needs a value gathered
from a run's far end, so the per-byte u32 arrays of minbpe_b200/csrc/k_split.cuh can stay in shared
memory, tile by tile.  Pinned against the installed `regex` module by tests/test_split_rules.py.

Forward scans, reset at every run start:
    since[i]   characters between the run start and i (0 at the run start)
    pk[i]      kind of the character in front of the run (255 at the start of the text)
    lead[i]    every character of the run up to and including i is a newline
Backward scans, reset at every run end:
    toend[i]   characters from i to the end of the run (1 at its last character)
    nlah[i]    some character of the run at or after i is a newline
    atend[i]   the run ends at the end of the text
"""
import numpy as np

from oracle.split_rules import AP, L, N, NL, O, SP, class_table, contraction_len


def _runs(kind):
    n = len(kind)
    first = np.ones(n, dtype=bool)
    first[1:] = kind[1:] != kind[:-1]
    last = np.ones(n, dtype=bool)
    last[:-1] = first[1:]
    return first, last


def _seg_forward(first, vals, op, ident):
    """Inclusive segmented scan (sequential restatement; the device uses the associative operator
    (f1,v1)+(f2,v2) = f2 ? (1,v2) : (f1, op(v1,v2)))."""
    out = np.empty_like(vals)
    acc = ident
    for i in range(len(vals)):
        acc = vals[i] if first[i] else op(acc, vals[i])
        out[i] = acc
    return out


def chunk_starts(text):
    n = len(text)
    if n == 0:
        return np.zeros(0, dtype=bool)
    cps = np.fromiter((ord(c) for c in text), dtype=np.int64, count=n)
    cls = class_table()[cps]
    kind = cls.copy()
    kind[cls == NL] = SP
    kind[cls == AP] = O
    first, last = _runs(kind)
    is_nl = cls == NL

    # ---- the six segmented scans ----
    since = _seg_forward(first, np.ones(n, dtype=np.int64), lambda a, b: a + b, 0) - 1
    prev_kind = np.concatenate([[255], kind[:-1]])
    pk = _seg_forward(first, np.where(first, prev_kind, 0), lambda a, b: a, 0)
    lead = _seg_forward(first, is_nl.astype(np.int64), lambda a, b: a & b, 1).astype(bool)
    rfirst = last[::-1]
    toend = _seg_forward(rfirst, np.ones(n, dtype=np.int64), lambda a, b: a + b, 0)[::-1]
    nlah = _seg_forward(rfirst, is_nl[::-1].astype(np.int64), lambda a, b: a | b, 0)[::-1].astype(bool)
    is_text_end = np.zeros(n, dtype=np.int64)
    is_text_end[-1] = 1
    atend = _seg_forward(rfirst, np.where(last, is_text_end, 0)[::-1], lambda a, b: a, 0)[::-1].astype(bool)

    start = np.zeros(n, dtype=bool)
    for i in range(n):   # every rule below reads index i, i-1, i-2 (and i+1, i+2 for the contraction letters) only
        k = kind[i]
        if i == 0:
            start[i] = True
        elif k == N:
            start[i] = since[i] % 3 == 0
        elif k == O:
            start[i] = since[i] == 0 and cps[i - 1] != 0x20
        elif k == SP:
            prev_oish = pk[i] == O
            in_lead_prev = since[i] > 0 and prev_oish and lead[i - 1]          # i-1 is a newline swallowed by the Oish chunk
            if prev_oish:
                is_w2s = (not is_nl[i]) and (since[i] == 0 or lead[i - 1])
            else:
                is_w2s = since[i] == 0
            is_w3s = (not nlah[i]) and ((since[i] > 0 and is_nl[i - 1] and not in_lead_prev) or is_w2s)
            rule_a = is_w2s and nlah[i]
            rule_b = is_w3s and (atend[i] or toend[i] >= 2)
            rule_c = toend[i] == 1 and (not atend[i]) and (not is_nl[i])
            start[i] = rule_a or rule_b or rule_c
        else:   # letters
            def single_oish_start(p):   # character p is a one-character Oish run that starts a match
                return p >= 0 and kind[p] == O and (p == 0 or kind[p - 1] != O) and kind[p + 1] != O and (p == 0 or cps[p - 1] != 0x20)
            if since[i] == 0:
                start[i] = not (cls[i - 1] == SP or single_oish_start(i - 1))
            elif since[i] <= 2:
                s = i - since[i]
                if s >= 1 and cps[s - 1] == 0x27 and single_oish_start(s - 1):
                    clen = contraction_len(cps, s - 1)
                    start[i] = clen != 0 and since[i] == clen - 1
    return start


def split(text):
    st = chunk_starts(text)
    cuts = np.flatnonzero(st).tolist() + [len(text)]
    return [text[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
