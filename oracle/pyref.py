"""
oracle/pyref.py — pure-Python restatement of the minbpe hot path for SMALL cases.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity: pinned by tests/test_oracle.py
against tests/golden/*.json generated from the unmodified reference.

Each function cites the reference lines (karpathy/minbpe @1acefe8) whose behaviour it restates.
"""


def get_stats(ids, counts=None):
    """base.py:13-22 — adjacent-pair histogram (overlaps counted), insertion ordered; updates
    and returns ``counts`` when given."""
    table = {} if counts is None else counts
    prev = None
    for cur in ids:
        if prev is not None:
            key = (prev, cur)
            table[key] = table.get(key, 0) + 1
        prev = cur
    return table


def merge(ids, pair, idx):
    """base.py:25-41 — greedy left-to-right, non-overlapping replacement of ``pair`` by ``idx``."""
    first, second = pair
    out = []
    k, n = 0, len(ids)
    while k < n:
        if k + 1 < n and ids[k] == first and ids[k + 1] == second:
            out.append(idx)
            k += 2
        else:
            out.append(ids[k])
            k += 1
    return out


def argmax_pair(stats):
    """basic.py:35 / regex.py:56 — max(stats, key=stats.get): highest count, earliest-inserted
    key wins ties.  Raises ValueError on an empty table exactly like max({})."""
    best, best_count = None, None
    for key, c in stats.items():
        if best_count is None or c > best_count:
            best, best_count = key, c
    if best is None:
        raise ValueError("max() iterable argument is empty")
    return best


def train(chunks, num_merges, first_idx=256):
    """basic.py:31-45 (``chunks`` = [whole text bytes]) and regex.py:49-66 (``chunks`` = the
    regex pieces).  Returns ([(pair, count)], final chunk id lists)."""
    streams = [list(c) for c in chunks]
    log = []
    for i in range(num_merges):
        stats = {}
        for s in streams:          # regex.py:51-54: one table, chunks in order
            get_stats(s, stats)
        pair = argmax_pair(stats)  # raises ValueError when no pair is left
        idx = first_idx + i
        streams = [merge(s, pair, idx) for s in streams]  # regex.py:60
        log.append((pair, stats[pair]))
    return log, streams


def encode_chunk(chunk_bytes, merges):
    """regex.py:92-109 / basic.py:57-74 — ``merges``: dict pair -> id.  Repeatedly merge the
    present pair with the lowest id until no present pair is mergeable."""
    ids = list(chunk_bytes)
    while len(ids) >= 2:
        stats = get_stats(ids)
        pair = min(stats, key=lambda p: merges.get(p, float("inf")))
        if pair not in merges:
            break
        ids = merge(ids, pair, merges[pair])
    return ids


def encode_ordinary(chunks, merges):
    """regex.py:111-121 — encode each chunk, concatenate in order."""
    out = []
    for c in chunks:
        out.extend(encode_chunk(c, merges))
    return out


def build_vocab(merges, special_tokens=None):
    """base.py:88-95."""
    vocab = {i: bytes([i]) for i in range(256)}
    for (p0, p1), idx in merges.items():
        vocab[idx] = vocab[p0] + vocab[p1]
    for s, idx in (special_tokens or {}).items():
        vocab[idx] = s.encode("utf-8")
    return vocab
