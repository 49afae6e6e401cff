"""
CPU stand-in for the engine's step API (tests only): lets tests/test_dist_gloo.py drive
minbpe_b200.dist.ShardedTrainer over gloo without a GPU.  Local work is done with plain Python
loops in the oracle's spirit (it is test infrastructure, never shipped): pair table as a dict,
merge by the greedy left-to-right rule, and the L / R / ZZ statistics delta of DESIGN.md computed
independently of the CUDA kernels.
"""
import numpy as np
import torch

CAND_NONE = (1 << 63) - 1


class CpuStepEngine:
    def __init__(self, data, offsets):
        self.ids = [int(b) for b in bytes(data)]
        self.start = [False] * len(self.ids)
        if self.ids:
            self.start[0] = True
        for o in ([] if offsets is None else offsets):
            self.start[int(o)] = True
        self.tab = {}
        self.log = []
        self.done_flag = False

    def new_i64(self, n):
        return torch.zeros(n, dtype=torch.int64)

    def pairs(self):
        for k in range(len(self.ids) - 1):
            if not self.start[k + 1]:
                yield k, (self.ids[k], self.ids[k + 1])

    def begin(self, dense):
        dense.zero_()
        for _, (a, b) in self.pairs():
            dense[a * 256 + b] += 1

    def table_(self, dense):
        self.tab = {(i // 256, i % 256): int(c) for i, c in enumerate(dense.tolist()) if c}

    def table(self, dense, num_merges, first_idx, poll_every):
        self.table_(dense)
        self.first_idx, self.V = first_idx, first_idx + num_merges
        self.max_iter = num_merges
        return 2 * self.V + 1

    def select(self, cand, rank):
        live = {p: c for p, c in self.tab.items() if c > 0}
        if self.done_flag or not live or len(self.log) >= self.max_iter:
            cand[0] = CAND_NONE
            self.best = 0
            return
        best = max(live.values())
        self.best = best
        tied = {p for p, c in live.items() if c == best}
        word = CAND_NONE
        if len(tied) == 1:
            (p0, p1), = tied
            word = (rank << 58) | (p0 << 29) | p1
        else:
            for _, p in self.pairs():  # first occurrence in this shard
                if p in tied:
                    word = (rank << 58) | (p[0] << 29) | p[1]
                    break
        cand[0] = word

    def merge(self, cand, delta):
        if self.done_flag or len(self.log) >= self.max_iter:
            return
        w = int(cand[0])
        if w == CAND_NONE:
            self.done_flag = True
            return
        a, b = (w >> 29) & 0x1FFFFFFF, w & 0x1FFFFFFF
        z = self.first_idx + len(self.log)
        self.log.append(((a, b), self.best))
        self.cur = (a, b, z)
        ids, st, n = self.ids, self.start, len(self.ids)
        # greedy left-to-right merge starts (base.py:33-40), chunk aware
        m = [False] * n
        i = 0
        while i < n:
            if ids[i] == a and i + 1 < n and not st[i + 1] and ids[i + 1] == b:
                m[i] = True
                i += 2
            else:
                i += 1
        V = self.V
        for p in range(n):
            if not m[p]:
                continue
            if p >= 1 and not st[p] and not (p >= 2 and m[p - 2]):
                delta[ids[p - 1]] += 1                      # L[x]
            if p + 2 < n and not st[p + 2]:
                if m[p + 2]:
                    delta[2 * V] += 1                       # ZZ
                else:
                    delta[V + ids[p + 2]] += 1              # R[y]
        out, ost = [], []
        i = 0
        while i < n:
            if m[i]:
                out.append(z); ost.append(st[i]); i += 2
            else:
                out.append(ids[i]); ost.append(st[i]); i += 1
        self.ids, self.start = out, ost

    def apply(self, delta):
        if self.done_flag or not hasattr(self, "cur") or self.cur is None:
            delta.zero_()
            return
        a, b, z = self.cur
        V = self.V
        d = delta.tolist()
        t = self.tab
        for x in range(V):
            if d[x]:
                if (x, a) != (a, b):
                    t[(x, a)] = t.get((x, a), 0) - d[x]
                t[(x, z)] = t.get((x, z), 0) + d[x]
            if d[V + x]:
                if (b, x) != (a, b):
                    t[(b, x)] = t.get((b, x), 0) - d[V + x]
                t[(z, x)] = t.get((z, x), 0) + d[V + x]
        if d[2 * V]:
            if (b, a) != (a, b):
                t[(b, a)] = t.get((b, a), 0) - d[2 * V]
            t[(z, z)] = t.get((z, z), 0) + d[2 * V]
        t[(a, b)] = 0
        assert all(c >= 0 for c in t.values())
        delta.zero_()
        self.cur = None

    def poll(self):
        return len(self.log), self.done_flag

    def result(self, cap):
        pairs = np.array([list(p) for p, _ in self.log], dtype=np.int32).reshape(-1, 2)
        counts = np.array([c for _, c in self.log], dtype=np.int64)
        return pairs, counts, len(self.log)
