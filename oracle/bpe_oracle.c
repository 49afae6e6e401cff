/*
 * oracle/bpe_oracle.c — CPU restatement of the minbpe hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (minbpe_b200 + libb200bpe.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against the
 * golden vectors in tests/golden/ that were produced by importing the unmodified
 * reference (tests/golden/make_golden.py), including the reference's own known-answer
 * test (tests/test_tokenizer.py:80-107, "aaabdaaabac" -> [258,100,258,97,99]).
 *
 * What is restated (reference file:line, relative to karpathy/minbpe @1acefe8):
 *   orc_get_stats   minbpe/base.py:13-22   adjacent pair histogram, overlaps counted,
 *                                          result in first-occurrence (dict insertion) order
 *   orc_merge       minbpe/base.py:25-41   greedy left-to-right non-overlapping replace
 *   orc_train       minbpe/basic.py:31-45  (one chunk) and minbpe/regex.py:49-66 (many
 *                                          chunks): stats over chunks in order, max() with
 *                                          first-inserted tie-break, per-chunk merge
 *   orc_encode      minbpe/regex.py:92-121 / minbpe/basic.py:57-74: per chunk, repeatedly
 *                                          merge the present pair with the lowest merge index
 *
 * Data layout: a corpus is one flat int32 array of token ids plus chunk start offsets
 * (offs[0]=0 < offs[1] < ... < offs[n_chunks-1] < n; chunk k = [offs[k], offs[k+1]) ).
 * No pair and no merge crosses a chunk boundary (regex.py:51-54,60).  BasicTokenizer is
 * the one-chunk case (n_chunks<=1 or offs==NULL).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ENOMEM -1
#define ORC_ECAP -2

/* ------------------------------------------------------------------------------------
 * insertion-ordered pair -> count map (what a Python dict keyed by (int,int) gives the
 * reference: lookup by key, iteration in first-insertion order)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *keys;   /* per ordinal: packed pair (p0<<32 | p1) in insertion order */
    int64_t *counts;  /* per ordinal */
    uint64_t n;       /* number of distinct pairs */
    uint64_t cap_ord; /* capacity of keys/counts */
    uint64_t *tab;    /* open addressing: ordinal+1, 0 = empty */
    uint64_t tab_mask;
} pairmap;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

static int pm_init(pairmap *m, uint64_t hint) {
    uint64_t t = 1024;
    while (t < hint * 2) t <<= 1;
    m->n = 0;
    m->cap_ord = t / 2;
    m->tab_mask = t - 1;
    m->keys = (uint64_t *)malloc(m->cap_ord * sizeof(uint64_t));
    m->counts = (int64_t *)malloc(m->cap_ord * sizeof(int64_t));
    m->tab = (uint64_t *)calloc(t, sizeof(uint64_t));
    return (m->keys && m->counts && m->tab) ? ORC_OK : ORC_ENOMEM;
}

static void pm_free(pairmap *m) {
    free(m->keys); free(m->counts); free(m->tab);
    m->keys = NULL; m->counts = NULL; m->tab = NULL;
}

static void pm_clear(pairmap *m) {
    memset(m->tab, 0, (m->tab_mask + 1) * sizeof(uint64_t));
    m->n = 0;
}

static int pm_grow(pairmap *m) {
    uint64_t t = (m->tab_mask + 1) * 2;
    uint64_t *nk = (uint64_t *)realloc(m->keys, (t / 2) * sizeof(uint64_t));
    if (!nk) return ORC_ENOMEM;
    m->keys = nk;
    int64_t *nc = (int64_t *)realloc(m->counts, (t / 2) * sizeof(int64_t));
    if (!nc) return ORC_ENOMEM;
    m->counts = nc;
    free(m->tab);
    m->tab = (uint64_t *)calloc(t, sizeof(uint64_t));
    if (!m->tab) return ORC_ENOMEM;
    m->tab_mask = t - 1;
    m->cap_ord = t / 2;
    for (uint64_t o = 0; o < m->n; ++o) {
        uint64_t h = mix64(m->keys[o]) & m->tab_mask;
        while (m->tab[h]) h = (h + 1) & m->tab_mask;
        m->tab[h] = o + 1;
    }
    return ORC_OK;
}

/* counts[pair] = counts.get(pair, 0) + w   (base.py:21) */
static int pm_add(pairmap *m, uint64_t key, int64_t w) {
    uint64_t h = mix64(key) & m->tab_mask;
    for (;;) {
        uint64_t o = m->tab[h];
        if (!o) break;
        if (m->keys[o - 1] == key) { m->counts[o - 1] += w; return ORC_OK; }
        h = (h + 1) & m->tab_mask;
    }
    if (m->n == m->cap_ord) {
        int rc = pm_grow(m);
        if (rc) return rc;
        h = mix64(key) & m->tab_mask;
        while (m->tab[h]) h = (h + 1) & m->tab_mask;
    }
    m->keys[m->n] = key;
    m->counts[m->n] = w;
    m->tab[h] = ++m->n;
    return ORC_OK;
}

static int64_t pm_find(const pairmap *m, uint64_t key) { /* ordinal or -1 */
    uint64_t h = mix64(key) & m->tab_mask;
    for (;;) {
        uint64_t o = m->tab[h];
        if (!o) return -1;
        if (m->keys[o - 1] == key) return (int64_t)(o - 1);
        h = (h + 1) & m->tab_mask;
    }
}

static inline uint64_t pack(int32_t a, int32_t b) {
    return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
}

/* start[k] != 0  <=>  token k is the first token of its chunk */
static uint8_t *make_start_flags(uint64_t n, const uint64_t *offs, uint64_t n_chunks) {
    uint8_t *s = (uint8_t *)calloc(n ? n : 1, 1);
    if (!s) return NULL;
    if (n) s[0] = 1;
    if (offs) for (uint64_t k = 0; k < n_chunks; ++k) if (offs[k] < n) s[offs[k]] = 1;
    return s;
}

/* accumulate the pair histogram of one stream into m  (base.py:19-22 applied per chunk,
 * regex.py:51-54: one dict across all chunks, chunks visited in order) */
static int stats_into(pairmap *m, const int32_t *ids, const uint8_t *start, uint64_t n,
                      const int64_t *weight_of_chunk) {
    int64_t w = 1;
    uint64_t chunk = (uint64_t)-1;
    for (uint64_t k = 0; k < n; ++k) {
        if (start[k]) { ++chunk; if (weight_of_chunk) w = weight_of_chunk[chunk]; continue; }
        int rc = pm_add(m, pack(ids[k - 1], ids[k]), w);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * base.py:13-22  get_stats
 * out_pairs[2*i], out_pairs[2*i+1], out_counts[i] in dict insertion order.
 * ---------------------------------------------------------------------------------- */
int orc_get_stats(const int32_t *ids, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
                  int32_t *out_pairs, int64_t *out_counts, uint64_t cap, uint64_t *n_pairs) {
    pairmap m;
    if (pm_init(&m, 1024)) return ORC_ENOMEM;
    uint8_t *start = make_start_flags(n, offs, n_chunks);
    if (!start) { pm_free(&m); return ORC_ENOMEM; }
    int rc = stats_into(&m, ids, start, n, NULL);
    if (!rc) {
        *n_pairs = m.n;
        if (m.n > cap) rc = ORC_ECAP;
        else for (uint64_t o = 0; o < m.n; ++o) {
            out_pairs[2 * o] = (int32_t)(m.keys[o] >> 32);
            out_pairs[2 * o + 1] = (int32_t)(m.keys[o] & 0xffffffffu);
            out_counts[o] = m.counts[o];
        }
    }
    free(start); pm_free(&m);
    return rc;
}

/* base.py:31-41 on a flat stream with chunk starts; in-place capable (out may alias ids).
 * Returns the new length; start flags are rewritten for the new stream. */
static uint64_t merge_stream(int32_t *out, uint8_t *out_start, const int32_t *ids,
                             const uint8_t *start, uint64_t n, int32_t a, int32_t b, int32_t idx) {
    uint64_t i = 0, j = 0;
    while (i < n) {
        /* "i < len(ids) - 1" inside a chunk: the next token exists and is not a chunk start */
        if (ids[i] == a && i + 1 < n && !start[i + 1] && ids[i + 1] == b) {
            uint8_t s = start[i];
            out[j] = idx; out_start[j] = s; ++j; i += 2;
        } else {
            uint8_t s = start[i];
            out[j] = ids[i]; out_start[j] = s; ++j; i += 1;
        }
    }
    return j;
}

/* ------------------------------------------------------------------------------------
 * base.py:25-41  merge   (out needs room for n ids)
 * ---------------------------------------------------------------------------------- */
int orc_merge(const int32_t *ids, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
              int32_t a, int32_t b, int32_t idx, int32_t *out, uint64_t *out_n) {
    uint8_t *start = make_start_flags(n, offs, n_chunks);
    uint8_t *ostart = (uint8_t *)malloc(n ? n : 1);
    if (!start || !ostart) { free(start); free(ostart); return ORC_ENOMEM; }
    *out_n = merge_stream(out, ostart, ids, start, n, a, b, idx);
    free(start); free(ostart);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------
 * basic.py:31-45 / regex.py:49-66  the training loop.
 *   ids/n/offs/n_chunks : initial stream (bytes widened to int32) and chunk starts
 *   weights             : NULL, or one multiplicity per chunk (the de-duplicated variant:
 *                         unique chunks in first-occurrence order, each counted weight
 *                         times; SURVEY.md §8c — same merges incl. tie-breaks)
 *   out_pairs[2*i..], out_counts[i] : pair chosen at merge i and stats[pair] (the number
 *                         the reference prints in verbose mode, basic.py:45)
 *   n_done              : merges actually performed; < num_merges iff the stream ran out
 *                         of pairs, where the reference raises ValueError (max() of an
 *                         empty dict, basic.py:35 / regex.py:56)
 *   final_ids/final_n   : optional (may be NULL): the stream after the last merge
 * ---------------------------------------------------------------------------------- */
int orc_train(const int32_t *ids_in, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
              const int64_t *weights, int32_t num_merges, int32_t first_idx,
              int32_t *out_pairs, int64_t *out_counts, int32_t *n_done,
              int32_t *final_ids, uint64_t *final_n) {
    int rc = ORC_OK;
    int32_t *ids = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    uint8_t *start = make_start_flags(n, offs, n_chunks);
    pairmap m;
    if (!ids || !start || pm_init(&m, 4096)) { free(ids); free(start); return ORC_ENOMEM; }
    memcpy(ids, ids_in, n * sizeof(int32_t));
    *n_done = 0;
    for (int32_t i = 0; i < num_merges; ++i) {
        pm_clear(&m);
        rc = stats_into(&m, ids, start, n, weights);
        if (rc) break;
        if (m.n == 0) break; /* reference: max({}) -> ValueError */
        /* max(stats, key=stats.get): the first key in insertion order with the max value */
        uint64_t best = 0;
        for (uint64_t o = 1; o < m.n; ++o) if (m.counts[o] > m.counts[best]) best = o;
        int32_t a = (int32_t)(m.keys[best] >> 32), b = (int32_t)(m.keys[best] & 0xffffffffu);
        out_pairs[2 * i] = a; out_pairs[2 * i + 1] = b;
        out_counts[i] = m.counts[best];
        n = merge_stream(ids, start, ids, start, n, a, b, first_idx + i);
        *n_done = i + 1;
    }
    if (!rc && final_ids) memcpy(final_ids, ids, n * sizeof(int32_t));
    if (final_n) *final_n = n;
    free(ids); free(start); pm_free(&m);
    return rc;
}

/* ------------------------------------------------------------------------------------
 * regex.py:92-121 (_encode_chunk + encode_ordinary) / basic.py:57-74 (one chunk).
 *   bytes/n, offs/n_chunks : text bytes and chunk starts
 *   merges[2*r], merges[2*r+1] : pair of merge rank r (its id is 256 + r)
 *   byte_perm : NULL or a 256-entry byte -> initial id map (gpt4.py:76-77,90-92 style)
 *   out_ids needs room for n ids.
 * ---------------------------------------------------------------------------------- */
int orc_encode(const uint8_t *bytes, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
               const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm,
               int32_t *out_ids, uint64_t *out_n) {
    pairmap ranks; /* pair -> rank, stored as count */
    if (pm_init(&ranks, (uint64_t)n_merges + 16)) return ORC_ENOMEM;
    for (int32_t r = 0; r < n_merges; ++r) {
        uint64_t key = pack(merges[2 * r], merges[2 * r + 1]);
        if (pm_find(&ranks, key) < 0) pm_add(&ranks, key, r); /* dict: later dup keys overwrite
            the value but load() assigns increasing idx; a duplicate line would overwrite with
            the higher idx (base.py:162).  Training never produces duplicates. */
        else ranks.counts[pm_find(&ranks, key)] = r;
    }
    uint64_t one_off = 0;
    if (!offs || n_chunks == 0) { offs = &one_off; n_chunks = n ? 1 : 0; }
    uint64_t w = 0;
    for (uint64_t c = 0; c < n_chunks; ++c) {
        uint64_t lo = offs[c], hi = (c + 1 < n_chunks) ? offs[c + 1] : n;
        int32_t *ids = out_ids + w; /* chunk is encoded in place in the output */
        uint64_t len = hi - lo;
        for (uint64_t k = 0; k < len; ++k)
            ids[k] = byte_perm ? byte_perm[bytes[lo + k]] : bytes[lo + k];
        while (len >= 2) {
            /* min(stats, key=lambda p: merges.get(p, inf)) */
            int64_t best_rank = -1; int32_t a = 0, b = 0;
            for (uint64_t k = 0; k + 1 < len; ++k) {
                int64_t o = pm_find(&ranks, pack(ids[k], ids[k + 1]));
                if (o >= 0 && (best_rank < 0 || ranks.counts[o] < best_rank)) {
                    best_rank = ranks.counts[o]; a = ids[k]; b = ids[k + 1];
                }
            }
            if (best_rank < 0) break; /* "pair not in self.merges" */
            int32_t idx = 256 + (int32_t)best_rank;
            uint64_t i = 0, j = 0;
            while (i < len) {
                if (ids[i] == a && i + 1 < len && ids[i + 1] == b) { ids[j++] = idx; i += 2; }
                else ids[j++] = ids[i++];
            }
            len = j;
        }
        w += len;
    }
    *out_n = w;
    pm_free(&ranks);
    return ORC_OK;
}

/* one timing-friendly step for bench.py's CPU baseline: stats + argmax + merge, in place.
 * ids/start are caller-owned working buffers (start[k]=1 at chunk starts). */
int orc_train_step(int32_t *ids, uint8_t *start, uint64_t *n_io, int32_t idx,
                   int32_t *pair_out, int64_t *count_out) {
    pairmap m;
    if (pm_init(&m, 4096)) return ORC_ENOMEM;
    int rc = stats_into(&m, ids, start, *n_io, NULL);
    if (!rc) {
        if (m.n == 0) rc = 1;
        else {
            uint64_t best = 0;
            for (uint64_t o = 1; o < m.n; ++o) if (m.counts[o] > m.counts[best]) best = o;
            int32_t a = (int32_t)(m.keys[best] >> 32), b = (int32_t)(m.keys[best] & 0xffffffffu);
            pair_out[0] = a; pair_out[1] = b; *count_out = m.counts[best];
            *n_io = merge_stream(ids, start, ids, start, *n_io, a, b, idx);
        }
    }
    pm_free(&m);
    return rc;
}

/* ------------------------------------------------------------------------------------
 * De-duplication of regex chunks (test-side accelerator for orc_train at corpus sizes the plain
 * loop cannot reach; SURVEY.md §8c "Dedup note"): regex.py:51-54 visits the chunks in order and
 * adds every chunk's pairs to ONE dict, so a chunk that occurs w times contributes w to each of its
 * pairs at the position of its FIRST occurrence in the insertion order.  Training on the distinct
 * chunks in first-occurrence order with weight w therefore gives the same dict (keys, order and
 * counts) at every iteration — tests/test_oracle.py::test_dedup_weights_equal_plain pins
 * orc_train(weights) == orc_train(plain) == the reference.
 *   out_bytes / out_offs / out_weights : distinct chunks back to back, their starts, multiplicities
 *   cap_chunks / cap_bytes             : capacities; ORC_ECAP when exceeded
 * ---------------------------------------------------------------------------------- */
static uint64_t hash_bytes(const uint8_t *p, uint64_t len) {
    uint64_t h = 0x9e3779b97f4a7c15ULL ^ (len * 0xff51afd7ed558ccdULL);
    uint64_t k = 0;
    while (len >= 8) { memcpy(&k, p, 8); h = mix64(h ^ k); p += 8; len -= 8; }
    if (len) { k = 0; memcpy(&k, p, len); h = mix64(h ^ k ^ (len << 56)); }
    return h;
}

int orc_dedup_chunks(const uint8_t *bytes, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
                     uint8_t *out_bytes, uint64_t cap_bytes, uint64_t *out_offs, int64_t *out_weights,
                     uint64_t cap_chunks, uint64_t *n_unique, uint64_t *n_out_bytes) {
    uint64_t tcap = 1 << 16, used = 0, wbytes = 0;
    uint64_t *tab = (uint64_t *)calloc(tcap, sizeof(uint64_t));   /* ordinal + 1, 0 = empty */
    uint64_t *hashes = (uint64_t *)malloc(cap_chunks * sizeof(uint64_t));
    if (!tab || !hashes) { free(tab); free(hashes); return ORC_ENOMEM; }
    int rc = ORC_OK;
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const uint64_t lo = offs[c], hi = (c + 1 < n_chunks) ? offs[c + 1] : n, len = hi - lo;
        const uint64_t hv = hash_bytes(bytes + lo, len);
        uint64_t s = hv & (tcap - 1);
        int found = 0;
        for (;;) {
            const uint64_t o = tab[s];
            if (!o) break;
            if (hashes[o - 1] == hv) {
                const uint64_t ulo = out_offs[o - 1], uhi = (o < used) ? out_offs[o] : wbytes;
                if (uhi - ulo == len && memcmp(out_bytes + ulo, bytes + lo, len) == 0) { out_weights[o - 1] += 1; found = 1; break; }
            }
            s = (s + 1) & (tcap - 1);
        }
        if (found) continue;
        if (used == cap_chunks || wbytes + len > cap_bytes) { rc = ORC_ECAP; break; }
        memcpy(out_bytes + wbytes, bytes + lo, len);
        out_offs[used] = wbytes; out_weights[used] = 1; hashes[used] = hv;
        wbytes += len;
        tab[s] = ++used;
        if (used * 2 > tcap) {   /* grow */
            const uint64_t ncap = tcap * 4;
            uint64_t *nt = (uint64_t *)calloc(ncap, sizeof(uint64_t));
            if (!nt) { rc = ORC_ENOMEM; break; }
            for (uint64_t o = 0; o < used; ++o) {
                uint64_t q = hashes[o] & (ncap - 1);
                while (nt[q]) q = (q + 1) & (ncap - 1);
                nt[q] = o + 1;
            }
            free(tab); tab = nt; tcap = ncap;
        }
    }
    *n_unique = used; *n_out_bytes = wbytes;
    free(tab); free(hashes);
    return rc;
}
