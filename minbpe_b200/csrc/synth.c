/*
 * synth.c — deterministic synthetic UTF-8 corpus generator for bench.py and the parity tests
 * (SURVEY.md §8d).  Not part of the BPE path; host-only C so a 1 GiB corpus takes seconds.
 *
 * Text model: a lexicon of 65,536 "words" (length 1..16, mean ~5; 90 % ASCII letters, 4 %
 * digit strings, 3 % two-byte Latin-1/Greek/Cyrillic letters, 2 % three-byte CJK/Hangul,
 * 1 % four-byte emoji) sampled Zipf(s=1.1) with the alias method; separators " " 82 %,
 * "\n" 6 %, ", " 5 %, ". " 4 %, "\n\n" 1 %, "  " / "\t" 1 %, 1 % suffixes ('s 'll 're n't)
 * glued to the previous word, and occasional "====" / "----" / "    " runs (they exercise
 * the equal-pair run-parity rule of merge, base.py:33-40).
 *
 * Determinism: the output is a concatenation of independent 1 MiB blocks, block k seeded
 * from (seed, k), each padded with spaces to exactly 1 MiB, so any number of threads gives
 * the same bytes.  Output is always valid UTF-8 (multi-byte words are never split).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEX 65536
#define MAXW 64 /* bytes per lexicon word: 16 code points * 4 bytes */
#define BLOCK (1u << 20)

typedef struct {
    uint8_t w[LEX][MAXW];
    uint8_t len[LEX];
    double prob[LEX];  /* alias method */
    uint32_t alias[LEX];
} lexicon;

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline uint32_t rnd(uint64_t *s, uint32_t n) { return (uint32_t)((splitmix(s) >> 11) % n); }
static inline double rnd01(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }

static int put_utf8(uint8_t *o, uint32_t cp) {
    if (cp < 0x80) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { o[0] = 0xC0 | (cp >> 6); o[1] = 0x80 | (cp & 63); return 2; }
    if (cp < 0x10000) { o[0] = 0xE0 | (cp >> 12); o[1] = 0x80 | ((cp >> 6) & 63); o[2] = 0x80 | (cp & 63); return 3; }
    o[0] = 0xF0 | (cp >> 18); o[1] = 0x80 | ((cp >> 12) & 63); o[2] = 0x80 | ((cp >> 6) & 63); o[3] = 0x80 | (cp & 63);
    return 4;
}

static void build_lexicon(lexicon *L, uint64_t seed) {
    uint64_t s = seed ^ 0x6c657869636f6eULL;
    static const char letters[] = "etaoinshrdlcumwfgypbvkjxqz"; /* frequency-ish order */
    for (int i = 0; i < LEX; ++i) {
        uint32_t kind = rnd(&s, 100);
        int ncp = 1; /* geometric-ish length, mean ~5, capped at 16 */
        while (ncp < 16 && rnd01(&s) < 0.80) ++ncp;
        int n = 0;
        if (kind < 90) {
            int upper_first = rnd(&s, 10) == 0, all_upper = rnd(&s, 50) == 0;
            for (int k = 0; k < ncp; ++k) {
                /* skew towards frequent letters */
                uint32_t r = rnd(&s, 26), r2 = rnd(&s, 26);
                char c = letters[r < r2 ? r : r2];
                if (all_upper || (upper_first && k == 0)) c = (char)(c - 32);
                L->w[i][n++] = (uint8_t)c;
            }
        } else if (kind < 94) {
            if (ncp > 6) ncp = 1 + ncp % 6;
            for (int k = 0; k < ncp; ++k) L->w[i][n++] = (uint8_t)('0' + rnd(&s, 10));
        } else if (kind < 97) {
            uint32_t script = rnd(&s, 3);
            for (int k = 0; k < ncp; ++k) {
                uint32_t cp = script == 0 ? 0xE0 + rnd(&s, 23)     /* à..ö   */
                            : script == 1 ? 0x3B1 + rnd(&s, 24)    /* Greek  */
                                          : 0x430 + rnd(&s, 32);   /* Cyrillic */
                n += put_utf8(&L->w[i][n], cp);
            }
        } else if (kind < 99) {
            uint32_t script = rnd(&s, 2);
            if (ncp > 8) ncp = 8;
            for (int k = 0; k < ncp; ++k) {
                uint32_t cp = script == 0 ? 0x4E00 + rnd(&s, 2000) : 0xAC00 + rnd(&s, 2000);
                n += put_utf8(&L->w[i][n], cp);
            }
        } else {
            if (ncp > 3) ncp = 1 + ncp % 3;
            for (int k = 0; k < ncp; ++k) n += put_utf8(&L->w[i][n], 0x1F600 + rnd(&s, 64));
        }
        L->len[i] = (uint8_t)n;
    }
    /* Zipf(1.1) weights by rank = index, alias tables (Vose) */
    static double p[LEX];
    static uint32_t small[LEX], large[LEX];
    double z = 0;
    for (int i = 0; i < LEX; ++i) { p[i] = 1.0 / pow((double)(i + 1), 1.1); z += p[i]; }
    int ns = 0, nl = 0;
    for (int i = 0; i < LEX; ++i) { p[i] = p[i] / z * LEX; if (p[i] < 1.0) small[ns++] = i; else large[nl++] = i; }
    while (ns && nl) {
        uint32_t a = small[--ns], g = large[--nl];
        L->prob[a] = p[a]; L->alias[a] = g;
        p[g] = (p[g] + p[a]) - 1.0;
        if (p[g] < 1.0) small[ns++] = g; else large[nl++] = g;
    }
    while (nl) { uint32_t g = large[--nl]; L->prob[g] = 1.0; L->alias[g] = g; }
    while (ns) { uint32_t a = small[--ns]; L->prob[a] = 1.0; L->alias[a] = a; }
}

static void gen_block(const lexicon *L, uint64_t seed, uint64_t k, uint8_t *out) {
    uint64_t s = seed * 0x100000001b3ULL + k * 0x9e3779b97f4a7c15ULL + 12345;
    splitmix(&s);
    uint32_t n = 0;
    for (;;) {
        uint8_t tmp[MAXW + 32];
        uint32_t t = 0;
        uint32_t col = rnd(&s, LEX);
        uint32_t wi = rnd01(&s) < L->prob[col] ? col : L->alias[col];
        memcpy(tmp, L->w[wi], L->len[wi]); t = L->len[wi];
        uint32_t r = rnd(&s, 1000);
        if (r < 10) { /* contraction suffix */
            static const char *suf[] = {"'s", "'ll", "'re", "n't", "'S", "'ve"};
            const char *q = suf[rnd(&s, 6)];
            size_t l = strlen(q); memcpy(tmp + t, q, l); t += (uint32_t)l;
        }
        r = rnd(&s, 1000);
        if (r < 820) tmp[t++] = ' ';
        else if (r < 880) tmp[t++] = '\n';
        else if (r < 930) { tmp[t++] = ','; tmp[t++] = ' '; }
        else if (r < 970) { tmp[t++] = '.'; tmp[t++] = ' '; }
        else if (r < 980) { tmp[t++] = '\n'; tmp[t++] = '\n'; }
        else if (r < 985) { tmp[t++] = ' '; tmp[t++] = ' '; }
        else if (r < 990) tmp[t++] = '\t';
        else { /* a run of one repeated symbol, length 2..13, then a newline */
            static const char sym[] = "=-* ._#";
            char c = sym[rnd(&s, 7)];
            uint32_t l = 2 + rnd(&s, 12);
            tmp[t++] = ' ';
            for (uint32_t q = 0; q < l; ++q) tmp[t++] = (uint8_t)c;
            tmp[t++] = '\n';
        }
        if (n + t > BLOCK) break;
        memcpy(out + n, tmp, t); n += t;
    }
    memset(out + n, ' ', BLOCK - n);
}

typedef struct { const lexicon *L; uint64_t seed, k0, k1, kbase; uint8_t *out; uint64_t nbytes; } job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    uint8_t *buf = (uint8_t *)malloc(BLOCK);
    for (uint64_t k = j->k0; k < j->k1; ++k) {
        uint64_t off = k * (uint64_t)BLOCK;
        if (off + BLOCK <= j->nbytes) gen_block(j->L, j->seed, j->kbase + k, j->out + off);
        else { /* final partial block: generate whole, copy a prefix cut at a char boundary */
            gen_block(j->L, j->seed, j->kbase + k, buf);
            uint64_t m = j->nbytes - off;
            uint64_t cut = m;
            while (cut > 0 && (buf[cut] & 0xC0) == 0x80) --cut; /* buf[cut] would start mid-char */
            memcpy(j->out + off, buf, cut);
            memset(j->out + off + cut, ' ', m - cut);
        }
    }
    free(buf);
    return NULL;
}

/* Fill out[0..nbytes) with the 1 MiB blocks first_block, first_block+1, ... of the corpus for `seed`
 * (a contiguous shard of that corpus: same lexicon, same text as a full generation would put there).
 * Returns 0 on success. */
int bpe_synth_generate_at(uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t nbytes, int n_threads) {
    lexicon *L = (lexicon *)malloc(sizeof(lexicon));
    if (!L) return -1;
    build_lexicon(L, seed);
    uint64_t nblocks = (nbytes + BLOCK - 1) / BLOCK;
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > nblocks) n_threads = (int)(nblocks ? nblocks : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    job *jobs = (job *)malloc(sizeof(job) * n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (job){L, seed, nblocks * t / n_threads, nblocks * (t + 1) / n_threads, first_block, out, nbytes};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs); free(L);
    return 0;
}

/* Fill out[0..nbytes) with the corpus for `seed` from its beginning. */
int bpe_synth_generate(uint64_t seed, uint8_t *out, uint64_t nbytes, int n_threads) {
    return bpe_synth_generate_at(seed, 0, out, nbytes, n_threads);
}
