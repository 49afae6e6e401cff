// k_split.cuh — the GPT-4 split pattern (regex.py:19, applied by regex.py:41 / :114) on the device.
//
//   '(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+
//
// A regex matcher is sequential, but for THIS pattern "does a match start at character i?" is a
// function of the local run structure (runs of letters / digits / whitespace / other, their
// starts, ends and a few neighbouring characters).  The rules are derived in DESIGN.md §9 and
// pinned against the `regex` module on 130k adversarial strings by tests/test_split_rules.py
// (numpy restatement: oracle/split_rules.py).  Here they run as
//   k_split_classify   byte -> {class, char-start}            (code-point class table from `regex`)
//   forward scan       run start (max), last newline (max), char count (sum)
//   backward scan      run end (min), next non-newline (min)
//   k_split_rules      per character: chunk start?  -> 1 flag byte
// followed either by "widen + mark" straight into the token stream (training needs no offsets at
// all) or by a flag compaction into chunk offsets (encode / host callers).
// All scans are 3-kernel tile scans (reduce, scan of tile sums by one block, down-sweep).
#pragma once
#include "common.cuh"

#define SC_L 0u
#define SC_N 1u
#define SC_NL 2u
#define SC_SP 3u
#define SC_AP 4u
#define SC_O 5u
#define SM_CLS 7u
#define SM_START 8u     // first byte of a UTF-8 sequence
#define SP_TILE 2048
#define SP_THREADS 256
#define SP_ITEMS (SP_TILE / SP_THREADS)
#define IDX_INF 0xffffffffu

// run kind: letters, digits, whitespace (SP or NL), other (O or apostrophe)
__device__ __forceinline__ u32 kind_of(u32 cls) { return cls == SC_NL ? SC_SP : (cls == SC_AP ? SC_O : cls); }

__device__ __forceinline__ u32 utf8_decode(const unsigned char *__restrict__ b, u64 i, u64 n) {
    const u32 c0 = b[i];
    if (c0 < 0x80) return c0;
    if (c0 < 0xE0) return ((c0 & 0x1f) << 6) | ((i + 1 < n ? b[i + 1] : 0) & 0x3f);
    if (c0 < 0xF0) return ((c0 & 0x0f) << 12) | (((i + 1 < n ? b[i + 1] : 0) & 0x3f) << 6) | ((i + 2 < n ? b[i + 2] : 0) & 0x3f);
    return ((c0 & 0x07) << 18) | (((i + 1 < n ? b[i + 1] : 0) & 0x3f) << 12) | (((i + 2 < n ? b[i + 2] : 0) & 0x3f) << 6) |
           ((i + 3 < n ? b[i + 3] : 0) & 0x3f);
}

// first byte of the character that byte i belongs to
__device__ __forceinline__ u64 char_start_of(const unsigned char *__restrict__ b, u64 i) {
    while (i > 0 && (b[i] & 0xC0) == 0x80) --i;
    return i;
}

// meta[i] = class of the character byte i belongs to | SM_START on its first byte
__global__ void __launch_bounds__(256) k_split_classify(const unsigned char *__restrict__ b, u64 n,
                                                        const unsigned char *__restrict__ cls_table,
                                                        unsigned char *__restrict__ meta) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const bool start = (b[i] & 0xC0) != 0x80;
        const u64 s = start ? i : char_start_of(b, i);
        u32 cp = utf8_decode(b, s, n);
        if (cp > 0x10ffffu) cp = 0xfffd;
        meta[i] = (unsigned char)(cls_table[cp] | (start ? SM_START : 0u));
    }
}

// ---- forward scan: run start (max), last newline + 1 (max), number of character starts (sum) ----
struct Fwd { u32 rs, nl, cnt; };
__device__ __forceinline__ Fwd fwd_op(Fwd a, Fwd b) { Fwd r; r.rs = max(a.rs, b.rs); r.nl = max(a.nl, b.nl); r.cnt = a.cnt + b.cnt; return r; }
__device__ __forceinline__ Fwd fwd_elem(const unsigned char *__restrict__ meta, u64 i) {
    const u32 m = meta[i];
    Fwd v;
    v.rs = (i == 0 || kind_of(meta[i - 1] & SM_CLS) != kind_of(m & SM_CLS)) ? (u32)i : 0u;
    v.nl = ((m & SM_CLS) == SC_NL) ? (u32)i + 1u : 0u;
    v.cnt = (m & SM_START) ? 1u : 0u;
    return v;
}
// ---- backward scan: run end, one past (min), next non-newline byte (min) ----
struct Bwd { u32 re, nnl; };
__device__ __forceinline__ Bwd bwd_op(Bwd a, Bwd b) { Bwd r; r.re = min(a.re, b.re); r.nnl = min(a.nnl, b.nnl); return r; }
__device__ __forceinline__ Bwd bwd_elem(const unsigned char *__restrict__ meta, u64 i, u64 n) {
    const u32 m = meta[i];
    Bwd v;
    v.re = (i + 1 == n || kind_of(meta[i + 1] & SM_CLS) != kind_of(m & SM_CLS)) ? (u32)i + 1u : IDX_INF;
    v.nnl = ((m & SM_CLS) != SC_NL) ? (u32)i : IDX_INF;
    return v;
}

template <typename T, typename OP>
__device__ __forceinline__ T block_reduce(T v, OP op, T *sm) {
    const u32 tid = threadIdx.x;
    sm[tid] = v;
    __syncthreads();
    for (int o = SP_THREADS / 2; o > 0; o >>= 1) { if (tid < (u32)o) sm[tid] = op(sm[tid], sm[tid + o]); __syncthreads(); }
    const T r = sm[0];
    __syncthreads();
    return r;
}

// tile aggregates (forward tiles in ascending order; the backward scan reads the same tiles right to left)
__global__ void __launch_bounds__(SP_THREADS) k_split_reduce(const unsigned char *__restrict__ meta, u64 n,
                                                             Fwd *__restrict__ fpart, Bwd *__restrict__ bpart) {
    __shared__ Fwd sf[SP_THREADS];
    __shared__ Bwd sb[SP_THREADS];
    const u64 base = (u64)blockIdx.x * SP_TILE + (u64)threadIdx.x * SP_ITEMS;
    Fwd f = {0u, 0u, 0u};
    Bwd g = {IDX_INF, IDX_INF};
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) {
        const u64 i = base + k;
        if (i < n) { f = fwd_op(f, fwd_elem(meta, i)); g = bwd_op(g, bwd_elem(meta, i, n)); }
    }
    f = block_reduce(f, fwd_op, sf);
    g = block_reduce(g, bwd_op, sb);
    if (threadIdx.x == 0) { fpart[blockIdx.x] = f; bpart[blockIdx.x] = g; }
}

// exclusive scans of the tile aggregates (one block): forward left-to-right, backward right-to-left
__global__ void __launch_bounds__(1024) k_split_scan_parts(Fwd *__restrict__ fpart, Bwd *__restrict__ bpart, u32 ntiles) {
    __shared__ Fwd sf[1024];
    __shared__ Bwd sb[1024];
    const u32 tid = threadIdx.x;
    const u32 per = (ntiles + 1023) / 1024;
    const u32 lo = min(ntiles, tid * per), hi = min(ntiles, lo + per);
    Fwd f = {0u, 0u, 0u};
    for (u32 t = lo; t < hi; ++t) f = fwd_op(f, fpart[t]);
    Bwd g = {IDX_INF, IDX_INF};
    for (u32 t = lo; t < hi; ++t) g = bwd_op(g, bpart[t]);
    sf[tid] = f; sb[tid] = g;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {   // inclusive: forward over lower tids, backward over higher tids
        Fwd fv = {0u, 0u, 0u}; Bwd gv = {IDX_INF, IDX_INF};
        if (tid >= (u32)o) fv = sf[tid - o];
        if (tid + o < 1024) gv = sb[tid + o];
        __syncthreads();
        sf[tid] = fwd_op(sf[tid], fv); sb[tid] = bwd_op(sb[tid], gv);
        __syncthreads();
    }
    Fwd frun = {0u, 0u, 0u};
    if (tid > 0) frun = sf[tid - 1];
    for (u32 t = lo; t < hi; ++t) { const Fwd x = fpart[t]; fpart[t] = frun; frun = fwd_op(frun, x); }
    Bwd grun = {IDX_INF, IDX_INF};
    if (tid + 1 < 1024) grun = sb[tid + 1];
    for (u32 t = hi; t > lo; --t) { const Bwd x = bpart[t - 1]; bpart[t - 1] = grun; grun = bwd_op(grun, x); }
}

// down-sweep: inclusive scan values for every byte
__global__ void __launch_bounds__(SP_THREADS) k_split_down(const unsigned char *__restrict__ meta, u64 n,
                                                           const Fwd *__restrict__ fpart, const Bwd *__restrict__ bpart,
                                                           u32 *__restrict__ o_rs, u32 *__restrict__ o_nl, u32 *__restrict__ o_cnt,
                                                           u32 *__restrict__ o_re, u32 *__restrict__ o_nnl) {
    __shared__ Fwd sf[SP_THREADS];
    __shared__ Bwd sb[SP_THREADS];
    const u32 tid = threadIdx.x;
    const u64 base = (u64)blockIdx.x * SP_TILE + (u64)tid * SP_ITEMS;
    Fwd fe[SP_ITEMS]; Bwd ge[SP_ITEMS];
    Fwd f = {0u, 0u, 0u};
    Bwd g = {IDX_INF, IDX_INF};
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) {
        const u64 i = base + k;
        if (i < n) { fe[k] = fwd_elem(meta, i); ge[k] = bwd_elem(meta, i, n); }
        else { fe[k] = {0u, 0u, 0u}; ge[k] = {IDX_INF, IDX_INF}; }
        f = fwd_op(f, fe[k]);
    }
#pragma unroll
    for (int k = SP_ITEMS - 1; k >= 0; --k) g = bwd_op(g, ge[k]);
    sf[tid] = f; sb[tid] = g;
    __syncthreads();
    for (int o = 1; o < SP_THREADS; o <<= 1) {
        Fwd fv = {0u, 0u, 0u}; Bwd gv = {IDX_INF, IDX_INF};
        if (tid >= (u32)o) fv = sf[tid - o];
        if (tid + o < SP_THREADS) gv = sb[tid + o];
        __syncthreads();
        sf[tid] = fwd_op(sf[tid], fv); sb[tid] = bwd_op(sb[tid], gv);
        __syncthreads();
    }
    Fwd frun = fpart[blockIdx.x];
    if (tid > 0) frun = fwd_op(frun, sf[tid - 1]);
    Bwd grun = bpart[blockIdx.x];
    if (tid + 1 < SP_THREADS) grun = bwd_op(grun, sb[tid + 1]);
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) {
        const u64 i = base + k;
        frun = fwd_op(frun, fe[k]);
        if (i < n) { o_rs[i] = frun.rs; o_nl[i] = frun.nl; o_cnt[i] = frun.cnt; }
    }
#pragma unroll
    for (int k = SP_ITEMS - 1; k >= 0; --k) {
        const u64 i = base + k;
        grun = bwd_op(grun, ge[k]);
        if (i < n) { o_re[i] = grun.re; o_nnl[i] = grun.nnl; }
    }
}

// contraction sets: bit 0 = matches (?i:[sdmt]), 1 = (?i:l), 2 = (?i:v), 3 = (?i:e), 4 = (?i:r)   (code points < 0x3000)
__device__ __forceinline__ u32 contr_bits(const unsigned char *__restrict__ contr, u32 cp) { return cp < 0x3000u ? contr[cp] : 0u; }

// length in characters (2 or 3) of a contraction whose apostrophe is followed by the characters at byte s, or 0
__device__ __forceinline__ u32 contraction_len(const unsigned char *__restrict__ b, u64 n, const unsigned char *__restrict__ contr, u64 s) {
    if (s >= n) return 0;
    const u32 c1 = utf8_decode(b, s, n);
    const u32 b1 = contr_bits(contr, c1);
    if (b1 & 1u) return 2;
    const u64 s2 = s + (c1 < 0x80 ? 1 : c1 < 0x800 ? 2 : c1 < 0x10000 ? 3 : 4);
    if (s2 >= n) return 0;
    const u32 b2 = contr_bits(contr, utf8_decode(b, s2, n));
    if (((b1 & 2u) && (b2 & 2u)) || ((b1 & 4u) && (b2 & 8u)) || ((b1 & 16u) && (b2 & 8u))) return 3;
    return 0;
}

// Is the apostrophe-free statement "a match starts at the single Oish character at byte p" true?
// (p is a one-character Oish run and no U+0020 precedes it)
__device__ __forceinline__ bool single_oish_start(const unsigned char *__restrict__ b, const unsigned char *__restrict__ meta,
                                                  const u32 *__restrict__ rs, const u32 *__restrict__ re, u64 p, u64 next) {
    if (kind_of(meta[p] & SM_CLS) != SC_O) return false;
    if (rs[p] != (u32)p || re[p] != (u32)next) return false;
    return p == 0 || b[p - 1] != 0x20;
}

// flag[i] = 1 iff a chunk starts at byte i (always a character start).  Rules: oracle/split_rules.py.
__global__ void __launch_bounds__(256) k_split_rules(const unsigned char *__restrict__ b, u64 n,
                                                     const unsigned char *__restrict__ meta, const unsigned char *__restrict__ contr,
                                                     const u32 *__restrict__ rs, const u32 *__restrict__ nl, const u32 *__restrict__ cnt,
                                                     const u32 *__restrict__ re, const u32 *__restrict__ nnl,
                                                     unsigned char *__restrict__ flag) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 m = meta[i];
        bool st = false;
        if (m & SM_START) {
            const u32 cls = m & SM_CLS, kind = kind_of(cls);
            const u64 s = rs[i], e = re[i];
            if (i == 0) st = true;
            else if (kind == SC_N) st = ((cnt[i] - cnt[s]) % 3u) == 0u;
            else if (kind == SC_O) st = (i == s) && (b[i - 1] != 0x20);
            else if (kind == SC_SP) {
                // whitespace run [s, e): leading newlines belong to a preceding Oish chunk
                const bool prev_oish = s > 0 && kind_of(meta[s - 1] & SM_CLS) == SC_O;
                const u64 w2s = prev_oish ? min((u64)nnl[s], e) : s;
                const u32 lnl1 = nl[e - 1];                       // last newline at or before e-1, plus one (0 = none)
                const bool has_nl = lnl1 != 0 && (u64)(lnl1 - 1) >= w2s && w2s < e;
                const u64 w3s = has_nl ? (u64)lnl1 : w2s;         // first byte after the last newline
                const u32 k = w3s < e ? cnt[e - 1] - cnt[w3s] + 1u : 0u;   // characters in the trailing spaces
                const u64 last = char_start_of(b, e - 1);
                st = (i == w2s && has_nl) ||
                     (i == w3s && k >= 1 && (e == n || k >= 2)) ||
                     (i == last && e < n && k >= 1);
            } else {   // letters
                if (i == s) {
                    const u64 p = char_start_of(b, i - 1);
                    const u32 pc = meta[p] & SM_CLS;
                    const bool absorbed = (pc == SC_SP) || single_oish_start(b, meta, rs, re, p, i);
                    st = !absorbed;
                } else {
                    // second or third letter of a run that follows a contraction apostrophe: "'s|foo", "'ll|ama"
                    const u32 d = cnt[i] - cnt[s];
                    if (d <= 2 && s > 0 && b[s - 1] == 0x27 && single_oish_start(b, meta, rs, re, s - 1, s)) {
                        const u32 clen = contraction_len(b, n, contr, s);
                        st = clen != 0 && d == clen - 1;
                    }
                }
            }
        }
        flag[i] = st ? 1 : 0;
    }
}

// text bytes -> token words with the chunk marks taken straight from the split flags
__global__ void __launch_bounds__(256) k_widen_marked(const unsigned char *__restrict__ src, const unsigned char *__restrict__ flag,
                                                      u32 *__restrict__ dst, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        dst[i] = (u32)src[i] | (flag[i] ? TOK_FLAG : 0u);
}

// ---- flag compaction: chunk offsets (3-kernel sum scan over the same tiles) ----
__global__ void __launch_bounds__(SP_THREADS) k_flag_reduce(const unsigned char *__restrict__ flag, u64 n, u32 *__restrict__ part) {
    __shared__ u32 sm[SP_THREADS];
    const u64 base = (u64)blockIdx.x * SP_TILE + (u64)threadIdx.x * SP_ITEMS;
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) if (base + k < n) c += flag[base + k];
    sm[threadIdx.x] = c;
    __syncthreads();
    for (int o = SP_THREADS / 2; o > 0; o >>= 1) { if (threadIdx.x < (u32)o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(1024) k_flag_scan_parts(const u32 *__restrict__ part, u64 *__restrict__ excl, u32 ntiles, u64 *total) {
    __shared__ u64 sm[1024];
    const u32 tid = threadIdx.x;
    const u32 per = (ntiles + 1023) / 1024;
    const u32 lo = min(ntiles, tid * per), hi = min(ntiles, lo + per);
    u64 sum = 0;
    for (u32 t = lo; t < hi; ++t) sum += part[t];
    sm[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { const u64 v = tid >= (u32)o ? sm[tid - o] : 0; __syncthreads(); sm[tid] += v; __syncthreads(); }
    u64 run = sm[tid] - sum;
    for (u32 t = lo; t < hi; ++t) { excl[t] = run; run += part[t]; }
    if (tid == 1023) *total = sm[1023];
}

__global__ void __launch_bounds__(SP_THREADS) k_flag_scatter(const unsigned char *__restrict__ flag, u64 n, const u64 *__restrict__ excl,
                                                             u64 *__restrict__ offs) {
    __shared__ u32 sm[SP_THREADS];
    const u32 tid = threadIdx.x;
    const u64 base = (u64)blockIdx.x * SP_TILE + (u64)tid * SP_ITEMS;
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) if (base + k < n) c += flag[base + k];
    sm[tid] = c;
    __syncthreads();
    for (int o = 1; o < SP_THREADS; o <<= 1) { const u32 v = tid >= (u32)o ? sm[tid - o] : 0; __syncthreads(); sm[tid] += v; __syncthreads(); }
    u64 dst = excl[blockIdx.x] + (sm[tid] - c);
#pragma unroll
    for (int k = 0; k < SP_ITEMS; ++k) if (base + k < n && flag[base + k]) offs[dst++] = base + k;
}
