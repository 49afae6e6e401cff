// k_merge_seg.cuh — the fused merge pass (base.py:25-41 + the statistics delta) for a != b on the
// SEGMENTED stream: every 4096-word segment is compacted in place by one CTA, independently of
// all other segments.  There is no prefix sum across segments, hence no serial dependency
// between CTAs: the kernel is a pure stream pass.
//
//   read   4 * count bytes per segment        (one 1-D bulk async copy, TMA / UBLKCP)
//   write  4 * new_count bytes, ONLY if the segment changed   (coalesced stores, same address range)
//   + 32-byte edge records per segment (first 3 / last 2 tokens, count), double-buffered, so a
//     CTA can see across its segment boundaries without reading a body that another CTA rewrites
//
// Warp roles: warp 8 lane 0 = producer (segment tickets, neighbour tokens, bulk copies through a
// 2-stage mbarrier ring, fully decoupled from the consumers through full/empty barriers);
// warps 0..7 = consumers (mark, warp-local scan, scatter to the staging tile, delta, copy-out).
// The marking and delta rules are the ones documented in k_merge.cuh.
#pragma once
#include "common.cuh"
#include "k_merge.cuh"
#include "k_seg.cuh"

#define MS_CWARPS 8
#define MS_CTHREADS (MS_CWARPS * 32)
#define MS_THREADS (MS_CTHREADS + 32)
#define MS_WSPAN (SEG_TOKENS / MS_CWARPS)   // 512 tokens per consumer warp
// Tuned on B200 (profiles/r1_summary.md §5): 2 stages x 16 KB + 16 KB staging + 4 KB delta cache =
// 53 KB -> 4 CTAs/SM at 56 registers (36 warps) beat 3 stages / 3 CTAs by 8 %.
#ifndef MS_STAGES
#define MS_STAGES 3
#endif
#ifndef MS_MINBLOCKS
#define MS_MINBLOCKS 4
#endif
#ifndef MS_DCACHE_LOG2
#define MS_DCACHE_LOG2 9
#endif
#define MS_PAD 4                            // body starts at word 4 of a stage (16-byte aligned)
#define MS_IN_WORDS (SEG_TOKENS + 8)
#define MS_DCACHE (1 << MS_DCACHE_LOG2)     // slots of the per-CTA delta cache (shared memory)
#define MS_SMEM_BYTES (MS_STAGES * MS_IN_WORDS * 4 + MS_DCACHE * 8 + 640)
#define MS_INVALID 0xffffffffu


// token i of a segment extended by its neighbours: i in [-2, count+3); h = {P0,P1,N0,N1,N2}
__device__ __noinline__ u32 seg_tok(const u32 *s, const u32 *h, u32 count, int i) {
    if (i < 0) return (i >= -2) ? h[-i - 1] : TOK_SENTINEL;
    if ((u32)i >= count) return ((u32)i - count < 3u) ? h[2 + (u32)i - count] : TOK_SENTINEL;
    return s[i];
}

// delta[idx] += 1 through a CTA-private shared-memory cache: the same few neighbour ids are hit by
// almost every merge of a dense iteration (global same-address atomics serialise in L2); the
// persistent CTA folds them here and flushes once at exit.
__device__ __noinline__ void delta_cache_add(u32 *s_dkey, u32 *s_dcnt, ull *delta, u32 idx) {
    u32 slot = (idx * 2654435761u) >> (32 - MS_DCACHE_LOG2);
#pragma unroll 1
    for (int probe = 0; probe < 4; ++probe) {
        u32 k = reinterpret_cast<volatile u32 *>(s_dkey)[slot];
        if (k == 0xffffffffu) {
            const u32 old = atomicCAS(&s_dkey[slot], 0xffffffffu, idx);
            k = (old == 0xffffffffu) ? idx : old;
        }
        if (k == idx) { atomicAdd(&s_dcnt[slot], 1u); return; }
        slot = (slot + 1) & (MS_DCACHE - 1);
    }
    atomicAdd(&delta[idx], 1ull);   // cache neighbourhood full
}

// statistics delta of the merge that starts at token i of the segment (rules: k_merge.cuh)
__device__ __noinline__ void delta_one(const u32 *s, const u32 *h, u32 count, int i, u32 a, u32 b, u32 V,
                                       u32 *s_dkey, u32 *s_dcnt, ull *delta) {
    const u32 t0 = s[i];
    u32 tm1, tm2, tp2, tp3;
    if (i >= 2 && (u32)i + 3 < count) {   // the usual case: all four neighbours inside the segment
        tm1 = s[i - 1]; tm2 = s[i - 2]; tp2 = s[i + 2]; tp3 = s[i + 3];
    } else {
        tm1 = seg_tok(s, h, count, i - 1); tm2 = seg_tok(s, h, count, i - 2);
        tp2 = seg_tok(s, h, count, i + 2); tp3 = seg_tok(s, h, count, i + 3);
    }
    const bool m_m2 = (((tm2 ^ a) & TOK_MASK) == 0) && tm1 == b;   // a merge starts at i-2
    const bool m_p2 = (((tp2 ^ a) & TOK_MASK) == 0) && tp3 == b;   // a merge starts at i+2
    if (tm1 != TOK_SENTINEL && !(t0 & TOK_FLAG) && !m_m2) delta_cache_add(s_dkey, s_dcnt, delta, tm1 & TOK_MASK);
    if (!(tp2 & TOK_FLAG))   // also false for the sentinel (end of stream)
        delta_cache_add(s_dkey, s_dcnt, delta, m_p2 ? 2u * V : V + tp2);
}

struct SegArgs {
    Ctl *ctl;
    u32 *buf0, *buf1;
    Edge *e0, *e1;
    ull *delta;   // [0,V) L, [V,2V) R, [2V] ZZ; NULL = plain merge
    u32 V;
    int force;
};

__global__ void __launch_bounds__(MS_THREADS, MS_MINBLOCKS) k_merge_seg(SegArgs A) {
    Ctl *ctl = A.ctl;
    if (!A.force && (ctl->done || ctl->overflow || ctl->iter >= ctl->max_iter)) return;
    if (ctl->a == ctl->b) return;  // pairs (a,a) take the pack + k_merge<true> path

    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32 *s_in = reinterpret_cast<u32 *>(smem_raw);                    // [MS_STAGES][MS_IN_WORDS]
    u64 *s_full = reinterpret_cast<u64 *>(s_in + MS_STAGES * MS_IN_WORDS);   // [MS_STAGES]
    u64 *s_empty = s_full + MS_STAGES;                                // [MS_STAGES]
    u32 *s_seg = reinterpret_cast<u32 *>(s_empty + MS_STAGES);        // [MS_STAGES] segment id or MS_INVALID
    u32 *s_cnt = s_seg + MS_STAGES;                                   // [MS_STAGES]
    u32 *s_halo = s_cnt + MS_STAGES;                                  // [MS_STAGES][8]: P0 P1 N0 N1 N2
    u32 *s_wtot2 = s_halo + MS_STAGES * 8;                            // [2][MS_CWARPS], by tile parity
    u32 *s_edge = s_wtot2 + 2 * MS_CWARPS;                            // [2][8] boundary tokens of the segment being written
    u32 *s_dkey = s_edge + 16;                            // [MS_DCACHE] delta index or 0xffffffff
    u32 *s_dcnt = s_dkey + MS_DCACHE;                                 // [MS_DCACHE]

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_ctrl = (warp == MS_CWARPS);
    const u64 n = ctl->n;
    u32 *__restrict__ w = ctl->cur ? A.buf1 : A.buf0;                 // compacted in place
    const Edge *__restrict__ e_cur = ctl->edge_cur ? A.e1 : A.e0;
    Edge *__restrict__ e_next = ctl->edge_cur ? A.e0 : A.e1;
    const u32 a = (u32)ctl->a, b = (u32)ctl->b, z = (u32)ctl->z;
    const u32 nseg = ctl->nseg;

    if (tid == 0) {
        for (int s = 0; s < MS_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], MS_CWARPS); }
        fence_mbar_init();
    }
    for (u32 i = tid; i < MS_DCACHE; i += MS_THREADS) { s_dkey[i] = 0xffffffffu; s_dcnt[i] = 0; }
    __syncthreads();

    if (is_ctrl) {
        // ================= producer: one thread keeps the 3-stage ring full =================
        if (lane == 0) {
            for (u32 j = 0;; ++j) {
                const u32 stage = j % MS_STAGES;
                if (j >= MS_STAGES) mbar_wait(&s_empty[stage], ((j / MS_STAGES) - 1) & 1u);
                u32 seg, cnt = 0;
                for (;;) {  // next non-empty segment; empty ones only need their (empty) edge carried over
                    seg = atomicAdd(&ctl->merge_ticket, 1u);
                    if (seg >= nseg) break;
                    cnt = e_cur[seg].count;
                    if (cnt) break;
                    Edge ed;
                    edge_from_tokens(ed, nullptr, 0);
                    e_next[seg] = ed;
                }
                if (seg >= nseg) { s_seg[stage] = MS_INVALID; mbar_arrive(&s_full[stage]); break; }
                u32 N[3], P[2];
                seg_neighbours(e_cur, seg, nseg, N, P);
                u32 *h = s_halo + stage * 8;
                h[0] = P[0]; h[1] = P[1]; h[2] = N[0]; h[3] = N[1]; h[4] = N[2];
                s_in[stage * MS_IN_WORDS + MS_PAD - 1] = P[0];   // s[-1], s[-2] for the consumers' unconditional reads
                s_in[stage * MS_IN_WORDS + MS_PAD - 2] = P[1];
                s_seg[stage] = seg; s_cnt[stage] = cnt;
                const u32 bytes = ((cnt + 3u) & ~3u) * 4u;
                mbar_arrive_expect_tx(&s_full[stage], bytes);   // release: the stores above are visible to waiters
                bulk_g2s(s_in + stage * MS_IN_WORDS + MS_PAD, w + (u64)seg * SEG_TOKENS, bytes, &s_full[stage]);
            }
        }
    } else {
        // ================= consumers =================
        ull cta_drops = 0;  // meaningful in thread 0
        // thread 0: the edge record of the previous changed segment is assembled one barrier later,
        // when every warp has finished its copy-out and deposited the boundary tokens in s_edge
        bool pending = false; u32 pend_par = 0;   // thread 0 only
        auto flush_edge = [&]() {   // called by warp 0 right after a block barrier
            const u32 par = __shfl_sync(0xffffffffu, pending ? pend_par + 1u : 0u, 0);
            if (!par) return;
            const u32 *se = s_edge + (par - 1u) * 8;
            const u32 cnt = se[5], sg = se[6];
            if (lane < 8) {
                u32 word = 0;
                if (lane < 3) word = (lane < cnt) ? se[lane] : TOK_SENTINEL;
                else if (lane == 3) word = cnt >= 2 ? se[3] : TOK_SENTINEL;
                else if (lane == 4) word = cnt >= 1 ? se[4] : TOK_SENTINEL;
                else if (lane == 5) word = cnt;
                reinterpret_cast<u32 *>(&e_next[sg])[lane] = word;   // Edge = f[3], l[2], count, pad[2]
            }
            pending = false;
        };
        for (u32 j = 0;; ++j) {
            const u32 stage = j % MS_STAGES;
            mbar_wait(&s_full[stage], (j / MS_STAGES) & 1u);
            const u32 seg = s_seg[stage];
            if (seg == MS_INVALID) break;
            const u32 count = s_cnt[stage];
            const u32 *s = s_in + stage * MS_IN_WORDS + MS_PAD;   // s[i] = token i of the segment
            const u32 *h = s_halo + stage * 8;
            u32 *s_wtot = s_wtot2 + (j & 1u) * MS_CWARPS;   // parity: an untouched segment skips barrier (2)
            // token i of the segment extended by its neighbours (i in [-2, count+3))
            auto tok = [&](int i) -> u32 { return seg_tok(s, h, count, i); };

            u32 t[4][4], mn[4], keep[4], lpre[4], rowoff[4], rowcnt[4];
            u32 wtot = 0, many = 0;
            bool plain = true;   // warp-uniform: no token of this warp's span is removed or replaced
            const u32 wbase = warp * MS_WSPAN;
            // interior warp: its 512 tokens and the one after them all belong to this segment -> no
            // bounds logic; neighbours come from shared memory (s[-1], s[-2] are filled by the producer)
            const bool interior = (wbase + MS_WSPAN + 4 <= count);
            u32 rem_any = 0;   // some token of the lane is dropped or lies past the end of the segment
            if (interior) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int li = (int)(wbase + r * 128 + lane * 4);
                    const uint4 q = *reinterpret_cast<const uint4 *>(s + li);
                    t[r][0] = q.x; t[r][1] = q.y; t[r][2] = q.z; t[r][3] = q.w;
                    const u32 nx = s[li + 4], pv = s[li - 1];
                    u32 m = 0;
                    m |= (((t[r][0] ^ a) & TOK_MASK) == 0 && t[r][1] == b) ? 1u : 0u;
                    m |= (((t[r][1] ^ a) & TOK_MASK) == 0 && t[r][2] == b) ? 2u : 0u;
                    m |= (((t[r][2] ^ a) & TOK_MASK) == 0 && t[r][3] == b) ? 4u : 0u;
                    m |= (((t[r][3] ^ a) & TOK_MASK) == 0 && nx == b) ? 8u : 0u;
                    const u32 pm = (((pv ^ a) & TOK_MASK) == 0 && t[r][0] == b) ? 1u : 0u;
                    const u32 d = ((m << 1) | pm) & 0xfu;
                    mn[r] = m;
                    keep[r] = d ^ 0xfu;
                    rem_any |= d;
                    many |= m;
                }
            } else if (wbase < count) {
                u32 nxt[4], pbit[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 li = wbase + r * 128 + lane * 4;
                    if (wbase + r * 128 < count) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(s + li);
                        t[r][0] = q.x; t[r][1] = q.y; t[r][2] = q.z; t[r][3] = q.w;
                        if (li + 4 > count) {  // lanes at / past the end see the following segment's tokens
#pragma unroll
                            for (int k = 0; k < 4; ++k) if (li + k >= count) t[r][k] = tok((int)(li + k));
                        }
                    } else {
                        t[r][0] = t[r][1] = t[r][2] = t[r][3] = TOK_SENTINEL;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 li = wbase + r * 128 + lane * 4;
                    u32 v = __shfl_down_sync(0xffffffffu, t[r][0], 1);
                    if (lane == 31) v = (li + 4 < count) ? s[li + 4] : tok((int)(li + 4));
                    nxt[r] = v;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    u32 m = 0;
                    m |= (((t[r][0] ^ a) & TOK_MASK) == 0 && t[r][1] == b) ? 1u : 0u;
                    m |= (((t[r][1] ^ a) & TOK_MASK) == 0 && t[r][2] == b) ? 2u : 0u;
                    m |= (((t[r][2] ^ a) & TOK_MASK) == 0 && t[r][3] == b) ? 4u : 0u;
                    m |= (((t[r][3] ^ a) & TOK_MASK) == 0 && nxt[r] == b) ? 8u : 0u;
                    mn[r] = m;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 li = wbase + r * 128 + lane * 4;
                    u32 pb = __shfl_up_sync(0xffffffffu, mn[r], 1) >> 3;
                    if (lane == 0) {
                        const u32 pv = (li >= 1) ? s[li - 1] : tok(-1);
                        pb = (((pv ^ a) & TOK_MASK) == 0 && t[r][0] == b) ? 1u : 0u;
                    }
                    pbit[r] = pb;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 li = wbase + r * 128 + lane * 4;
                    const u32 d = ((mn[r] << 1) | pbit[r]) & 0xfu;
                    const int rem = (int)count - (int)li;
                    const u32 valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                    mn[r] &= valid;             // a merge only starts at a token this segment owns
                    keep[r] = ~d & valid;
                    rem_any |= keep[r] ^ 0xfu;
                    many |= mn[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) { t[r][0] = t[r][1] = t[r][2] = t[r][3] = TOK_SENTINEL; mn[r] = 0; keep[r] = 0; }
                rem_any = 1;
            }
            if (wbase < count) {
                many = __any_sync(0xffffffffu, many != 0) ? 1u : 0u;
                if (!__any_sync(0xffffffffu, rem_any != 0)) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) { lpre[r] = 4 * lane; rowoff[r] = 128 * r; rowcnt[r] = 128; }
                    wtot = MS_WSPAN;
                    plain = (many == 0);
                } else {
                    // removed tokens per lane (0..4) -> three ballots give the exclusive prefix
                    plain = false;
                    u32 run = 0;
                    const u32 lt = (1u << lane) - 1u;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        rowoff[r] = run;
                        if (!__any_sync(0xffffffffu, keep[r] != 0xfu)) {   // row keeps all of its 128 tokens
                            lpre[r] = 4 * lane;
                            rowcnt[r] = 128;
                            run += 128;
                        } else {
                            const u32 gone = 4u - __popc(keep[r]);
                            const u32 b0 = __ballot_sync(0xffffffffu, gone & 1u), b1 = __ballot_sync(0xffffffffu, gone & 2u),
                                      b2 = __ballot_sync(0xffffffffu, gone & 4u);
                            lpre[r] = 4 * lane - (__popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt));
                            rowcnt[r] = 128 - (__popc(b0) + 2 * __popc(b1) + 4 * __popc(b2));
                            run += rowcnt[r];
                        }
                    }
                    wtot = run;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) { mn[r] = 0; keep[r] = 0; lpre[r] = 0; rowoff[r] = 0; rowcnt[r] = 0; }
            }
            // ---- statistics delta of this warp's merge starts (reads the stage before anyone rewrites it) ----
            if (A.delta && many) {
                u32 mall = mn[0] | (mn[1] << 4) | (mn[2] << 8) | (mn[3] << 12);
#pragma unroll 1
                while (mall) {   // one pass per merge start of this lane
                    const int bit = __ffs(mall) - 1;
                    mall &= mall - 1;
                    delta_one(s, h, count, (int)(wbase + (bit >> 2) * 128 + lane * 4 + (bit & 3)), a, b, A.V, s_dkey, s_dcnt, A.delta);
                }
            }
            // bit 31 of the warp total = "this warp replaces a token" (a merge whose tail lies in
            // the next warp / segment changes a token without removing one)
            if (lane == 0) s_wtot[warp] = wtot | (many << 31);
            named_bar_sync(1, MS_CTHREADS);  // (1) warp totals visible; nobody reads another warp's span of the stage any more
            if (warp == 0) flush_edge();

            u32 woff = 0, new_count = 0, chg = 0;
#pragma unroll
            for (u32 k = 0; k < MS_CWARPS; ++k) {
                const u32 v = s_wtot[k];
                chg |= v >> 31;
                new_count += v & 0x7fffffffu;
                if (k < warp) woff += v & 0x7fffffffu;
            }
            const bool changed = chg || (new_count != count);
            if (!changed) {
                // untouched segment: nothing to write, the edge record carries over
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[stage]);
                if (tid == 0) e_next[seg] = e_cur[seg];
                continue;
            }

            // ---- compact in place, row by row, inside the stage; clean rows are left as they are ----
            u32 *sw = s_in + stage * MS_IN_WORDS + MS_PAD;
            if (wtot | many) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (__any_sync(0xffffffffu, keep[r] != 0xfu || mn[r] != 0)) {
                        u32 dst = wbase + r * 128 + lpre[r];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if ((keep[r] >> k) & 1u) {
                                sw[dst] = ((mn[r] >> k) & 1u) ? (z | (t[r][k] & TOK_FLAG)) : t[r][k];
                                ++dst;
                            }
                        }
                    }
                }
                __syncwarp();
                // ---- copy-out: row r holds rowcnt[r] tokens at the start of its 128-word slot ----
                u32 *__restrict__ segp = w + (u64)seg * SEG_TOKENS;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 c = rowcnt[r];
                    if (c) {
                        const u32 gdst = woff + rowoff[r];
                        const u32 *sp = sw + wbase + r * 128;
                        u32 *gp = segp + gdst;
                        if ((gdst & 3u) == 0) {   // 16-byte aligned on both sides
                            if (4 * lane + 4 <= c) reinterpret_cast<uint4 *>(gp)[lane] = reinterpret_cast<const uint4 *>(sp)[lane];
                            else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) if (4 * lane + k < c) gp[4 * lane + k] = sp[4 * lane + k];
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) { const u32 i = 32 * q + lane; if (i < c) gp[i] = sp[i]; }
                        }
                    }
                }
                // boundary tokens of the new segment that live in this warp's rows -> s_edge
                if (lane < 5) {
                    const u32 back = 5 - lane;   // lanes 3,4: positions new_count-2, new_count-1
                    const bool want = lane < 3 ? (lane < new_count) : (new_count >= back);
                    const u32 p = lane < 3 ? lane : new_count - back;
                    if (want && p >= woff && p < woff + wtot) {
                        const u32 q = p - woff;
                        const int r = (q >= rowoff[3]) ? 3 : (q >= rowoff[2]) ? 2 : (q >= rowoff[1]) ? 1 : 0;
                        const u32 ro = r == 3 ? rowoff[3] : r == 2 ? rowoff[2] : r == 1 ? rowoff[1] : rowoff[0];
                        s_edge[(j & 1u) * 8 + lane] = sw[wbase + r * 128 + (q - ro)];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);   // this warp is done with the stage
            if (tid == 0) {
                u32 *se = s_edge + (j & 1u) * 8;
                se[5] = new_count; se[6] = seg; se[7] = 1u;
                pend_par = j & 1u; pending = true;
                cta_drops += count - new_count;
            }
        }
        named_bar_sync(1, MS_CTHREADS);
        if (warp == 0) flush_edge();
        if (A.delta)
            for (u32 i = tid; i < MS_DCACHE; i += MS_CTHREADS)
                if (s_dkey[i] != 0xffffffffu && s_dcnt[i]) atomicAdd(&A.delta[s_dkey[i]], (ull)s_dcnt[i]);
        named_bar_sync(1, MS_CTHREADS);
        // ---- exit: the last CTA out publishes the new stream length and flips the edge arrays ----
        if (tid == 0) {
            if (cta_drops) atomicAdd(&ctl->drops, cta_drops);
            __threadfence();
            const u32 e = atomicAdd(&ctl->merge_exit, 1u);
            if (e == gridDim.x - 1) {
                __threadfence();
                const ull dropped = *(volatile ull *)&ctl->drops;
                ctl->sum_in += n; ctl->sum_out += n - dropped;
                ctl->n = n - dropped;
                ctl->drops = 0;
                ctl->edge_cur ^= 1u;
                ctl->iter += 1;
                ctl->epoch += 1;
                ctl->merge_ticket = 0; ctl->merge_exit = 0;
            }
        }
    }
}
