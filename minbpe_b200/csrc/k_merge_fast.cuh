// k_merge_fast.cuh — the fused merge pass for a != b (the common case), warp-specialised:
//
//   control warp (warp 8)   takes tile tickets, keeps a 3-deep ring of 16 KB tiles in flight with
//                           1-D bulk async copies (TMA, cp.async.bulk -> UBLKCP) signalled through
//                           mbarriers, and runs the decoupled look-back for the tile offset
//   compute warps (0..7)    each owns a contiguous 512-token span of the tile (4 rows of 128 =
//                           one 16-byte shared-memory load per lane per row): mark, warp-local
//                           scan, scatter into the padded staging tile, statistics delta, and the
//                           coalesced copy-out
//
// Two block barriers per tile: (1) warp totals -> offsets / tile aggregate, (2) tile offset ->
// copy-out.  Barrier (1) of the next tile also protects the staging tile, barrier (2) frees the
// input stage for the next bulk copy.  Same arithmetic as k_merge<false>; see k_merge.cuh for the
// marking / delta rules and DESIGN.md for the byte accounting (read 4n, write 4(n-c)).
#pragma once
#include "common.cuh"
#include "k_merge.cuh"

#define MF_CWARPS 8
#define MF_CTHREADS (MF_CWARPS * 32)
#define MF_THREADS (MF_CTHREADS + 32)
#define MF_TILE 4096
#define MF_WSPAN (MF_TILE / MF_CWARPS)   // 512 tokens per compute warp
#define MF_STAGES 3
#define MF_HALO 4                        // tokens staged before and after the tile
#define MF_IN_WORDS (MF_TILE + 2 * MF_HALO)
#define MF_OUT_WORDS (MF_TILE + MF_TILE / 32 + 8)
#define MF_SMEM_BYTES (MF_STAGES * MF_IN_WORDS * 4 + MF_OUT_WORDS * 4 + 256)

__global__ void __launch_bounds__(MF_THREADS, 3) k_merge_fast(MergeArgs A) {
    Ctl *ctl = A.ctl;
    if (!A.force && (ctl->done || ctl->iter >= ctl->max_iter)) return;
    if (ctl->a == ctl->b) return;  // the (a,a) instance k_merge<true> handles this launch

    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32 *s_in = reinterpret_cast<u32 *>(smem_raw);                       // [MF_STAGES][MF_IN_WORDS]
    u32 *s_out = s_in + MF_STAGES * MF_IN_WORDS;                         // [MF_OUT_WORDS]
    u64 *s_full = reinterpret_cast<u64 *>(s_out + MF_OUT_WORDS);         // [MF_STAGES] mbarriers
    u32 *s_tileid = reinterpret_cast<u32 *>(s_full + MF_STAGES);         // [MF_STAGES]
    u32 *s_wtot = s_tileid + MF_STAGES + 1;                              // [MF_CWARPS]
    u64 *s_tileoff = reinterpret_cast<u64 *>(s_wtot + MF_CWARPS);        // [1] (8-byte aligned: see offsets)
    u32 *s_agg = reinterpret_cast<u32 *>(s_tileoff + 1);                 // [1]

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_ctrl = (warp == MF_CWARPS);
    const u64 n = ctl->n;
    const u32 *__restrict__ w = ctl->cur ? A.buf1 : A.buf0;
    u32 *__restrict__ out = ctl->cur ? A.buf0 : A.buf1;
    const u32 a = (u32)ctl->a, b = (u32)ctl->b, z = (u32)ctl->z;
    const u32 epoch = ctl->epoch;
    const u32 ntiles = (u32)((n + MF_TILE - 1) / MF_TILE);

    // ---- control-warp helper: take a ticket and start the bulk copy of that tile into `stage` ----
    auto refill = [&](u32 stage) {
        const u32 tile = atomicAdd(&ctl->merge_ticket, 1u);
        s_tileid[stage] = tile;
        if (tile < ntiles) {
            u32 *dst = s_in + stage * MF_IN_WORDS;
            const u64 ts = (u64)tile * MF_TILE;
            if (tile == 0) {   // no left halo: tokens [0, TILE+HALO) land at word MF_HALO
                mbar_arrive_expect_tx(&s_full[stage], (MF_TILE + MF_HALO) * 4);
                bulk_g2s(dst + MF_HALO, w, (MF_TILE + MF_HALO) * 4, &s_full[stage]);
            } else {           // tokens [ts-HALO, ts+TILE+HALO)
                mbar_arrive_expect_tx(&s_full[stage], MF_IN_WORDS * 4);
                bulk_g2s(dst, w + ts - MF_HALO, MF_IN_WORDS * 4, &s_full[stage]);
            }
        }
    };

    if (tid == 0) {
        for (int s = 0; s < MF_STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (is_ctrl && lane == 0) for (u32 s = 0; s < MF_STAGES; ++s) refill(s);
    __syncthreads();

    for (u32 j = 0;; ++j) {
        const u32 stage = j % MF_STAGES;
        const u32 tile = s_tileid[stage];
        if (tile >= ntiles) break;  // tickets are monotone: every later stage is past the end too
        const u64 ts = (u64)tile * MF_TILE;
        const bool full_tile = (ts + MF_TILE + MF_HALO <= n);
        const u32 *s = s_in + stage * MF_IN_WORDS;  // s[MF_HALO + i] = token ts + i

        // registers of the compute warps that live across the barriers
        u32 t[4][4], mn[4], keep[4], lpre[4], rowoff[4];
        u32 wtot = 0;
        bool anyd = false;
        const u32 lbase = MF_HALO + warp * MF_WSPAN + lane * 4;  // word index of the lane's first token (row 0)

        if (!is_ctrl) {
            mbar_wait(&s_full[stage], (j / MF_STAGES) & 1u);
            // ---- load from shared: one 16-byte read per lane per row ----
            u32 nxt[4], pbit[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint4 q = *reinterpret_cast<const uint4 *>(s + lbase + r * 128);
                t[r][0] = q.x; t[r][1] = q.y; t[r][2] = q.z; t[r][3] = q.w;
            }
            if (!full_tile) {  // last tile(s): everything at or past n reads as the sentinel
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u64 g = ts + warp * MF_WSPAN + r * 128 + lane * 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (g + k >= n) t[r][k] = TOK_SENTINEL;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                u32 v = __shfl_down_sync(0xffffffffu, t[r][0], 1);
                if (lane == 31) {
                    v = s[lbase + r * 128 + 4];
                    if (!full_tile && ts + warp * MF_WSPAN + r * 128 + 128 >= n) v = TOK_SENTINEL;
                }
                nxt[r] = v;
            }
            // ---- mark: bit k = a merge starts at the lane's k-th token ----
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                u32 m = 0;
                m |= (((t[r][0] ^ a) & TOK_MASK) == 0 && t[r][1] == b) ? 1u : 0u;
                m |= (((t[r][1] ^ a) & TOK_MASK) == 0 && t[r][2] == b) ? 2u : 0u;
                m |= (((t[r][2] ^ a) & TOK_MASK) == 0 && t[r][3] == b) ? 4u : 0u;
                m |= (((t[r][3] ^ a) & TOK_MASK) == 0 && nxt[r] == b) ? 8u : 0u;
                mn[r] = m;
            }
            // is the lane's first token the tail of a merge that starts one position earlier?
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                u32 pb = __shfl_up_sync(0xffffffffu, mn[r], 1) >> 3;
                if (lane == 0) {
                    const u64 g = ts + warp * MF_WSPAN + r * 128;
                    const u32 pv = s[lbase + r * 128 - 1];
                    pb = (g > 0 && ((pv ^ a) & TOK_MASK) == 0 && t[r][0] == b) ? 1u : 0u;
                }
                pbit[r] = pb;
            }
            // ---- warp-local scan of kept tokens ----
            u32 dany = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const u32 d = ((mn[r] << 1) | pbit[r]) & 0xfu;
                u32 valid = 0xfu;
                if (!full_tile) {
                    const long long rem = (long long)n - (long long)(ts + warp * MF_WSPAN + r * 128 + lane * 4);
                    valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                }
                keep[r] = ~d & valid;
                dany |= d | (valid ^ 0xfu);
            }
            anyd = __any_sync(0xffffffffu, dany != 0);
            if (!anyd) {  // nothing dropped in this warp's 512 tokens: positions are identity
#pragma unroll
                for (int r = 0; r < 4; ++r) { lpre[r] = 4 * lane; rowoff[r] = 128 * r; }
                wtot = MF_WSPAN;
            } else {
                u32 run = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const u32 c = __popc(keep[r]);
                    u32 incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += y; }
                    lpre[r] = incl - c;
                    rowoff[r] = run;
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                wtot = run;
            }
            if (lane == 0) s_wtot[warp] = wtot;
        }
        __syncthreads();  // (1) warp totals visible; staging tile free (previous copy-out finished)

        if (is_ctrl) {
            // ---- tile aggregate, decoupled look-back, publish the tile offset ----
            u32 v = (lane < MF_CWARPS) ? s_wtot[lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const u64 excl = tile_lookback(A.desc, tile, epoch, (u64)v);
            if (lane == 0) {
                *s_tileoff = excl; *s_agg = v;
                if (tile == ntiles - 1) ctl->n_next = excl + v;
            }
        } else {
            u32 woff = 0;
            for (u32 k = 0; k < warp; ++k) woff += s_wtot[k];
            // ---- compact into the padded staging tile ----
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                u32 dst = woff + rowoff[r] + lpre[r];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((keep[r] >> k) & 1u) {
                        const u32 v = ((mn[r] >> k) & 1u) ? (z | (t[r][k] & TOK_FLAG)) : t[r][k];
                        s_out[stage_idx(dst)] = v;
                        ++dst;
                    }
                }
            }
            // ---- statistics delta of every merge start (rules: k_merge.cuh / DESIGN.md) ----
            if (A.delta) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (mn[r]) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if ((mn[r] >> k) & 1u) {
                                const u32 i = lbase + r * 128 + k;           // shared index of p
                                const u64 p = ts + (i - MF_HALO);
                                // tokens around p, masked past the end of the stream
                                const u32 tm1 = (p >= 1) ? s[i - 1] : TOK_SENTINEL;
                                const u32 tm2 = (p >= 2) ? s[i - 2] : TOK_SENTINEL;
                                const u32 tp2 = (p + 2 < n) ? s[i + 2] : TOK_SENTINEL;
                                const u32 tp3 = (p + 3 < n) ? s[i + 3] : TOK_SENTINEL;
                                const bool m_m2 = (((tm2 ^ a) & TOK_MASK) == 0) && tm1 == b;   // merge at p-2
                                const bool m_p2 = (((tp2 ^ a) & TOK_MASK) == 0) && tp3 == b;   // merge at p+2
                                if (p >= 1 && !(t[r][k] & TOK_FLAG) && !m_m2) atomicAdd(&A.delta[tm1 & TOK_MASK], 1ull);
                                if (p + 2 < n && !(tp2 & TOK_FLAG)) {
                                    if (m_p2) atomicAdd(&A.delta[2 * (u64)A.V], 1ull);
                                    else atomicAdd(&A.delta[(u64)A.V + tp2], 1ull);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // (2) tile offset known; input stage no longer read

        if (is_ctrl) {
            if (lane == 0) refill(stage);  // next ticket into the stage just released
        } else {
            const u64 tileoff = *s_tileoff;
            const u32 agg = *s_agg;
            // element i = q*256 + tid lives at padded index i + (i >> 5) = q*264 + tid + (tid >> 5)
            u32 *__restrict__ dstp = out + tileoff + tid;
            const u32 *srcp = s_out + tid + (tid >> 5);
#pragma unroll
            for (int q = 0; q < MF_TILE / MF_CTHREADS; ++q)
                if (q * MF_CTHREADS + tid < agg) dstp[q * MF_CTHREADS] = srcp[q * (MF_CTHREADS + MF_CTHREADS / 32)];
        }
    }

    // ---- exit: the last CTA out flips the ping-pong buffers and resets the tickets ----
    if (tid == 0) {
        __threadfence();
        const u32 e = atomicAdd(&ctl->merge_exit, 1u);
        if (e == gridDim.x - 1) {
            __threadfence();
            const u64 n_new = ntiles ? *(volatile u64 *)&ctl->n_next : 0;
            ctl->sum_in += n; ctl->sum_out += n_new;
            ctl->n = n_new;
            ctl->cur ^= 1u;
            ctl->iter += 1;
            ctl->epoch += 1;
            ctl->merge_ticket = 0; ctl->merge_exit = 0;
        }
    }
}
