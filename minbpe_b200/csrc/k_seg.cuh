// k_seg.cuh — the segmented token stream (common.cuh "segmented stream"): edge records, packing
// back to a contiguous stream, and helpers for kernels that walk the stream in order.
#pragma once
#include "common.cuh"

__device__ __forceinline__ const Edge *edges_cur(const Ctl *ctl, const Edge *e0, const Edge *e1) {
    return ctl->edge_cur ? e1 : e0;
}

// first token of the first non-empty segment after t (TOK_SENTINEL at the end of the stream)
__device__ __forceinline__ u32 seg_next_first(const Edge *e, u32 t, u32 nseg) {
    for (u32 s = t + 1; s < nseg; ++s)
        if (e[s].count) return e[s].f[0];
    return TOK_SENTINEL;
}

// The next three tokens after segment t (N[0] nearest) and the previous two (P[0] nearest),
// walking over short / empty segments.  TOK_SENTINEL past either end of the stream.
__device__ __forceinline__ void seg_neighbours(const Edge *e, u32 t, u32 nseg, u32 N[3], u32 P[2]) {
    N[0] = N[1] = N[2] = TOK_SENTINEL;
    P[0] = P[1] = TOK_SENTINEL;
    int got = 0;
    for (u32 s = t + 1; s < nseg && got < 3; ++s) {
        const u32 c = e[s].count;
        for (u32 k = 0; k < c && k < 3 && got < 3; ++k) N[got++] = e[s].f[k];
    }
    got = 0;
    for (long long s = (long long)t - 1; s >= 0 && got < 2; --s) {
        const u32 c = e[s].count;
        if (c >= 1 && got < 2) P[got++] = e[s].l[1];
        if (c >= 2 && got < 2) P[got++] = e[s].l[0];
    }
}

__device__ __forceinline__ void edge_from_tokens(Edge &out, const u32 *tok, u32 count) {
    out.count = count;
#pragma unroll
    for (int k = 0; k < 3; ++k) out.f[k] = ((u32)k < count) ? tok[k] : TOK_SENTINEL;
    out.l[1] = count >= 1 ? tok[count - 1] : TOK_SENTINEL;
    out.l[0] = count >= 2 ? tok[count - 2] : TOK_SENTINEL;
    out.pad[0] = out.pad[1] = 0;
}

// Edge records of a CONTIGUOUS stream of ctl->n tokens in the current buffer (every segment full
// except the last).  Runs after a load and after every contiguous pass; mode 1 = only when the
// stream was just packed (ctl->contig), in which case the last block clears the flag.
__global__ void __launch_bounds__(256) k_build_edges(const u32 *__restrict__ buf0, const u32 *__restrict__ buf1,
                                                     Ctl *ctl, Edge *e0, int only_if_contig) {
    if (only_if_contig && !ctl->contig) return;
    const u32 *__restrict__ w = ctl->cur ? buf1 : buf0;
    const u64 n = ctl->n;
    const u32 nseg = (u32)((n + SEG_TOKENS - 1) / SEG_TOKENS);
    for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg; t += gridDim.x * blockDim.x) {
        const u64 base = (u64)t * SEG_TOKENS;
        const u32 count = (u32)((n - base < SEG_TOKENS) ? (n - base) : SEG_TOKENS);
        Edge ed;
        edge_from_tokens(ed, w + base, count);
        e0[t] = ed;
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); last = (atomicAdd(&ctl->gather_exit, 1u) == gridDim.x - 1); }
    __syncthreads();
    if (last && threadIdx.x == 0) { ctl->gather_exit = 0; ctl->nseg = nseg; ctl->edge_cur = 0; ctl->contig = 0; }
}

// gate shared by the packing kernels: forced by the host, or device-driven for a pair (a,a)
__device__ __forceinline__ bool pack_wanted(const Ctl *ctl, int force) {
    if (force) return true;
    return !ctl->done && !ctl->overflow && ctl->iter < ctl->max_iter && ctl->a == ctl->b;
}

// exclusive prefix sum of the segment counts (one block of 32 warps; every warp owns a contiguous
// slice and walks it 32 records at a time, so the strided 32-byte records are fetched in parallel)
__global__ void __launch_bounds__(1024) k_scan_counts(const Ctl *ctl, const Edge *e0, const Edge *e1,
                                                      u64 *__restrict__ offs, int force) {
    if (!pack_wanted(ctl, force)) return;
    const Edge *e = edges_cur(ctl, e0, e1);
    const u32 nseg = ctl->nseg;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 per = ((nseg + 31) / 32 + 31) & ~31u;          // slice per warp, a multiple of 32
    const u32 lo = min(nseg, warp * per), hi = min(nseg, lo + per);
    u64 sum = 0;
    for (u32 t = lo + lane; t < hi; t += 32) sum += e[t].count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ u64 s_w[32];
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    u64 run = 0;
    for (u32 k = 0; k < warp; ++k) run += s_w[k];
    for (u32 t0 = lo; t0 < hi; t0 += 32) {
        const u32 t = t0 + lane;
        const u32 c = (t < hi) ? e[t].count : 0u;
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += v; }
        if (t < hi) offs[t] = run + (incl - c);
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// Pack the segmented stream into `dst` (contiguous).  flip=1: dst is the other ping-pong buffer
// and becomes the current, contiguous stream (ctl->contig = 1, edges stale until k_build_edges).
__global__ void __launch_bounds__(256) k_gather(Ctl *ctl, u32 *buf0, u32 *buf1, const Edge *e0, const Edge *e1,
                                                const u64 *__restrict__ offs, u32 *dst_arg, int force, int flip) {
    if (!pack_wanted(ctl, force)) return;
    const Edge *e = edges_cur(ctl, e0, e1);
    const u32 *__restrict__ w = ctl->cur ? buf1 : buf0;
    u32 *__restrict__ dst = flip ? (ctl->cur ? buf0 : buf1) : dst_arg;
    const u32 nseg = ctl->nseg;
    const u32 wpb = blockDim.x >> 5;
    for (u32 t = blockIdx.x * wpb + (threadIdx.x >> 5); t < nseg; t += gridDim.x * wpb) {   // one warp per segment
        const u32 c = e[t].count;
        const u32 *__restrict__ src = w + (u64)t * SEG_TOKENS;
        u32 *__restrict__ d = dst + offs[t];
        for (u32 i = threadIdx.x & 31; i < c; i += 32) d[i] = src[i];
    }
    if (!flip) return;
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); last = (atomicAdd(&ctl->gather_exit, 1u) == gridDim.x - 1); }
    __syncthreads();
    if (last && threadIdx.x == 0) { ctl->gather_exit = 0; ctl->cur ^= 1u; ctl->contig = 1; }
}
