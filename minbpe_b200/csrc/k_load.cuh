// k_load.cuh — moving a corpus into / out of the int32 token stream in HBM.
// basic.py:25-26 (`list(text.encode("utf-8"))`) and regex.py:41-44 (one id list per regex chunk)
// become: widen the text bytes to 32-bit words, then set bit 31 on the first token of each chunk.
#pragma once
#include "common.cuh"

// 16 text bytes -> 16 token words per thread per step.  perm: NULL or 256-entry byte -> id map.
__global__ void __launch_bounds__(256) k_widen_bytes(const unsigned char *__restrict__ src, u32 *__restrict__ dst,
                                                     u64 n, const unsigned char *__restrict__ perm) {
    const u64 nvec = n / 16;
    for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (u64)gridDim.x * blockDim.x) {
        const uint4 q = reinterpret_cast<const uint4 *>(src)[v];
        const u32 words[4] = {q.x, q.y, q.z, q.w};
        uint4 *o = reinterpret_cast<uint4 *>(dst + v * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u32 b0 = words[j] & 255u, b1 = (words[j] >> 8) & 255u, b2 = (words[j] >> 16) & 255u, b3 = words[j] >> 24;
            if (perm) { b0 = perm[b0]; b1 = perm[b1]; b2 = perm[b2]; b3 = perm[b3]; }
            o[j] = make_uint4(b0, b1, b2, b3);
        }
    }
    // tail
    const u64 t0 = nvec * 16;
    for (u64 i = t0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 b = src[i];
        dst[i] = perm ? perm[b] : b;
    }
}

// dst[offs[i] - seg_base] |= FLAG for offsets inside [seg_base, seg_base + seg_len)
__global__ void k_set_flags(u32 *__restrict__ dst, const u64 *__restrict__ offs, u64 k, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (u64)gridDim.x * blockDim.x) {
        const u64 o = offs[i];
        if (o < n) dst[o] |= TOK_FLAG;
    }
}

// int32 ids -> token words (ids must be in [0, 2^31 - 1)); err[0] set on a bad id, err[1] = largest id
__global__ void k_copy_ids(const int *__restrict__ src, u32 *__restrict__ dst, u64 n, u32 *err) {
    u32 mx = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const int v = src[i];
        if (v < 0 || v == 0x7fffffff) err[0] = 1;
        else mx = max(mx, (u32)v);
        dst[i] = (u32)v & TOK_MASK;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(&err[1], mx);
}

// token words -> int32 ids (chunk marks stripped)
__global__ void k_strip_flags(const u32 *__restrict__ src, int *__restrict__ dst, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        dst[i] = (int)(src[i] & TOK_MASK);
}

// encode: lowest merge rank whose pair is present in the table (regex.py:99 / basic.py:64,
// min(stats, key=merges.get)).  found = min rank via atomicMin; the finishing kernel turns it
// into the next (a, b, z) or sets done.
__global__ void __launch_bounds__(256) k_select_rank(const int *__restrict__ merges, int n_merges, Table t, Ctl *ctl) {
    if (ctl->done || ctl->overflow) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    u64 hit = POS_NONE;
    if (r < n_merges) {
        const u64 slot = table_find(t, pack_pair((u32)merges[2 * r], (u32)merges[2 * r + 1]));
        if (slot != POS_NONE && t.counts[slot] != 0) hit = (u64)r;
    }
    // warp min, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 y = __shfl_xor_sync(0xffffffffu, hit, o); hit = y < hit ? y : hit; }
    if (lane_id() == 0 && hit != POS_NONE) atomicMin((ull *)&ctl->found_pos, (ull)hit);
}

__global__ void k_select_rank_finish(const int *__restrict__ merges, Ctl *ctl) {
    if (ctl->done || ctl->overflow) return;
    const u64 r = ctl->found_pos;
    if (r == POS_NONE) { ctl->done = 1; return; }
    ctl->a = merges[2 * r]; ctl->b = merges[2 * r + 1]; ctl->z = 256 + (int)r;
    ctl->found_pos = POS_NONE;
}
