// k_decode.cuh — decode (basic.py:51-55, regex.py:78-90): the bytes of vocab[id] for every id, back to
// back.  A sum scan of the token lengths (tile aggregates, one-block scan of the aggregates, in-tile
// scan) gives every token its output offset; every thread then copies the bytes of its own tokens.
#pragma once
#include "common.cuh"

#define DC_THREADS 256
#define DC_ITEMS 8
#define DC_TILE (DC_THREADS * DC_ITEMS)
#define DC_INVALID 0xffffffffu   // vocab_len[id] of an id that is not in the vocabulary

// length of token i, or 0 and *first_bad = min(first_bad, i) when the id is not in the vocabulary
__device__ __forceinline__ u32 decode_len(const int *__restrict__ ids, u64 i, const u32 *__restrict__ vlen, u32 V, ull *first_bad) {
    const int id = ids[i];
    u32 len = ((u32)id < V) ? vlen[id] : DC_INVALID;
    if (len == DC_INVALID) { atomicMin(first_bad, (ull)i); len = 0; }
    return len;
}

__global__ void __launch_bounds__(DC_THREADS) k_decode_reduce(const int *__restrict__ ids, u64 n, const u32 *__restrict__ vlen, u32 V,
                                                              u64 *__restrict__ part, ull *first_bad) {
    __shared__ u64 sm[DC_THREADS];
    const u64 base = (u64)blockIdx.x * DC_TILE + (u64)threadIdx.x * DC_ITEMS;
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < DC_ITEMS; ++k) if (base + k < n) sum += decode_len(ids, base + k, vlen, V, first_bad);
    sm[threadIdx.x] = sum;
    __syncthreads();
    for (int o = DC_THREADS / 2; o > 0; o >>= 1) { if (threadIdx.x < (u32)o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}

// exclusive scan of the tile sums (one block); total[0] = number of output bytes
__global__ void __launch_bounds__(1024) k_decode_scan_parts(u64 *__restrict__ part, u64 ntiles, u64 *total) {
    __shared__ u64 sm[1024];
    const u32 tid = threadIdx.x;
    const u64 per = (ntiles + 1023) / 1024;
    const u64 lo = min(ntiles, (u64)tid * per), hi = min(ntiles, lo + per);
    u64 sum = 0;
    for (u64 t = lo; t < hi; ++t) sum += part[t];
    sm[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { const u64 v = tid >= (u32)o ? sm[tid - o] : 0; __syncthreads(); sm[tid] += v; __syncthreads(); }
    u64 run = sm[tid] - sum;
    for (u64 t = lo; t < hi; ++t) { const u64 x = part[t]; part[t] = run; run += x; }
    if (tid == 1023) *total = sm[1023];
}

__global__ void __launch_bounds__(DC_THREADS) k_decode_copy(const int *__restrict__ ids, u64 n, const u64 *__restrict__ vstart,
                                                            const u32 *__restrict__ vlen, u32 V, const unsigned char *__restrict__ vbytes,
                                                            const u64 *__restrict__ part, unsigned char *__restrict__ out, u64 cap) {
    __shared__ u64 sm[DC_THREADS];
    const u32 tid = threadIdx.x;
    const u64 base = (u64)blockIdx.x * DC_TILE + (u64)tid * DC_ITEMS;
    u32 len[DC_ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < DC_ITEMS; ++k) {
        len[k] = 0;
        if (base + k < n) {
            const int id = ids[base + k];
            const u32 l = ((u32)id < V) ? vlen[id] : DC_INVALID;
            len[k] = (l == DC_INVALID) ? 0u : l;
        }
        sum += len[k];
    }
    sm[tid] = sum;
    __syncthreads();
    for (int o = 1; o < DC_THREADS; o <<= 1) { const u64 v = tid >= (u32)o ? sm[tid - o] : 0; __syncthreads(); sm[tid] += v; __syncthreads(); }
    u64 dst = part[blockIdx.x] + (sm[tid] - sum);
#pragma unroll
    for (int k = 0; k < DC_ITEMS; ++k) {
        if (len[k]) {
            const unsigned char *src = vbytes + vstart[ids[base + k]];
            for (u32 j = 0; j < len[k]; ++j) if (dst + j < cap) out[dst + j] = src[j];
            dst += len[k];
        }
    }
}
