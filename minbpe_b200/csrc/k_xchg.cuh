// k_xchg.cuh — the per-merge exchanges of the sharded (multi-GPU) training loop, written directly over
// NVLink peer memory instead of two NCCL all-reduces per merge (regex.py:49-63 across shards).
//
// Every rank owns one exchange block in its own HBM (cudaMalloc + CUDA IPC, mapped into every peer):
//
//     XHdr                     sequence number, arrival flags, candidate slots (one per source rank)
//     delta[0], delta[1]       this rank's LOCAL statistics delta of the merge in flight (2V+1 uint64 each),
//                              double-buffered by the parity of the sequence number
//
// Per merge `s` (s = XHdr.seq, identical on all ranks because every decision that gates an iteration is taken
// from the replicated pair table):
//
//   k_xchg_cand   (only when the arg-max is tied)  every rank PUSHES its candidate word into slot [me] of every
//                 peer's header, then a release-store of the flag s+1; it waits for the N flags in its OWN
//                 header and takes the minimum — the lowest rank that sees a tied pair wins, which is the
//                 reference's first-occurrence rule (basic.py:35) across shards.  A unique max needs no
//                 exchange at all: every rank already holds the same answer.
//   k_merge_seg   adds its statistics delta into the local delta[s & 1].
//   k_xchg_apply  signals "my delta[s & 1] is complete" to every peer (release-store of s+1 into their
//                 dflag[me]), waits for the N flags in its own header, then every thread PULLS its two
//                 entries from all N ranks (coalesced 8-byte loads over NVLink), sums them and applies the
//                 result to the replicated table — the all-reduce and the table update are one kernel.  It also
//                 zeroes the local delta[(s & 1) ^ 1]: all peers have signalled s+1, so they have finished
//                 apply(s-1), the last reader of that buffer, and merge(s+1) is its next writer.
//
// Flags only ever grow (s+1), so nothing is reset and no second barrier is needed.  Peer data is read with
// ld.relaxed.sys after an ld.acquire.sys of the flag; flags are written with st.release.sys by the thread that
// wrote the data (candidates) or after the kernel boundary that completed it (delta).
#pragma once
#include "common.cuh"
#include "k_stats.cuh"

struct XArgs {
    unsigned char *peer[XCHG_MAX_RANKS];   // base of every rank's exchange block (peer[rank] = the local one)
    int world, rank;
    u64 delta_stride;                      // bytes between delta[0] and delta[1]
};

__device__ __forceinline__ XHdr *x_hdr(const XArgs &X, int r) { return reinterpret_cast<XHdr *>(X.peer[r]); }
__device__ __forceinline__ ull *x_delta(const XArgs &X, int r, u32 parity) {
    return reinterpret_cast<ull *>(X.peer[r] + XCHG_HDR_BYTES + (u64)parity * X.delta_stride);
}

#ifndef BPE_SIMT_EMU
__device__ __forceinline__ void st_release_sys_u64(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed_sys_u64(const void *p) {
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(void *p, u64 v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
#else   // CPU SIMT emulator (tests/emu): peers are other OS threads, system-scope accesses are C++ atomics
static inline void st_release_sys_u64(u64 *p, u64 v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline u64 ld_acquire_sys_u64(const u64 *p) { emu::spin(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline u64 ld_relaxed_sys_u64(const void *p) { return __atomic_load_n(reinterpret_cast<const u64 *>(p), __ATOMIC_RELAXED); }
static inline void st_relaxed_sys_u64(void *p, u64 v) { __atomic_store_n(reinterpret_cast<u64 *>(p), v, __ATOMIC_RELAXED); }
#endif

// Wait until *flag >= target.  A peer that died would leave this kernel spinning for ever and the GPU wedged:
// after ~20 s of SM clocks the wait gives up and raises ctl->overflow = 2, which gates every later kernel of the
// loop off; bpe_step_poll reports it.
#define XCHG_TIMEOUT_CYCLES 40000000000ll
__device__ __forceinline__ bool x_wait_cycles(const u64 *flag, u64 target, long long cycles) {
    const long long t0 = clock64();
    u32 spins = 0;
    while (ld_acquire_sys_u64(flag) < target) {
        if ((++spins & 0xfffu) == 0 && clock64() - t0 > cycles) return false;
    }
    return true;
}
__device__ __forceinline__ bool x_wait(const u64 *flag, u64 target, Ctl *ctl) {
    if (x_wait_cycles(flag, target, XCHG_TIMEOUT_CYCLES)) return true;
    ctl->overflow = 2;
    return false;
}

// Handshake over the freshly mapped blocks (bpe_xchg_probe): every rank PUSHES a round number into every peer's
// header and PULLS the peer's magic word, with a short timeout.  result[0] = 1 when all peers answered and every pull
// returned the magic; the host falls back to the NCCL exchange otherwise.
#define XCHG_MAGIC 0x6270655f78636867ull   /* "bpe_xchg" */
__global__ void __launch_bounds__(32) k_xchg_probe(XArgs X, long long timeout_cycles, u32 *result) {
    XHdr *me = x_hdr(X, X.rank);
    const u64 round = me->probe_round + 1;
    const int r = (int)threadIdx.x;
    bool ok = true;
    if (r < X.world) {
        st_release_sys_u64(&x_hdr(X, r)->pflag[X.rank], round);
        ok = x_wait_cycles(&me->pflag[r], round, timeout_cycles);
        if (ok) ok = ld_relaxed_sys_u64(&x_hdr(X, r)->magic) == XCHG_MAGIC;
    }
    const bool all = __all_sync(0xffffffffu, ok);
    if (r == 0) { me->probe_round = round; result[0] = all ? 1u : 0u; }
}

// ---- tie-break across shards (one warp) ----------------------------------------------------------
__global__ void __launch_bounds__(32) k_xchg_cand(Ctl *ctl, XArgs X, int *log_pairs, long long *log_counts) {
    if (ctl->done || ctl->overflow || ctl->iter >= ctl->max_iter) return;
    if (ctl->n_tied <= 1) return;                       // unique max: recorded by k_argmax, identically on every rank
    XHdr *me = x_hdr(X, X.rank);
    const u64 s1 = (u64)me->seq + 1;
    const int r = (int)threadIdx.x;
    long long word = CAND_NONE;
    if (ctl->a >= 0) word = ((long long)X.rank << 58) | ((long long)ctl->a << 29) | (long long)ctl->b;
    if (r < X.world) {
        XHdr *p = x_hdr(X, r);
        st_relaxed_sys_u64(&p->cand[X.rank], (u64)word);
        st_release_sys_u64(&p->cflag[X.rank], s1);      // same thread: the candidate is visible before the flag
    }
    long long got = CAND_NONE;
    if (r < X.world) {
        if (x_wait(&me->cflag[r], s1, ctl)) got = (long long)ld_relaxed_sys_u64(&me->cand[r]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long y = __shfl_xor_sync(0xffffffffu, got, o); got = y < got ? y : got; }
    if (r == 0) {
        if (ctl->overflow) return;                      // a peer never answered
        if (got == CAND_NONE) ctl->done = 1;            // no rank has a pair left
        else record_selection(ctl, (int)((got >> 29) & 0x1fffffff), (int)(got & 0x1fffffff), ctl->best_count, log_pairs, log_counts);
    }
}

// ---- all-reduce(SUM) of the delta vector fused with the table update -----------------------------
// One thread per token id x (as k_apply_delta).  `present` = the rank's pair-presence bitmap (tie-break
// filter, k_stats.cuh), fed from the LOCAL delta before the sum.
__global__ void __launch_bounds__(256) k_xchg_apply(Table t, Ctl *ctl, XArgs X, u32 V, u32 *present) {
    XHdr *me = x_hdr(X, X.rank);
    if (ctl->overflow || ctl->iter == me->applied) return;   // no merge ran since the last round (same decision on every rank)
    const u32 s = me->seq;
    const u64 s1 = (u64)s + 1;
    const u32 par = s & 1u;
    if (blockIdx.x == 0 && (int)threadIdx.x < X.world)  // the merge kernels of this round have completed (kernel boundary)
        st_release_sys_u64(&x_hdr(X, (int)threadIdx.x)->dflag[X.rank], s1);
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < X.world && !x_wait(&me->dflag[threadIdx.x], s1, ctl)) s_ok = 0;
    __syncthreads();
    if (!s_ok) return;                                  // block-uniform; ctl->overflow = 2 stops the loop

    const u32 a = (u32)ctl->a, b = (u32)ctl->b, z = (u32)ctl->z;
    const u64 kab = pack_pair(a, b);
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    ull *mine = x_delta(X, X.rank, par), *prev = x_delta(X, X.rank, par ^ 1u);
    if (x < V) {
        ull l = 0, r = 0;
        for (int q = 0; q < X.world; ++q) {
            const ull *d = x_delta(X, q, par);
            l += ld_relaxed_sys_u64(d + x);
            r += ld_relaxed_sys_u64(d + V + x);
        }
        if (present) {                                   // pairs this merge created in THIS shard
            if (mine[x]) present_set(present, pack_pair(x, z));
            if (mine[V + x]) present_set(present, pack_pair(z, x));
        }
        prev[x] = 0; prev[V + x] = 0;
        if (l && table_reserve(ctl)) {
            table_sub(t, pack_pair(x, a), kab, l);
            const u64 sl = table_upsert(t, pack_pair(x, z), nullptr);
            atomicAdd((ull *)&t.counts[sl], l);
        }
        if (r && table_reserve(ctl)) {
            table_sub(t, pack_pair(b, x), kab, r);
            const u64 sl = table_upsert(t, pack_pair(z, x), nullptr);
            atomicAdd((ull *)&t.counts[sl], r);
        }
    }
    if (x == 0) {
        ull zz = 0;
        for (int q = 0; q < X.world; ++q) zz += ld_relaxed_sys_u64(x_delta(X, q, par) + 2ull * V);
        if (present && mine[2ull * V]) present_set(present, pack_pair(z, z));
        prev[2ull * V] = 0;
        if (zz && table_reserve(ctl)) {
            table_sub(t, pack_pair(b, a), kab, zz);
            const u64 sl = table_upsert(t, pack_pair(z, z), nullptr);
            atomicAdd((ull *)&t.counts[sl], zz);
        }
        const u64 sl = table_find(t, kab);
        if (sl != POS_NONE) t.counts[sl] = 0;
    }
    // the last block out closes the round
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&me->exit_count, 1u) == gridDim.x - 1) { me->exit_count = 0; me->applied = ctl->iter; __threadfence(); me->seq = s + 1; }
    }
}
