// encode_host.inl — bpe_encode (regex.py:92-121 / basic.py:57-74) on the stream kernels.
//
// "Repeatedly merge the present pair with the lowest merge index" over every chunk is the same
// as applying the merges in rank order to the whole marked stream (chunks never interact, and a
// merge only creates pairs that contain its own new id, whose ranks are all higher — SURVEY.md
// A6).  So encode is the training loop with arg-max replaced by "lowest rank whose pair count in
// the table is non-zero": one k_select_rank + k_merge + k_apply_delta round per applicable rank.

// Device scratch of the general path, kept with the (scratch) handle between calls: _encode_chunk (regex.py:92-109) and
// short encode() calls come one after the other with the same merges, and eight cudaMalloc / cudaFree pairs plus a
// rebuilt rank table per call would cost more than the kernels.  Buffers only grow; the rank table is rebuilt when the
// merges (length + 64-bit hash of the array) change.
struct GenEnc {
    unsigned char *d_bytes = nullptr; u64 bytes_cap = 0;
    u64 *d_offs = nullptr; u64 offs_cap = 0;
    u64 *d_list = nullptr; u64 list_cap = 0;
    u64 *d_keys = nullptr; u32 *d_ranks = nullptr; int *d_merges = nullptr; u64 tcap = 0;
    ull *d_cnt = nullptr;
    unsigned char *d_perm = nullptr;
    u64 merges_hash = 0; int n_merges = -1;      // the merges d_keys / d_ranks were built from (-1: none)
};

static void gen_free(bpe_handle *h) {
    GenEnc *G = h->gen;
    if (!G) return;
    cudaFree(G->d_bytes); cudaFree(G->d_offs); cudaFree(G->d_list); cudaFree(G->d_keys); cudaFree(G->d_ranks); cudaFree(G->d_merges);
    cudaFree(G->d_cnt); cudaFree(G->d_perm);
    delete G;
    h->gen = nullptr;
}

static u64 hash_words(const void *p, size_t nbytes) {     // nbytes is a multiple of 8 (pairs of int32)
    const unsigned char *b = (const unsigned char *)p;
    u64 hsh = 0x243f6a8885a308d3ull ^ (u64)nbytes;
    for (size_t i = 0; i + 8 <= nbytes; i += 8) {
        u64 w;
        memcpy(&w, b + i, 8);
        hsh = (hsh ^ w) * 0x9e3779b97f4a7c15ull;
        hsh ^= hsh >> 29;
    }
    return hsh;
}

template <class T>
static cudaError_t gen_grow(T *&ptr, u64 &cap, u64 want) {     // elements; the stream is idle between calls
#ifdef BPE_SIMT_EMU
    // emulator build (tests/emu): exactly the bytes of this call in front of the guard page, as the per-call allocations had
    cudaFree(ptr); ptr = nullptr; cap = 0;
    const u64 w = std::max<u64>(want, 1);
#else
    if (want <= cap && ptr) return cudaSuccess;
    cudaFree(ptr); ptr = nullptr; cap = 0;
    const u64 w = want + want / 4 + 256;
#endif
    const cudaError_t e = cudaMalloc(&ptr, w * sizeof(T));
    if (e == cudaSuccess) cap = w;
    return e;
}

// scratch + rank table for this call; *d_perm_out = device copy of byte_perm or NULL
static int gen_prepare(bpe_handle *h, uint64_t n, uint64_t n_chunks, const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm,
                       const unsigned char **d_perm_out) {
    if (!h->gen) h->gen = new (std::nothrow) GenEnc();
    GenEnc *G = h->gen;
    if (!G) return fail(h, BPE_ERR_INTERNAL, "out of host memory");
#define GEN_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    GEN_CU(gen_grow(G->d_bytes, G->bytes_cap, n));
    GEN_CU(gen_grow(G->d_offs, G->offs_cap, n_chunks));
    GEN_CU(gen_grow(G->d_list, G->list_cap, n / ENC_LOCAL + 1));
    if (!G->d_cnt) GEN_CU(cudaMalloc(&G->d_cnt, 16));
    if (!G->d_perm) GEN_CU(cudaMalloc(&G->d_perm, 256));
    *d_perm_out = nullptr;
    if (byte_perm) {
        GEN_CU(cudaMemcpyAsync(G->d_perm, byte_perm, 256, cudaMemcpyHostToDevice, h->stream));
        *d_perm_out = G->d_perm;
    }
    const u64 tcap = next_pow2(std::max<u64>(1024, 4ull * (u64)n_merges));
    const u64 mh = hash_words(merges, (size_t)n_merges * 8);
    if (G->n_merges == n_merges && G->merges_hash == mh && G->tcap >= tcap) return BPE_OK;
    G->n_merges = -1;                             // nothing is valid until the table is rebuilt
    if (tcap > G->tcap) {
        cudaFree(G->d_keys); cudaFree(G->d_ranks); cudaFree(G->d_merges);
        G->d_keys = nullptr; G->d_ranks = nullptr; G->d_merges = nullptr; G->tcap = 0;
        GEN_CU(cudaMalloc(&G->d_keys, tcap * 8));
        GEN_CU(cudaMalloc(&G->d_ranks, tcap * 4));
        GEN_CU(cudaMalloc(&G->d_merges, tcap * 2));   // tcap >= 4 * n_merges pairs of 8 bytes / 4
        G->tcap = tcap;
    }
    GEN_CU(cudaMemsetAsync(G->d_keys, 0xff, G->tcap * 8, h->stream));
    GEN_CU(cudaMemsetAsync(G->d_ranks, 0, G->tcap * 4, h->stream));
    if (n_merges) {
        GEN_CU(cudaMemcpyAsync(G->d_merges, merges, (size_t)n_merges * 8, cudaMemcpyHostToDevice, h->stream));
        k_rank_table_build<<<(n_merges + 255) / 256, 256, 0, h->stream>>>(G->d_merges, n_merges, G->d_keys, G->d_ranks, G->tcap - 1);
        h->tm.kernel_launches += 1;
        h->tm.h2d_bytes += (u64)n_merges * 8;
    }
    GEN_CU(cudaStreamSynchronize(h->stream));     // `merges` is the caller's buffer
    G->n_merges = n_merges; G->merges_hash = mh;
#undef GEN_CU
    return BPE_OK;
}

// Chunk-parallel encode: one thread per chunk (one CTA per long chunk), ids written at the chunk's
// byte offset, holes squeezed out per segment, then the ordinary pack + read-back.
// *handled = 0 when some chunk exceeds ENC_LONG_MAX tokens (caller uses the stream rounds instead).
static int encode_chunks_on(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
                            const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm, int32_t *out_ids,
                            uint64_t out_cap, uint64_t *out_n, int *handled) {
    int rc;
    *handled = 0;
    const u64 one = 0;
    if (!offs || n_chunks == 0) { offs = &one; n_chunks = 1; }
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0;
    if ((rc = ensure_stream_capacity(h, n))) return rc;
    const unsigned char *d_perm = nullptr;
    if ((rc = gen_prepare(h, n, n_chunks, merges, n_merges, byte_perm, &d_perm))) return rc;
    GenEnc *G = h->gen;
    unsigned char *d_bytes = G->d_bytes;
    u64 *d_offs = G->d_offs, *d_list = G->d_list, *d_keys = G->d_keys;
    u32 *d_ranks = G->d_ranks;
    ull *d_cnt = G->d_cnt;
    const u64 tcap = G->tcap;
#define ENC_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    ENC_CU(cudaMemcpyAsync(d_bytes, bytes, n, cudaMemcpyHostToDevice, h->stream));
    ENC_CU(cudaMemcpyAsync(d_offs, offs, n_chunks * 8, cudaMemcpyHostToDevice, h->stream));
    h->tm.h2d_bytes += n + n_chunks * 8;
    ENC_CU(cudaMemsetAsync(d_cnt, 0, 16, h->stream));
    RankTable rt = {d_keys, d_ranks, tcap - 1};
    k_encode_chunks<<<(unsigned)((n_chunks + 127) / 128), 128, 0, h->stream>>>(d_bytes, d_offs, n_chunks, n, rt, d_perm,
                                                                             h->buf[0], d_list, d_cnt);
    ull cnt[2] = {0, 0};
    ENC_CU(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, h->stream));
    ENC_CU(cudaStreamSynchronize(h->stream));
    h->tm.kernel_launches += 1;
    if (cnt[1]) return BPE_OK;                  // a chunk longer than ENC_LONG_MAX: not handled here
    if (cnt[0]) {
        ENC_CU(cudaFuncSetAttribute(k_encode_long, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ENC_LONG_MAX * 4));
        k_encode_long<<<(unsigned)std::min<u64>(cnt[0], (u64)h->sms * 2), 256, 2 * ENC_LONG_MAX * 4, h->stream>>>(
            d_bytes, d_offs, n_chunks, n, d_list, cnt[0], rt, d_perm, h->buf[0]);
        h->tm.kernel_launches += 1;
    }
    const u32 nseg = (u32)((n + SEG_TOKENS - 1) / SEG_TOKENS);
    k_compact_holes<<<std::max(1u, std::min<u32>(nseg, (u32)h->sms * 8)), 256, 0, h->stream>>>(h->buf[0], n, h->edge[0]);
    k_total_count<<<1, 1024, 0, h->stream>>>(h->ctl, h->edge[0], nseg);
    h->tm.kernel_launches += 2;
    ENC_CU(cudaGetLastError());
    ENC_CU(cudaStreamSynchronize(h->stream));
#undef ENC_CU
    h->loaded = true; h->bytes_only = false; h->table_valid = false;
    const u64 h2d = h->tm.h2d_bytes;
    rc = bpe_read_stream(h, out_ids, out_cap, out_n);
    h->tm.h2d_bytes = h2d;
    *handled = 1;
    return rc;
}

// Apply the merges in rank order to the loaded BYTE stream ("repeatedly merge the present pair with the lowest merge
// index", regex.py:97-108, over all chunks at once) with the training kernels: select the lowest present rank, merge,
// update the pair table incrementally.  On return the stream is the encoded text and the table equals
// get_stats(stream) — which is also exactly the state of a training run after these merges (resume: bpe_replay).
static int replay_merges_on(bpe_handle *h, const int32_t *merges, int32_t n_merges) {
    int rc;
    const u32 V = std::max(256u + (u32)n_merges, h->opt_vocab_cap);
    if ((rc = ensure_delta(h, V))) return rc;
    int *d_merges = nullptr;
    CU(cudaMalloc(&d_merges, (size_t)n_merges * 8));
    cudaError_t e = cudaMemcpyAsync(d_merges, merges, (size_t)n_merges * 8, cudaMemcpyHostToDevice, h->stream);
    h->tm.h2d_bytes += (u64)n_merges * 8;
    if (e != cudaSuccess) { cudaFree(d_merges); return fail(h, BPE_ERR_CUDA, cudaGetErrorString(e)); }
    const u64 cap = auto_table_cap(h, 0);
    if ((rc = build_table(h, cap)) || (rc = pull_ctl(h))) { cudaFree(d_merges); return rc; }
    h->h_ctl->overflow = 0;
    h->h_ctl->table_limit = (u64)(TABLE_MAX_LOAD * (double)cap);
    h->h_ctl->iter = 0; h->h_ctl->done = 0; h->h_ctl->max_iter = 0xffffffffu; h->h_ctl->found_pos = POS_NONE;
    if ((rc = push_ctl(h))) { cudaFree(d_merges); return rc; }
    // at most one round per merge rank
    int rounds_left = n_merges;
    while (rounds_left > 0 && !h->h_ctl->done) {
        const int k = std::min(h->opt_batch, rounds_left);
        const int iters_before = (int)h->h_ctl->iter;
        maybe_repack(h);
        for (int i = 0; i < k; ++i) {
            k_select_rank<<<(n_merges + 255) / 256, 256, 0, h->stream>>>(d_merges, n_merges, h->table, h->ctl);
            k_select_rank_finish<<<1, 1, 0, h->stream>>>(d_merges, h->ctl);
            launch_merge(h, h->delta, 0);
            k_apply_delta<<<(h->V + 255) / 256, 256, 0, h->stream>>>(h->table, h->ctl, h->delta, h->V, 0, 0, 0, 1, 0);
            h->tm.kernel_launches += 3;
        }
        if ((rc = pull_ctl(h))) { cudaFree(d_merges); return rc; }
        if (h->h_ctl->overflow) {
            // rounds after the overflowing one were skipped on the device: give them back
            const int performed = (int)h->h_ctl->iter - iters_before;
            if ((rc = handle_overflow(h))) { cudaFree(d_merges); return rc; }
            rounds_left -= std::max(performed, 1);
            continue;
        }
        rounds_left -= k;
    }
    cudaFree(d_merges);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string("replay: ") + cudaGetErrorString(le));
    return BPE_OK;
}

// Resume training (base.py:140-165 load() followed by more train()): bring a freshly loaded byte stream to the state
// a training run has after `n_merges` merges, so that bpe_train(more, first_idx = 256 + n_merges) continues it.
extern "C" int bpe_replay(bpe_handle *h, const int32_t *merges, int32_t n_merges) {
    if (!h || n_merges < 0 || (n_merges && !merges)) return BPE_ERR_ARG;
    if (!h->loaded || !h->bytes_only) return fail(h, BPE_ERR_STATE, "bpe_replay needs a freshly loaded byte stream");
    CU(cudaSetDevice(h->device));
    h->tm.kernel_launches = 0;
    int rc = pull_ctl(h);
    if (rc) return rc;
    if (n_merges == 0 || h->h_ctl->n < 2) return BPE_OK;
    if ((rc = replay_merges_on(h, merges, n_merges))) return rc;
    // leave the control block as bpe_train expects it
    h->h_ctl->done = 0; h->h_ctl->iter = 0; h->h_ctl->max_iter = 0; h->h_ctl->found_pos = POS_NONE;
    if ((rc = push_ctl(h))) return rc;
    h->table_valid = true;
    h->bytes_only = false;
    h->max_id = std::max(h->max_id, 255u + (u32)n_merges);
    return BPE_OK;
}

static int encode_on(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
                     const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm, int32_t *out_ids,
                     uint64_t out_cap, uint64_t *out_n) {
    int rc;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0;
    if ((rc = check_offsets(h, offs, n_chunks, n))) return rc;
    if ((rc = ensure_stream_capacity(h, n))) return rc;
    unsigned char *d_perm = nullptr;
    if (byte_perm) {
        if (!h->gen) h->gen = new (std::nothrow) GenEnc();
        if (!h->gen) return fail(h, BPE_ERR_INTERNAL, "out of host memory");
        if (!h->gen->d_perm) CU(cudaMalloc(&h->gen->d_perm, 256));
        d_perm = h->gen->d_perm;
        CU(cudaMemcpyAsync(d_perm, byte_perm, 256, cudaMemcpyHostToDevice, h->stream));
    }
    rc = load_bytes_into(h, h->buf[0], bytes, n, d_perm);
    if (rc) return rc;
    if ((rc = mark_chunks(h, h->buf[0], offs, n_chunks, n))) return rc;
    if ((rc = reset_ctl_for_stream(h, n))) return rc;
    if ((rc = build_edges(h, n))) return rc;
    h->loaded = true; h->bytes_only = true; h->table_valid = false;
    if ((rc = pull_ctl(h))) return rc;

    if (n_merges > 0 && n >= 2 && (rc = replay_merges_on(h, merges, n_merges))) return rc;
    const u64 kept_d2h = h->tm.d2h_bytes;
    rc = bpe_read_stream(h, out_ids, out_cap, out_n);
    h->tm.d2h_bytes += kept_d2h;
    return rc;
}

// The general path: thread per chunk / CTA per long chunk (k_encode.cuh), or rank-ordered rounds on the stream kernels
// for one huge chunk (BasicTokenizer on a long text) or an oversized chunk.  Runs on a scratch handle that stays with
// the parent (a stream loaded for training is not disturbed, and no handle is created per call).
static int encode_general(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *chunk_offsets, uint64_t n_chunks,
                          const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm, int32_t *out_ids, uint64_t out_cap,
                          uint64_t *out_n) {
    if (!h->enc_scratch) {
        int rc0 = bpe_create(h->device, &h->enc_scratch);
        if (rc0) return fail(h, rc0, std::string("bpe_encode: ") + bpe_last_error(nullptr));
    }
    bpe_handle *c = h->enc_scratch;
    c->opt_batch = h->opt_batch; c->opt_table_log2 = h->opt_table_log2;
    int handled = 0;
    int rc = BPE_OK;
    if (n >= 1 && n_merges > 0 && (n_chunks >= 1 || n <= ENC_LONG_MAX))
        rc = encode_chunks_on(c, bytes, n, chunk_offsets, n_chunks, merges, n_merges, byte_perm, out_ids, out_cap, out_n, &handled);
    if (!rc && !handled)
        rc = encode_on(c, bytes, n, chunk_offsets, n_chunks, merges, n_merges, byte_perm, out_ids, out_cap, out_n);
    if (rc) h->err = c->err;
    h->tm.h2d_bytes += c->tm.h2d_bytes; h->tm.d2h_bytes += c->tm.d2h_bytes; h->tm.kernel_launches += c->tm.kernel_launches;
    return rc;
}
