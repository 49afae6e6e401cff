// encode_host.inl — bpe_encode (regex.py:92-121 / basic.py:57-74) on the stream kernels.
//
// "Repeatedly merge the present pair with the lowest merge index" over every chunk is the same
// as applying the merges in rank order to the whole marked stream (chunks never interact, and a
// merge only creates pairs that contain its own new id, whose ranks are all higher — SURVEY.md
// A6).  So encode is the training loop with arg-max replaced by "lowest rank whose pair count in
// the table is non-zero": one k_select_rank + k_merge + k_apply_delta round per applicable rank.

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
    unsigned char *d_bytes = nullptr, *d_perm = nullptr;
    u64 *d_offs = nullptr, *d_keys = nullptr, *d_list = nullptr;
    u32 *d_ranks = nullptr;
    int *d_merges = nullptr;
    ull *d_cnt = nullptr;
    const u64 tcap = next_pow2(std::max<u64>(1024, 4ull * (u64)n_merges));
    const u64 list_cap = n / ENC_LOCAL + 1;
    auto cleanup = [&]() {
        cudaFree(d_bytes); cudaFree(d_perm); cudaFree(d_offs); cudaFree(d_keys); cudaFree(d_list); cudaFree(d_ranks);
        cudaFree(d_merges); cudaFree(d_cnt);
    };
#define ENC_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    ENC_CU(cudaMalloc(&d_bytes, n));
    ENC_CU(cudaMalloc(&d_offs, n_chunks * 8));
    ENC_CU(cudaMalloc(&d_keys, tcap * 8));
    ENC_CU(cudaMalloc(&d_ranks, tcap * 4));
    ENC_CU(cudaMalloc(&d_list, list_cap * 8));
    ENC_CU(cudaMalloc(&d_cnt, 16));
    ENC_CU(cudaMalloc(&d_merges, std::max<size_t>(8, (size_t)n_merges * 8)));
    if (byte_perm) { ENC_CU(cudaMalloc(&d_perm, 256)); ENC_CU(cudaMemcpyAsync(d_perm, byte_perm, 256, cudaMemcpyHostToDevice, h->stream)); }
    ENC_CU(cudaMemcpyAsync(d_bytes, bytes, n, cudaMemcpyHostToDevice, h->stream));
    ENC_CU(cudaMemcpyAsync(d_offs, offs, n_chunks * 8, cudaMemcpyHostToDevice, h->stream));
    if (n_merges) ENC_CU(cudaMemcpyAsync(d_merges, merges, (size_t)n_merges * 8, cudaMemcpyHostToDevice, h->stream));
    h->tm.h2d_bytes = n + n_chunks * 8 + (u64)n_merges * 8;
    ENC_CU(cudaMemsetAsync(d_keys, 0xff, tcap * 8, h->stream));
    ENC_CU(cudaMemsetAsync(d_ranks, 0, tcap * 4, h->stream));
    ENC_CU(cudaMemsetAsync(d_cnt, 0, 16, h->stream));
    if (n_merges) k_rank_table_build<<<(n_merges + 255) / 256, 256, 0, h->stream>>>(d_merges, n_merges, d_keys, d_ranks, tcap - 1);
    RankTable rt = {d_keys, d_ranks, tcap - 1};
    k_encode_chunks<<<(unsigned)((n_chunks + 127) / 128), 128, 0, h->stream>>>(d_bytes, d_offs, n_chunks, n, rt, d_perm,
                                                                             h->buf[0], d_list, d_cnt);
    ull cnt[2] = {0, 0};
    ENC_CU(cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, h->stream));
    ENC_CU(cudaStreamSynchronize(h->stream));
    h->tm.kernel_launches += 2;
    if (cnt[1]) { cleanup(); return BPE_OK; }   // a chunk longer than ENC_LONG_MAX: not handled here
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
    cleanup();
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
        CU(cudaMalloc(&d_perm, 256));
        CU(cudaMemcpyAsync(d_perm, byte_perm, 256, cudaMemcpyHostToDevice, h->stream));
    }
    rc = load_bytes_into(h, h->buf[0], bytes, n, d_perm);
    if (d_perm) { cudaStreamSynchronize(h->stream); cudaFree(d_perm); }
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
