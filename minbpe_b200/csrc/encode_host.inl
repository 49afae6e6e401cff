// encode_host.inl — bpe_encode (regex.py:92-121 / basic.py:57-74) on the stream kernels.
//
// "Repeatedly merge the present pair with the lowest merge index" over every chunk is the same
// as applying the merges in rank order to the whole marked stream (chunks never interact, and a
// merge only creates pairs that contain its own new id, whose ranks are all higher — SURVEY.md
// A6).  So encode is the training loop with arg-max replaced by "lowest rank whose pair count in
// the table is non-zero": one k_select_rank + k_merge + k_apply_delta round per applicable rank.

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

    if (n_merges > 0 && n >= 2) {
        const u32 V = 256u + (u32)n_merges;
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
        if (le != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string("encode: ") + cudaGetErrorString(le));
    }
    const u64 kept_d2h = h->tm.d2h_bytes;
    rc = bpe_read_stream(h, out_ids, out_cap, out_n);
    h->tm.d2h_bytes += kept_d2h;
    return rc;
}

extern "C" int bpe_encode(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *chunk_offsets,
                          uint64_t n_chunks, const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm,
                          int32_t *out_ids, uint64_t out_cap, uint64_t *out_n) {
    if (!h || !out_n) return BPE_ERR_ARG;
    if (!bytes && n) return fail(h, BPE_ERR_ARG, "bytes is NULL");
    if (n_merges < 0 || (n_merges && !merges)) return fail(h, BPE_ERR_ARG, "bad merges");
    if (n >= (1ull << 36)) return fail(h, BPE_ERR_ARG, "input too long");
    CU(cudaSetDevice(h->device));
    // encode works on its own scratch state so that a stream loaded for training is untouched
    bpe_handle *c = nullptr;
    int rc = bpe_create(h->device, &c);
    if (rc) return fail(h, rc, std::string("bpe_encode: ") + bpe_last_error(nullptr));
    c->opt_batch = h->opt_batch; c->opt_table_log2 = h->opt_table_log2;
    rc = encode_on(c, bytes, n, chunk_offsets, n_chunks, merges, n_merges, byte_perm, out_ids, out_cap, out_n);
    if (rc) h->err = c->err;
    h->tm.h2d_bytes = c->tm.h2d_bytes; h->tm.d2h_bytes = c->tm.d2h_bytes; h->tm.kernel_launches = c->tm.kernel_launches;
    bpe_destroy(c);
    return rc;
}

