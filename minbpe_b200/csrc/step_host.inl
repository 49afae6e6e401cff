// step_host.inl — step-wise training for the sharded (multi-GPU) loop.  One process per GPU; the
// host (minbpe_b200/dist.py) interleaves these calls with two tiny collectives per merge:
//
//     bpe_step_begin   local byte-pair histogram            -> all-reduce(SUM) of 65536 counts
//     bpe_step_table   build the replicated global table
//   per merge:
//     bpe_step_select  arg-max (+ local first occurrence)   -> all-reduce(MIN) of one int64
//     bpe_step_merge   commit the winner, merge the shard   -> all-reduce(SUM) of the delta vector
//     bpe_step_apply   apply the summed delta to the table
//   every few merges:
//     bpe_step_poll    host sync: done?, table head-room, segment re-packing
//
// Everything is enqueued on one stream (bpe_set_stream: the caller's, e.g. torch's current stream,
// so NCCL collectives issued by torch are ordered with these kernels without host syncs).

extern "C" int bpe_set_stream(bpe_handle *h, void *cuda_stream) {
    if (!h) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return BPE_OK;
}

extern "C" int bpe_step_begin(bpe_handle *h, uint64_t *dense_dev) {
    if (!h || !dense_dev) return BPE_ERR_ARG;
    if (!h->loaded || !h->bytes_only) return fail(h, BPE_ERR_STATE, "bpe_step_begin needs a freshly loaded byte stream");
    CU(cudaSetDevice(h->device));
    CU(cudaMemsetAsync(dense_dev, 0, 65536 * 8, h->stream));
    CU(cudaMemsetAsync(h->d_err, 0, 4, h->stream));
    h->tm.kernel_launches = 0;
    {
        int rc0 = hist_dense(h, (ull *)dense_dev);
        if (rc0) return rc0;
    }
    // which pairs occur in THIS shard (before the caller sums the histogram across ranks)
    if (!h->d_present) CU(cudaMalloc(&h->d_present, (1u << PRESENT_LOG2) / 8));
    CU(cudaMemsetAsync(h->d_present, 0, (1u << PRESENT_LOG2) / 8, h->stream));
    k_present_init<<<65536 / 256, 256, 0, h->stream>>>((const ull *)dense_dev, h->d_present);
    CU(cudaGetLastError());
    h->tm.kernel_launches += 1;
    return BPE_OK;
}

// grow the table until `iters` worst-case iterations (2V+1 new pairs each) fit under the load limit
static int step_headroom(bpe_handle *h, int iters) {
    int rc;
    for (;;) {
        const u64 cap = h->table.mask + 1;
        const double room = TABLE_MAX_LOAD * (double)cap - (double)h->h_ctl->table_used;
        if (room >= (double)iters * (2.0 * h->V + 1)) return BPE_OK;
        if ((rc = rehash_table(h, cap * 2))) return rc;
        if ((rc = pull_ctl(h))) return rc;
        h->h_ctl->table_limit = (u64)(TABLE_MAX_LOAD * (double)(cap * 2));
        if ((rc = push_ctl(h))) return rc;
    }
}

extern "C" int bpe_step_table(bpe_handle *h, const uint64_t *dense_dev, int32_t num_merges, int32_t first_idx, int32_t poll_every) {
    if (!h || !dense_dev || num_merges <= 0 || first_idx < 0 || poll_every <= 0) return BPE_ERR_ARG;
    if ((u64)first_idx + (u64)num_merges >= (1ull << 29)) return fail(h, BPE_ERR_ARG, "sharded training needs ids < 2^29");
    CU(cudaSetDevice(h->device));
    int rc;
    const u32 V = (u32)first_idx + (u32)num_merges;
    h->V = V;   // the delta vector is the caller's buffer in this mode
    if (h->log_cap < num_merges) {
        if (h->log_pairs) cudaFree(h->log_pairs);
        if (h->log_counts) cudaFree(h->log_counts);
        h->log_pairs = nullptr; h->log_counts = nullptr; h->log_cap = 0;
        CU(cudaMalloc(&h->log_pairs, (size_t)num_merges * 8));
        CU(cudaMalloc(&h->log_counts, (size_t)num_merges * 8));
        h->log_cap = num_merges;
    }
    u32 bad = 0;
    CU(cudaMemcpyAsync(&bad, h->d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    const u64 cap = auto_table_cap(h, 0);
    if (!h->table.keys || h->table.mask + 1 != cap) {
        free_table(h, h->table);
        if ((rc = alloc_table(h, h->table, cap, false))) return rc;
    } else {
        CU(cudaMemsetAsync(h->table.keys, 0xff, cap * 8, h->stream));
        CU(cudaMemsetAsync(h->table.counts, 0, cap * 8, h->stream));
    }
    const ull zero = 0;
    CU(cudaMemcpyAsync(&h->ctl->table_used, &zero, 8, cudaMemcpyHostToDevice, h->stream));
    k_dense_to_table<<<65536 / 256, 256, 0, h->stream>>>((const ull *)dense_dev, h->table, h->ctl);
    CU(cudaGetLastError());
    if ((rc = pull_ctl(h))) return rc;
    if (bad) return fail(h, BPE_ERR_INTERNAL, "byte stream contains ids >= 256");
    h->h_ctl->iter = 0; h->h_ctl->done = 0; h->h_ctl->first_idx = (u32)first_idx; h->h_ctl->max_iter = (u32)num_merges;
    h->h_ctl->sum_in = 0; h->h_ctl->sum_out = 0; h->h_ctl->overflow = 0;
    h->h_ctl->table_limit = (u64)(TABLE_MAX_LOAD * (double)cap);
    if ((rc = push_ctl(h))) return rc;
    h->table_valid = false;
    h->tm.merge_kernel_ms = 0;
    if (h->xchg) {   // ctl->iter restarts at 0; the buffer of the previous run's last round is still dirty (the caller has
                     // synchronised the ranks, so no peer reads it any more)
        CU(cudaMemsetAsync(h->xchg + offsetof(XHdr, applied), 0, 4, h->stream));
        CU(cudaMemsetAsync(h->xchg + XCHG_HDR_BYTES, 0, 2 * h->xchg_stride, h->stream));
    }
    h->step_poll_every = poll_every;
    h->tm.kernel_launches += 1;
    return step_headroom(h, poll_every);
}

extern "C" int bpe_step_select(bpe_handle *h, int64_t *cand_dev, int32_t rank) {
    if (!h || !cand_dev || rank < 0 || rank >= 32) return BPE_ERR_ARG;
    k_argmax<<<h->argmax_grid, 256, 0, h->stream>>>(h->table, h->ctl, h->partials, h->log_pairs, h->log_counts);
    k_tie_present<<<h->argmax_grid, 256, 0, h->stream>>>(h->table, h->ctl, h->d_present);
    k_find_first<<<h->ff_grid, 256, 0, h->stream>>>(h->buf[0], h->buf[1], h->edge[0], h->edge[1], h->table, h->ctl,
                                                    h->log_pairs, h->log_counts, 1);
    k_pack_candidate<<<1, 1, 0, h->stream>>>(h->ctl, (long long *)cand_dev, rank);
    h->tm.kernel_launches += 4;
    CU(cudaGetLastError());
    return BPE_OK;
}

extern "C" int bpe_step_merge(bpe_handle *h, const int64_t *cand_dev, uint64_t *delta_dev) {
    if (!h || !cand_dev || !delta_dev) return BPE_ERR_ARG;
    k_commit_candidate<<<1, 1, 0, h->stream>>>(h->ctl, (const long long *)cand_dev, h->log_pairs, h->log_counts);
    h->tm.kernel_launches += 1;
    timed_merge(h, (ull *)delta_dev);
    // pairs this merge created in THIS shard (the delta is still local here; the caller sums it next)
    k_present_update<<<(h->V + 255) / 256, 256, 0, h->stream>>>((const ull *)delta_dev, h->V, h->ctl, h->d_present);
    h->tm.kernel_launches += 1;
    CU(cudaGetLastError());
    return BPE_OK;
}

extern "C" int bpe_step_apply(bpe_handle *h, uint64_t *delta_dev) {
    if (!h || !delta_dev) return BPE_ERR_ARG;
    k_apply_delta<<<(h->V + 255) / 256, 256, 0, h->stream>>>(h->table, h->ctl, (ull *)delta_dev, h->V, 0, 0, 0, 1, 0);
    h->tm.kernel_launches += 1;
    CU(cudaGetLastError());
    return BPE_OK;
}

extern "C" int bpe_step_delta_len(bpe_handle *h, uint64_t *len) {
    if (!h || !len) return BPE_ERR_ARG;
    *len = 2ull * h->V + 1;
    return BPE_OK;
}

// Host synchronisation point of the sharded loop: reports progress, keeps `poll_every` worst-case
// iterations of table head-room, re-packs sparse segments.  Identical decisions on every rank as
// far as the table goes (same table everywhere); re-packing is a local matter.
extern "C" int bpe_step_poll(bpe_handle *h, int32_t *iters_done, int32_t *exhausted) {
    if (!h || !iters_done || !exhausted) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    int rc = pull_ctl(h);
    if (rc) return rc;
    drain_kernel_events(h);
    if (h->h_ctl->overflow == 2) return fail(h, BPE_ERR_INTERNAL, "NVLink exchange timed out waiting for a peer rank");
    if (h->h_ctl->overflow) return fail(h, BPE_ERR_INTERNAL, "pair table overflowed despite the head-room guarantee");
    *iters_done = (int32_t)h->h_ctl->iter;
    *exhausted = h->h_ctl->done != 0;
    h->tm.tokens_in = h->h_ctl->sum_in; h->tm.tokens_out = h->h_ctl->sum_out;
    h->tm.table_slots = h->table.mask + 1; h->tm.table_used = h->h_ctl->table_used;
    if ((rc = step_headroom(h, h->step_poll_every))) return rc;
    maybe_repack(h);
    return BPE_OK;
}

// Copy the merges performed so far (pairs + global counts) to the host.
extern "C" int bpe_step_result(bpe_handle *h, int32_t *out_pairs, int64_t *out_counts, int32_t cap, int32_t *n_done) {
    if (!h || !n_done) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    int rc = pull_ctl(h);
    if (rc) return rc;
    const int done = (int)h->h_ctl->iter;
    *n_done = done;
    if (done > cap) return fail(h, BPE_ERR_CAPACITY, "output buffers too small");
    if (done > 0) {
        CU(cudaMemcpyAsync(out_pairs, h->log_pairs, (size_t)done * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(out_counts, h->log_counts, (size_t)done * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->h_ctl->first_idx + done > 256) h->bytes_only = false;
    return BPE_OK;
}


// ================================================================================================
// The same loop with the per-merge exchanges done by our own kernels over NVLink peer memory
// (k_xchg.cuh) instead of two NCCL all-reduces issued by the host: no host call per merge at all.
//
//   bpe_xchg_create   allocate this rank's exchange block, return its CUDA IPC handle (64 bytes)
//   (host: all-gather the handles of all ranks, e.g. torch.distributed)
//   bpe_xchg_attach   map every peer's block
//   bpe_step_begin / (all-reduce of the byte-pair histogram, once) / bpe_step_table   as above
//   bpe_step_fused    enqueue n merge iterations: arg-max, tie filter, local first occurrence, candidate
//                     exchange (ties only), merge, delta exchange fused with the table update
//   bpe_step_poll / bpe_step_result   as above
// ================================================================================================
static void xchg_close_peers(bpe_handle *h) {
    for (int r = 0; r < h->xargs.world; ++r)
        if (r != h->xargs.rank && h->xargs.peer[r]) { cudaIpcCloseMemHandle(h->xargs.peer[r]); h->xargs.peer[r] = nullptr; }
    h->xchg_attached = false;
}

static void xchg_release(bpe_handle *h) {
    xchg_close_peers(h);
    if (h->xchg) cudaFree(h->xchg);
    h->xchg = nullptr; h->xchg_bytes = 0; h->xchg_V = 0;
    memset(&h->xargs, 0, sizeof(h->xargs));
}

extern "C" int bpe_xchg_create(bpe_handle *h, int32_t world, int32_t rank, int32_t vocab_cap, uint8_t *ipc_handle_out) {
    if (!h || world < 1 || world > XCHG_MAX_RANKS || rank < 0 || rank >= world || vocab_cap < 256 || !ipc_handle_out) return BPE_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    xchg_release(h);
    const u64 stride = ((2ull * (u64)vocab_cap + 1) * 8 + 255) & ~255ull;
    const u64 bytes = XCHG_HDR_BYTES + 2 * stride;
    CU(cudaMalloc(&h->xchg, bytes));
    CU(cudaMemset(h->xchg, 0, bytes));
    {
        const u64 magic = XCHG_MAGIC;
        CU(cudaMemcpy(h->xchg + offsetof(XHdr, magic), &magic, 8, cudaMemcpyHostToDevice));
    }
    h->xchg_bytes = bytes; h->xchg_stride = stride; h->xchg_V = (u32)vocab_cap;
    h->xargs.world = world; h->xargs.rank = rank; h->xargs.delta_stride = stride;
    h->xargs.peer[rank] = h->xchg;
    if (world > 1) {
        cudaIpcMemHandle_t ipc;
        CU(cudaIpcGetMemHandle(&ipc, h->xchg));
        memcpy(ipc_handle_out, &ipc, 64);
    } else {              // nobody maps the block of a single rank: no IPC needed (containers may refuse it)
        memset(ipc_handle_out, 0, 64);
        h->xchg_attached = true;
    }
    return BPE_OK;
}

// Unmap the peers' blocks.  A block must not be freed (bpe_xchg_create again, bpe_destroy) while another process
// still maps it: every rank detaches, the host synchronises the ranks, then the blocks may go.
extern "C" int bpe_xchg_detach(bpe_handle *h) {
    if (!h) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    xchg_close_peers(h);
    return BPE_OK;
}

extern "C" int bpe_xchg_attach(bpe_handle *h, const uint8_t *all_handles) {
    if (!h || !all_handles) return BPE_ERR_ARG;
    if (!h->xchg) return fail(h, BPE_ERR_STATE, "call bpe_xchg_create first");
    CU(cudaSetDevice(h->device));
    for (int r = 0; r < h->xargs.world; ++r) {
        if (r == h->xargs.rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, all_handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(h, BPE_ERR_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e) +
                                             " (peer access between the GPUs of one NVLink box is required)");
        h->xargs.peer[r] = (unsigned char *)p;
    }
    h->xchg_attached = true;
    return BPE_OK;
}

// Handshake with every mapped peer (push a flag, wait for theirs, pull their magic word) with a short timeout:
// *ok = 1 when the peer-memory path works in both directions.  Every rank must call it (it waits for the others).
extern "C" int bpe_xchg_probe(bpe_handle *h, int32_t timeout_ms, int32_t *ok) {
    if (!h || !ok || timeout_ms <= 0) return BPE_ERR_ARG;
    *ok = 0;
    if (!h->xchg_attached) return fail(h, BPE_ERR_STATE, "bpe_xchg_probe needs bpe_xchg_create + bpe_xchg_attach");
    CU(cudaSetDevice(h->device));
    int khz = 1500000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
    const long long cycles = (long long)timeout_ms * (long long)khz;
    k_xchg_probe<<<1, 32, 0, h->stream>>>(h->xargs, cycles, h->d_err);
    u32 res = 0;
    CU(cudaMemcpyAsync(&res, h->d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *ok = (int32_t)res;
    return BPE_OK;
}

extern "C" int bpe_step_fused(bpe_handle *h, int32_t n_iters) {
    if (!h || n_iters < 0) return BPE_ERR_ARG;
    if (!h->xchg_attached) return fail(h, BPE_ERR_STATE, "bpe_step_fused needs bpe_xchg_create + bpe_xchg_attach");
    if (h->V != h->xchg_V) return fail(h, BPE_ERR_STATE, "exchange block was created for another vocabulary capacity");
    if (!h->d_present) return fail(h, BPE_ERR_STATE, "call bpe_step_begin / bpe_step_table first");
    for (int i = 0; i < n_iters; ++i) {
        k_argmax<<<h->argmax_grid, 256, 0, h->stream>>>(h->table, h->ctl, h->partials, h->log_pairs, h->log_counts);
        k_tie_present<<<h->argmax_grid, 256, 0, h->stream>>>(h->table, h->ctl, h->d_present);
        k_find_first<<<h->ff_grid, 256, 0, h->stream>>>(h->buf[0], h->buf[1], h->edge[0], h->edge[1], h->table, h->ctl,
                                                        h->log_pairs, h->log_counts, 1);
        k_xchg_cand<<<1, 32, 0, h->stream>>>(h->ctl, h->xargs, h->log_pairs, h->log_counts);
        timed_merge(h, nullptr, true);
        k_xchg_apply<<<(h->V + 255) / 256, 256, 0, h->stream>>>(h->table, h->ctl, h->xargs, h->V, h->d_present);
        h->tm.kernel_launches += 5;
    }
    CU(cudaGetLastError());
    return BPE_OK;
}
