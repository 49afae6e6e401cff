// split_host.inl — the GPT-4 split pattern on the device (k_split.cuh): chunk starts for training
// (marks written straight into the token stream, no offsets array) and for encode / host callers
// (compacted chunk offsets).

extern "C" int bpe_gpt4_tables(bpe_handle *h, const uint8_t *cls_table, const uint8_t *contr_table) {
    if (!h || !cls_table || !contr_table) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    if (!h->d_cls) CU(cudaMalloc(&h->d_cls, 0x110000));
    if (!h->d_contr) CU(cudaMalloc(&h->d_contr, 0x3000));
    CU(cudaMemcpyAsync(h->d_cls, cls_table, 0x110000, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_contr, contr_table, 0x3000, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BPE_OK;
}

// Working set of one split call, carved out of a slab that stays with the handle (grow-only; given
// back when the stream buffers have to grow, and in bpe_destroy): the text, its 1-byte class array,
// the flag bytes (offsets path only) and the tile aggregates — 3 bytes per text byte.
struct SplitWork {
    unsigned char *bytes = nullptr, *meta = nullptr, *flag = nullptr;
    SplFwd *fpart = nullptr; SplBwd *bpart = nullptr;
};

static void split_slab_release(bpe_handle *h) {
    if (h->split_slab) cudaFree(h->split_slab);
    h->split_slab = nullptr; h->split_cap = 0;
}

static int split_carve(bpe_handle *h, u64 n, SplitWork &W) {
    const u64 a = (n + 255) & ~255ull;
    const u64 ntiles = (n + SP_TILE - 1) / SP_TILE;
    const u64 fbytes = (ntiles * sizeof(SplFwd) + 255) & ~255ull, bbytes = (ntiles * sizeof(SplBwd) + 255) & ~255ull;
    const u64 need = 3 * a + fbytes + bbytes + 256;
    if (need > h->split_cap) {
        split_slab_release(h);
        cudaError_t e = cudaMalloc(&h->split_slab, need);
        if (e != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string("cudaMalloc(split working set): ") + cudaGetErrorString(e));
        h->split_cap = need;
    }
    unsigned char *p = h->split_slab;
    W.bytes = p; p += a;
    W.meta = p; p += a;
    W.flag = p; p += a;
    W.fpart = (SplFwd *)p; p += fbytes;
    W.bpart = (SplBwd *)p;
    return BPE_OK;
}

// host text (n bytes of UTF-8) -> W.bytes on the device; then either W.flag[i] = 1 at every chunk start
// (tokens == nullptr) or tokens[i] = byte | chunk mark.  With `hits` (flags only): the occurrences of the handle's
// special tokens (special_host.inl) are found first and become text boundaries of the split — every occurrence one
// chunk, the text on either side split on its own (regex.py:152-163); hits / which return them.
static int split_run(bpe_handle *h, const uint8_t *bytes, u64 n, SplitWork &W, u32 *tokens,
                     std::vector<u64> *hits = nullptr, std::vector<unsigned char> *which = nullptr) {
    if (!h->d_cls) return fail(h, BPE_ERR_STATE, "call bpe_gpt4_tables first");
    if (n == 0) return BPE_OK;
    if (n >= 0xfffffff0ull) return fail(h, BPE_ERR_ARG, "device split handles at most 4 GiB - 16 per call");
    int rc = split_carve(h, n, W);
    if (rc) return rc;
    CU(cudaMemcpyAsync(W.bytes, bytes, n, cudaMemcpyHostToDevice, h->stream));
    h->tm.h2d_bytes += n;
    const u32 ntiles = (u32)((n + SP_TILE - 1) / SP_TILE);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->opt_kernel_timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
    k_split_classify<<<h->sms * 8, 256, 0, h->stream>>>(W.bytes, n, h->d_cls, W.meta);
    bool with_b = false;
    if (hits && !tokens) {
        if ((rc = spec_find(h, W.bytes, n, *hits, *which))) return rc;
        if (!hits->empty()) {
            if ((rc = spec_upload_hits(h, *hits))) return rc;
            k_special_meta<<<grid_for(hits->size(), 256, h->sms * 8), 256, 0, h->stream>>>(W.meta, h->spec->d_hit, hits->size());
            h->tm.kernel_launches += 1;
            with_b = true;
        }
    }
    if (with_b) {
        k_split_reduce<true><<<ntiles, SP_THREADS, 0, h->stream>>>(W.meta, n, W.fpart, W.bpart);
        k_split_scan_parts<<<1, 1024, 0, h->stream>>>(W.fpart, W.bpart, ntiles);
        if (h->opt_split_pattern == 1) k_split_apply<false, true, 1><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, W.flag, nullptr);
        else k_split_apply<false, true, 0><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, W.flag, nullptr);
        k_special_flags<<<grid_for(hits->size(), 256, h->sms * 8), 256, 0, h->stream>>>(W.flag, h->spec->d_hit, hits->size(), n);
        h->tm.kernel_launches += 1;
    } else {
        k_split_reduce<false><<<ntiles, SP_THREADS, 0, h->stream>>>(W.meta, n, W.fpart, W.bpart);
        k_split_scan_parts<<<1, 1024, 0, h->stream>>>(W.fpart, W.bpart, ntiles);
        if (h->opt_split_pattern == 1) {
            if (tokens) k_split_apply<true, false, 1><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, nullptr, tokens);
            else k_split_apply<false, false, 1><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, W.flag, nullptr);
        } else {
            if (tokens) k_split_apply<true, false, 0><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, nullptr, tokens);
            else k_split_apply<false, false, 0><<<ntiles, SP_THREADS, 0, h->stream>>>(W.bytes, W.meta, n, h->d_contr, W.fpart, W.bpart, W.flag, nullptr);
        }
    }
    h->tm.kernel_launches += 4;
    if (e0) {   // BPE_OPT_KERNEL_TIMING: device time of the four split kernels -> bpe_timing.init_ms
        cudaEventRecord(e1, h->stream); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); h->tm.init_ms += ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    CU(cudaGetLastError());
    return BPE_OK;
}

// flags -> compacted offsets on the device (d_offs has room for n entries); *n_chunks on the host
static int flags_to_offsets(bpe_handle *h, const unsigned char *d_flag, u64 n, u64 *d_offs, u64 *n_chunks, u64 text_base = 0) {
    *n_chunks = 0;
    if (n == 0) return BPE_OK;
    const u32 ntiles = (u32)((n + SP_TILE - 1) / SP_TILE);
    u32 *part = nullptr; u64 *excl = nullptr, *d_total = nullptr;
    CU(cudaMalloc(&part, (size_t)ntiles * 4));
    cudaError_t e = cudaMalloc(&excl, (size_t)ntiles * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_total, 8);
    if (e == cudaSuccess) {
        k_flag_reduce<<<ntiles, SP_THREADS, 0, h->stream>>>(d_flag, n, part);
        k_flag_scan_parts<<<1, 1024, 0, h->stream>>>(part, excl, ntiles, d_total);
        k_flag_scatter<<<ntiles, SP_THREADS, 0, h->stream>>>(d_flag, n, excl, d_offs, text_base);
        h->tm.kernel_launches += 3;
        e = cudaMemcpyAsync(n_chunks, d_total, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(part); cudaFree(excl); cudaFree(d_total);
    if (e != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string("flags_to_offsets: ") + cudaGetErrorString(e));
    return BPE_OK;
}

// ---- texts of any size: pieces cut at provable chunk boundaries ---------------------------------------
// One split call handles < 4 GiB (32-bit tile arithmetic, 3 bytes of working set per text byte).  Longer
// texts are cut where an ASCII letter is followed by U+0020: no alternative of the GPT-4 pattern matches a
// letter followed by a space inside one chunk and the pattern has no look-behind, so
// findall(left) + findall(right) == findall(whole) at such a point (property test in
// tests/test_host.py::test_parallel_split and tests/test_gpu_split.py).  Pieces default to 1 GiB.
#define SPLIT_PIECE_BYTES (1ull << 30)

// end of the piece that starts at s: n, or the last safe cut in (s + piece/2, s + piece] that is a multiple of 4
// (the rule kernel stores the token words of a piece as 16-byte vectors: every piece must start 16-byte aligned
// in the stream buffer); 0 = none found.  With special tokens a cut must not fall inside an occurrence of one.
static u64 split_piece_end(const uint8_t *b, u64 n, u64 s, u64 piece, const SpecSet *spec = nullptr) {
    if (n - s <= piece) return n;
    const u64 hi = (s + piece) & ~3ull, lo = s + piece / 2;
    for (u64 p = hi; p > lo; p -= 4) {
        const uint8_t c = b[p - 1];
        if (b[p] == 0x20 && ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) && !spec_covers(spec, b, n, p)) return p;
    }
    return 0;
}

extern "C" int bpe_split_gpt4(bpe_handle *h, const uint8_t *bytes, uint64_t n, uint64_t *out_offsets, uint64_t cap,
                              uint64_t *n_chunks) {
    if (!h || !n_chunks || (!bytes && n)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    *n_chunks = 0;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0; h->tm.init_ms = 0;
    if (n == 0) return BPE_OK;
    const u64 piece = g_split_piece_override ? g_split_piece_override : SPLIT_PIECE_BYTES;
    u64 *d_offs = nullptr;
    int rc = BPE_OK;
    u64 total = 0;
    for (u64 s = 0; s < n && !rc;) {
        const u64 e = split_piece_end(bytes, n, s, piece);
        if (!e) { rc = fail(h, BPE_ERR_ARG, "no letter+space cut point within a piece of the text (cannot split it piecewise)"); break; }
        const u64 m = e - s;
        SplitWork W;
        rc = split_run(h, bytes + s, m, W, nullptr);
        if (!rc && !d_offs && cudaMalloc(&d_offs, std::min(n, piece) * 8) != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, "cudaMalloc offsets");
        u64 k = 0;
        if (!rc) rc = flags_to_offsets(h, W.flag, m, d_offs, &k, s);
        if (!rc) {
            if (total + k > cap) { total += k; rc = fail(h, BPE_ERR_CAPACITY, "offsets buffer too small"); break; }
            cudaError_t ce = cudaMemcpyAsync(out_offsets + total, d_offs, k * 8, cudaMemcpyDeviceToHost, h->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
            if (ce != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, cudaGetErrorString(ce));
            h->tm.d2h_bytes += k * 8;
            total += k;
        }
        s = e;
    }
    *n_chunks = total;
    cudaStreamSynchronize(h->stream);
    cudaFree(d_offs);
    return rc;
}

// regex.py:41-44 entirely on the device: upload the text, split it with the GPT-4 pattern, widen to
// the token stream with the chunk marks set.  Equivalent to bpe_load_stream(bytes, n, offsets of
// re.findall(GPT4_SPLIT_PATTERN, text)).
extern "C" int bpe_load_text_gpt4(bpe_handle *h, const uint8_t *bytes, uint64_t n, uint64_t *n_chunks) {
    if (!h || (!bytes && n)) return BPE_ERR_ARG;
    if (n >= (1ull << 36)) return fail(h, BPE_ERR_ARG, "stream too long (limit 2^36 tokens)");
    CU(cudaSetDevice(h->device));
    h->tm.h2d_bytes = 0; h->tm.kernel_launches = 0; h->tm.init_ms = 0;
    h->loaded = false; h->table_valid = false;
    int rc = ensure_stream_capacity(h, n);
    if (rc) return rc;
    u64 chunks = 0;
    const u64 piece = g_split_piece_override ? g_split_piece_override : SPLIT_PIECE_BYTES;
    u32 *part = nullptr; u64 *excl = nullptr, *d_total = nullptr;
    for (u64 s = 0; s < n;) {
        const u64 e = split_piece_end(bytes, n, s, piece);
        if (!e) return fail(h, BPE_ERR_ARG, "no letter+space cut point within a piece of the text (cannot split it piecewise)");
        const u64 m = e - s;
        SplitWork W;
        if (!n_chunks) {   // the usual case: token words with their chunk marks straight from the rule kernel
            rc = split_run(h, bytes + s, m, W, h->buf[0] + s);
            if (rc) return rc;
        } else {           // chunk count requested: flags, one counting pass over them, then widen + mark
            rc = split_run(h, bytes + s, m, W, nullptr);
            if (rc) return rc;
            k_widen_marked<<<h->sms * 8, 256, 0, h->stream>>>(W.bytes, W.flag, h->buf[0] + s, m);
            h->tm.kernel_launches += 1;
            const u32 ntiles = (u32)((m + SP_TILE - 1) / SP_TILE);
            u64 k = 0;
            if (cudaMalloc(&part, (size_t)ntiles * 4) == cudaSuccess && cudaMalloc(&excl, (size_t)ntiles * 8) == cudaSuccess &&
                cudaMalloc(&d_total, 8) == cudaSuccess) {
                k_flag_reduce<<<ntiles, SP_THREADS, 0, h->stream>>>(W.flag, m, part);
                k_flag_scan_parts<<<1, 1024, 0, h->stream>>>(part, excl, ntiles, d_total);
                cudaMemcpyAsync(&k, d_total, 8, cudaMemcpyDeviceToHost, h->stream);
            }
            cudaStreamSynchronize(h->stream);
            cudaFree(part); cudaFree(excl); cudaFree(d_total);
            part = nullptr; excl = nullptr; d_total = nullptr;
            chunks += k;
        }
        CU(cudaGetLastError());
        // a pageable source buffer is staged by the runtime: the copy has returned, the kernels may still run, and
        // the slab is reused by the next piece on the same stream — stream order keeps that safe
        s = e;
    }
    if (n_chunks) *n_chunks = chunks;
    if ((rc = reset_ctl_for_stream(h, n))) return rc;
    if ((rc = build_edges(h, n))) return rc;
    h->loaded = true; h->bytes_only = true; h->max_id = 255;
    return BPE_OK;
}
