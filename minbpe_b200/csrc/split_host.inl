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

struct SplitWork {
    unsigned char *meta = nullptr, *flag = nullptr;
    u32 *rs = nullptr, *nl = nullptr, *cnt = nullptr, *re = nullptr, *nnl = nullptr;
    Fwd *fpart = nullptr; Bwd *bpart = nullptr;
    void release() {
        cudaFree(meta); cudaFree(flag); cudaFree(rs); cudaFree(nl); cudaFree(cnt); cudaFree(re); cudaFree(nnl);
        cudaFree(fpart); cudaFree(bpart);
        *this = SplitWork();
    }
};

// d_bytes (device, n bytes of UTF-8) -> W.flag[i] = 1 at every chunk start.  Caller releases W.
static int split_flags(bpe_handle *h, const unsigned char *d_bytes, u64 n, SplitWork &W) {
    if (!h->d_cls) return fail(h, BPE_ERR_STATE, "call bpe_gpt4_tables first");
    if (n == 0) return BPE_OK;
    if (n >= 0xfffffff0ull) return fail(h, BPE_ERR_ARG, "device split handles at most 4 GiB - 16 per call");
    const u32 ntiles = (u32)((n + SP_TILE - 1) / SP_TILE);
#define SP_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { W.release(); return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    SP_CU(cudaMalloc(&W.meta, n)); SP_CU(cudaMalloc(&W.flag, n));
    SP_CU(cudaMalloc(&W.rs, n * 4)); SP_CU(cudaMalloc(&W.nl, n * 4)); SP_CU(cudaMalloc(&W.cnt, n * 4));
    SP_CU(cudaMalloc(&W.re, n * 4)); SP_CU(cudaMalloc(&W.nnl, n * 4));
    SP_CU(cudaMalloc(&W.fpart, (size_t)ntiles * sizeof(Fwd))); SP_CU(cudaMalloc(&W.bpart, (size_t)ntiles * sizeof(Bwd)));
    const int g = h->sms * 8;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->opt_kernel_timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
    k_split_classify<<<g, 256, 0, h->stream>>>(d_bytes, n, h->d_cls, W.meta);
    k_split_reduce<<<ntiles, SP_THREADS, 0, h->stream>>>(W.meta, n, W.fpart, W.bpart);
    k_split_scan_parts<<<1, 1024, 0, h->stream>>>(W.fpart, W.bpart, ntiles);
    k_split_down<<<ntiles, SP_THREADS, 0, h->stream>>>(W.meta, n, W.fpart, W.bpart, W.rs, W.nl, W.cnt, W.re, W.nnl);
    k_split_rules<<<g, 256, 0, h->stream>>>(d_bytes, n, W.meta, h->d_contr, W.rs, W.nl, W.cnt, W.re, W.nnl, W.flag);
    h->tm.kernel_launches += 5;
    if (e0) {   // BPE_OPT_KERNEL_TIMING: device time of the five split kernels -> bpe_timing.init_ms
        cudaEventRecord(e1, h->stream); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); h->tm.init_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    SP_CU(cudaGetLastError());
#undef SP_CU
    return BPE_OK;
}

// flags -> compacted offsets on the device (d_offs has room for n entries); *n_chunks on the host
static int flags_to_offsets(bpe_handle *h, const unsigned char *d_flag, u64 n, u64 *d_offs, u64 *n_chunks) {
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
        k_flag_scatter<<<ntiles, SP_THREADS, 0, h->stream>>>(d_flag, n, excl, d_offs);
        h->tm.kernel_launches += 3;
        e = cudaMemcpyAsync(n_chunks, d_total, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(part); cudaFree(excl); cudaFree(d_total);
    if (e != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string("flags_to_offsets: ") + cudaGetErrorString(e));
    return BPE_OK;
}

extern "C" int bpe_split_gpt4(bpe_handle *h, const uint8_t *bytes, uint64_t n, uint64_t *out_offsets, uint64_t cap,
                              uint64_t *n_chunks) {
    if (!h || !n_chunks || (!bytes && n)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    *n_chunks = 0;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0;
    if (n == 0) return BPE_OK;
    unsigned char *d_bytes = nullptr;
    CU(cudaMalloc(&d_bytes, n));
    cudaError_t e = cudaMemcpyAsync(d_bytes, bytes, n, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { cudaFree(d_bytes); return fail(h, BPE_ERR_CUDA, cudaGetErrorString(e)); }
    h->tm.h2d_bytes = n;
    SplitWork W;
    int rc = split_flags(h, d_bytes, n, W);
    u64 *d_offs = nullptr;
    if (!rc && cudaMalloc(&d_offs, n * 8) != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, "cudaMalloc offsets");
    if (!rc) rc = flags_to_offsets(h, W.flag, n, d_offs, n_chunks);
    if (!rc) {
        if (*n_chunks > cap) rc = fail(h, BPE_ERR_CAPACITY, "offsets buffer too small");
        else {
            e = cudaMemcpyAsync(out_offsets, d_offs, *n_chunks * 8, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, cudaGetErrorString(e));
            h->tm.d2h_bytes = *n_chunks * 8;
        }
    }
    cudaStreamSynchronize(h->stream);
    W.release(); cudaFree(d_offs); cudaFree(d_bytes);
    return rc;
}

// regex.py:41-44 entirely on the device: upload the text, split it with the GPT-4 pattern, widen to
// the token stream with the chunk marks set.  Equivalent to bpe_load_stream(bytes, n, offsets of
// re.findall(GPT4_SPLIT_PATTERN, text)).
extern "C" int bpe_load_text_gpt4(bpe_handle *h, const uint8_t *bytes, uint64_t n, uint64_t *n_chunks) {
    if (!h || (!bytes && n)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    h->tm.h2d_bytes = 0; h->tm.kernel_launches = 0;
    h->loaded = false; h->table_valid = false;
    int rc = ensure_stream_capacity(h, n);
    if (rc) return rc;
    u64 chunks = 0;
    if (n) {
        unsigned char *d_bytes = nullptr;
        CU(cudaMalloc(&d_bytes, n));
        cudaError_t e = cudaMemcpyAsync(d_bytes, bytes, n, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { cudaFree(d_bytes); return fail(h, BPE_ERR_CUDA, cudaGetErrorString(e)); }
        h->tm.h2d_bytes = n;
        SplitWork W;
        rc = split_flags(h, d_bytes, n, W);
        if (!rc) {
            k_widen_marked<<<h->sms * 8, 256, 0, h->stream>>>(d_bytes, W.flag, h->buf[0], n);
            h->tm.kernel_launches += 1;
            if (n_chunks) {   // only counted on request (one extra pass over the flags)
                const u32 ntiles = (u32)((n + SP_TILE - 1) / SP_TILE);
                u32 *part = nullptr; u64 *excl = nullptr, *d_total = nullptr;
                if (cudaMalloc(&part, (size_t)ntiles * 4) == cudaSuccess && cudaMalloc(&excl, (size_t)ntiles * 8) == cudaSuccess &&
                    cudaMalloc(&d_total, 8) == cudaSuccess) {
                    k_flag_reduce<<<ntiles, SP_THREADS, 0, h->stream>>>(W.flag, n, part);
                    k_flag_scan_parts<<<1, 1024, 0, h->stream>>>(part, excl, ntiles, d_total);
                    cudaMemcpyAsync(&chunks, d_total, 8, cudaMemcpyDeviceToHost, h->stream);
                }
                cudaStreamSynchronize(h->stream);
                cudaFree(part); cudaFree(excl); cudaFree(d_total);
            }
            e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, std::string("bpe_load_text_gpt4: ") + cudaGetErrorString(e));
        }
        W.release(); cudaFree(d_bytes);
        if (rc) return rc;
    }
    if (n_chunks) *n_chunks = chunks;
    if ((rc = reset_ctl_for_stream(h, n))) return rc;
    if ((rc = build_edges(h, n))) return rc;
    h->loaded = true; h->bytes_only = true;
    return BPE_OK;
}
