// decode_host.inl — bpe_decode: ids -> concatenated vocabulary bytes on the device (k_decode.cuh).

extern "C" int bpe_decode(bpe_handle *h, const int32_t *ids, uint64_t n_ids, const uint8_t *vocab_bytes, uint64_t vocab_nbytes,
                          const uint64_t *vocab_start, const uint32_t *vocab_len, int32_t V, uint8_t *out, uint64_t cap,
                          uint64_t *out_n, int64_t *bad_index) {
    if (!h || !out_n || (!ids && n_ids) || V < 0 || (V && (!vocab_start || !vocab_len)) || (vocab_nbytes && !vocab_bytes)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    *out_n = 0;
    if (bad_index) *bad_index = -1;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0;
    if (n_ids == 0) return BPE_OK;
    const u64 ntiles = (n_ids + DC_TILE - 1) / DC_TILE;
    int *d_ids = nullptr; unsigned char *d_vb = nullptr, *d_out = nullptr; u64 *d_vs = nullptr, *d_part = nullptr, *d_total = nullptr;
    u32 *d_vl = nullptr; ull *d_bad = nullptr;
    int rc = BPE_OK;
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(d_ids); cudaFree(d_vb); cudaFree(d_out); cudaFree(d_vs); cudaFree(d_part); cudaFree(d_total); cudaFree(d_vl); cudaFree(d_bad);
    };
#define DC_CU(call) do { e = (call); if (e != cudaSuccess) { cleanup(); return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e)); } } while (0)
    DC_CU(cudaMalloc(&d_ids, n_ids * 4));
    DC_CU(cudaMalloc(&d_vb, vocab_nbytes ? vocab_nbytes : 1));
    DC_CU(cudaMalloc(&d_vs, (size_t)(V ? V : 1) * 8));
    DC_CU(cudaMalloc(&d_vl, (size_t)(V ? V : 1) * 4));
    DC_CU(cudaMalloc(&d_part, ntiles * 8));
    DC_CU(cudaMalloc(&d_total, 8));
    DC_CU(cudaMalloc(&d_bad, 8));
    DC_CU(cudaMemcpyAsync(d_ids, ids, n_ids * 4, cudaMemcpyHostToDevice, h->stream));
    if (vocab_nbytes) DC_CU(cudaMemcpyAsync(d_vb, vocab_bytes, vocab_nbytes, cudaMemcpyHostToDevice, h->stream));
    if (V) {
        DC_CU(cudaMemcpyAsync(d_vs, vocab_start, (size_t)V * 8, cudaMemcpyHostToDevice, h->stream));
        DC_CU(cudaMemcpyAsync(d_vl, vocab_len, (size_t)V * 4, cudaMemcpyHostToDevice, h->stream));
    }
    DC_CU(cudaMemsetAsync(d_bad, 0xff, 8, h->stream));
    h->tm.h2d_bytes = n_ids * 4 + vocab_nbytes + (u64)V * 12;
    k_decode_reduce<<<(unsigned)ntiles, DC_THREADS, 0, h->stream>>>(d_ids, n_ids, d_vl, (u32)V, d_part, d_bad);
    k_decode_scan_parts<<<1, 1024, 0, h->stream>>>(d_part, ntiles, d_total);
    h->tm.kernel_launches = 2;
    u64 total = 0; ull bad = ~0ull;
    DC_CU(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, h->stream));
    DC_CU(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, h->stream));
    DC_CU(cudaStreamSynchronize(h->stream));
    if (bad != ~0ull) {   // regex.py:87 raises ValueError("invalid token id"), basic.py:53 KeyError: the caller maps this
        if (bad_index) *bad_index = (int64_t)bad;
        cleanup();
        return fail(h, BPE_ERR_ARG, "invalid token id");
    }
    *out_n = total;
    if (total > cap || (total && !out)) { cleanup(); return fail(h, BPE_ERR_CAPACITY, "output buffer too small"); }
    if (total) {
        DC_CU(cudaMalloc(&d_out, total));
        k_decode_copy<<<(unsigned)ntiles, DC_THREADS, 0, h->stream>>>(d_ids, n_ids, d_vs, d_vl, (u32)V, d_vb, d_part, d_out, total);
        h->tm.kernel_launches += 1;
        DC_CU(cudaGetLastError());
        DC_CU(cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, h->stream));
        DC_CU(cudaStreamSynchronize(h->stream));
        h->tm.d2h_bytes = total;
    }
#undef DC_CU
    cleanup();
    return rc;
}
