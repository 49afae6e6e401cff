// encode2_host.inl — host side of the memoised chunk encode (k_encode2.cuh): persistent state in the handle
// (rank table, memo table, id pool, scratch), bpe_encode_text_gpt4 (split + encode fused on the device, no
// offsets anywhere) and the piece loop shared with bpe_encode(chunk_offsets).

struct EncState {
    // rank table of the merges the memo was built for
    u64 merges_hash = 0; int n_merges = -1; bool has_perm = false; u64 perm_hash = 0;
    u64 spec_hash = 0;                // special tokens seeded into the memo (0 = none)
    int *d_merges = nullptr; u64 *d_rkeys = nullptr; u32 *d_rranks = nullptr; u64 rt_cap = 0;
    unsigned char *d_perm = nullptr;
    MemoSlot *memo = nullptr; u64 memo_cap = 0;
    u32 *pool = nullptr; u64 pool_cap = 0;
    u32 *tmp = nullptr; u64 tmp_cap = 0;          // per-piece id area (direct chunks)
    u32 *new_list = nullptr;
    u64 *direct_list = nullptr; u64 direct_cap = 0;
    PosSlot *posmap = nullptr; u64 pos_cap = 0;
    EncCtl *ctl = nullptr;
    u32 *part = nullptr; u64 *excl = nullptr; u64 *d_total = nullptr; u64 parts_cap = 0;
    int *d_ids = nullptr; u64 ids_cap = 0;
    u64 *d_offs = nullptr; u64 offs_cap = 0;     // staging of host chunk offsets (bpe_encode with offsets)
    // statistics of the last call (bpe_encode_stats)
    u64 st_chunks_new = 0, st_direct = 0, st_long = 0, st_pieces = 0, st_fallbacks = 0, st_retries = 0, st_tmp_ids = 0;
    double st_kernel_ms = 0;
};

#define ENC2_MEMO_LOG2 22            /* 4 Mi slots x 64 B = 256 MiB */
#define ENC2_POOL_IDS (64ull << 20)  /* 64 Mi ids = 256 MiB */

static u64 fnv64(const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    u64 h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

static void enc2_free(bpe_handle *h) {
    EncState *S = h->enc;
    if (!S) return;
    cudaFree(S->d_merges); cudaFree(S->d_rkeys); cudaFree(S->d_rranks); cudaFree(S->d_perm); cudaFree(S->memo); cudaFree(S->pool); cudaFree(S->tmp);
    cudaFree(S->new_list); cudaFree(S->direct_list); cudaFree(S->posmap); cudaFree(S->ctl); cudaFree(S->part); cudaFree(S->excl);
    cudaFree(S->d_total); cudaFree(S->d_ids); cudaFree(S->d_offs);
    delete S;
    h->enc = nullptr;
}

#define E2CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, BPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

static Enc2 enc2_args(const EncState *S) {
    Enc2 E;
    E.text = nullptr; E.flag = nullptr; E.n = 0;
    E.memo = S->memo; E.memo_mask = S->memo_cap - 1; E.memo_limit = S->memo_cap / 2;
    E.pool = S->pool; E.pool_cap = S->pool_cap;
    E.tmp = S->tmp; E.tmp_cap = S->tmp_cap;
    E.new_list = S->new_list; E.new_cap = (u32)S->memo_cap;
    E.direct_list = S->direct_list; E.direct_cap = S->direct_cap;
    E.posmap = nullptr; E.pos_mask = 0;
    E.ctl = S->ctl;
    return E;
}

// empty table; the special tokens of the current call (if any) are entries from the start
static int enc2_reset_memo(bpe_handle *h) {
    EncState *S = h->enc;
    E2CU(cudaMemsetAsync(S->memo, 0, S->memo_cap * sizeof(MemoSlot), h->stream));
    E2CU(cudaMemsetAsync(S->ctl, 0, sizeof(EncCtl), h->stream));
    const SpecSet *P = h->spec;
    if (S->spec_hash && P && P->k) {
        k_enc_seed_specials<<<1, SPEC_MAX, 0, h->stream>>>(enc2_args(S), P->d_blob, P->d_off, P->d_ids, P->k);
        h->tm.kernel_launches += 1;
    }
    return BPE_OK;
}

// rank table + memo for this merges table (kept while the caller keeps passing the same merges)
static int enc2_prepare(bpe_handle *h, const int32_t *merges, int32_t n_merges, const uint8_t *perm, bool use_spec = false) {
    if (!h->enc) h->enc = new (std::nothrow) EncState();
    EncState *S = h->enc;
    if (!S) return fail(h, BPE_ERR_INTERNAL, "out of host memory");
    if (!S->memo || !S->pool || !S->new_list || !S->ctl || !S->d_total || !S->d_perm) {   // (an earlier call may have run out of memory half-way)
        S->memo_cap = 1ull << (h->opt_memo_log2 ? h->opt_memo_log2 : ENC2_MEMO_LOG2);
        S->pool_cap = ENC2_POOL_IDS;
        S->n_merges = -1;
        if (!S->memo) E2CU(cudaMalloc(&S->memo, S->memo_cap * sizeof(MemoSlot)));
        if (!S->pool) E2CU(cudaMalloc(&S->pool, S->pool_cap * 4));
        if (!S->new_list) E2CU(cudaMalloc(&S->new_list, S->memo_cap * 4));
        if (!S->ctl) E2CU(cudaMalloc(&S->ctl, sizeof(EncCtl)));
        if (!S->d_total) E2CU(cudaMalloc(&S->d_total, 8));
        if (!S->d_perm) E2CU(cudaMalloc(&S->d_perm, 256));
    }
    const u64 mh = fnv64(merges, (size_t)n_merges * 8), ph = perm ? fnv64(perm, 256) : 0;
    const u64 sh = (use_spec && h->spec && h->spec->k) ? h->spec->hash : 0;
    if (S->n_merges == n_merges && S->merges_hash == mh && S->has_perm == (perm != nullptr) && S->perm_hash == ph && S->spec_hash == sh)
        return BPE_OK;
    S->spec_hash = sh;
    S->n_merges = -1;                         // nothing is valid until the end of this function
    const u64 tcap = next_pow2(std::max<u64>(1024, 4ull * (u64)n_merges));
    if (tcap > S->rt_cap) {
        cudaFree(S->d_rkeys); cudaFree(S->d_rranks); cudaFree(S->d_merges);
        S->d_rkeys = nullptr; S->d_rranks = nullptr; S->d_merges = nullptr; S->rt_cap = 0;
        E2CU(cudaMalloc(&S->d_rkeys, tcap * 8));
        E2CU(cudaMalloc(&S->d_rranks, tcap * 4));
        E2CU(cudaMalloc(&S->d_merges, tcap * 2));   // tcap >= 4 * n_merges entries of 8 bytes / 4
        S->rt_cap = tcap;
    }
    E2CU(cudaMemsetAsync(S->d_rkeys, 0xff, S->rt_cap * 8, h->stream));
    E2CU(cudaMemsetAsync(S->d_rranks, 0, S->rt_cap * 4, h->stream));
    if (n_merges) {
        E2CU(cudaMemcpyAsync(S->d_merges, merges, (size_t)n_merges * 8, cudaMemcpyHostToDevice, h->stream));
        k_rank_table_build<<<(n_merges + 255) / 256, 256, 0, h->stream>>>(S->d_merges, n_merges, S->d_rkeys, S->d_rranks, S->rt_cap - 1);
    }
    if (perm) E2CU(cudaMemcpyAsync(S->d_perm, perm, 256, cudaMemcpyHostToDevice, h->stream));
    int rc = enc2_reset_memo(h);
    if (rc) return rc;
    E2CU(cudaStreamSynchronize(h->stream));   // `merges` / `perm` are the caller's buffers
    S->merges_hash = mh; S->n_merges = n_merges; S->has_perm = perm != nullptr; S->perm_hash = ph;
    return BPE_OK;
}

// One piece: text bytes and chunk-start flags are on the device (m bytes).  Ids are appended to out[*written..).
// *fell_back = 1 when the piece needs the general path (nothing was written for it).
static int enc2_piece(bpe_handle *h, const unsigned char *d_text, const unsigned char *d_flag, u64 m,
                      int32_t *out, u64 cap, u64 *written, int *fell_back) {
    EncState *S = h->enc;
    *fell_back = 0;
    if (m == 0) return BPE_OK;
    const u32 ntiles = (u32)((m + E2_TILE - 1) / E2_TILE);
    if (ntiles > S->parts_cap) {
        cudaFree(S->part); cudaFree(S->excl); S->part = nullptr; S->excl = nullptr; S->parts_cap = 0;
        E2CU(cudaMalloc(&S->part, (size_t)ntiles * 4));
        E2CU(cudaMalloc(&S->excl, (size_t)ntiles * 8));
        S->parts_cap = ntiles;
    }
    const u64 dcap = m / 16 + 65536;
    if (dcap > S->direct_cap) {
        cudaFree(S->direct_list); S->direct_list = nullptr; S->direct_cap = 0;
        E2CU(cudaMalloc(&S->direct_list, dcap * 8));
        S->direct_cap = dcap;
    }
    // per-piece id area for the direct chunks: a guess first (an eighth of the bytes), the worst case (one id per byte)
    // if the guess turns out too small
    {
        const u64 floor_ids = h->opt_memo_log2 ? 1024 : (1ull << 22);      // the test hook also shrinks this
        const u64 want = std::max<u64>(floor_ids, m / 8);
        if (S->tmp_cap < want) {
            cudaFree(S->tmp); S->tmp = nullptr; S->tmp_cap = 0;
            E2CU(cudaMalloc(&S->tmp, want * 4));
            S->tmp_cap = want;
        }
    }
    const RankTable rt = {S->d_rkeys, S->d_rranks, S->rt_cap - 1};
    const unsigned char *perm = S->has_perm ? S->d_perm : nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->opt_kernel_timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
    EncCtl hc;
    Enc2 E;
    bool bad = false;
    u64 total = 0, n_new_all = 0;
    for (int attempt = 0;; ++attempt) {
        // per-piece counters
        E2CU(cudaMemcpyAsync(&hc, S->ctl, sizeof(EncCtl), cudaMemcpyDeviceToHost, h->stream));
        E2CU(cudaStreamSynchronize(h->stream));
        hc.n_new = 0; hc.n_direct = 0; hc.n_long = 0; hc.fail = 0; hc.tmp_used = 0;
        E2CU(cudaMemcpyAsync(S->ctl, &hc, sizeof(EncCtl), cudaMemcpyHostToDevice, h->stream));
        E = enc2_args(S);
        E.text = d_text; E.flag = d_flag; E.n = m;

        k_enc_insert<<<ntiles, E2_THREADS, 0, h->stream>>>(E);
        E2CU(cudaMemcpyAsync(&hc, S->ctl, sizeof(EncCtl), cudaMemcpyDeviceToHost, h->stream));
        E2CU(cudaStreamSynchronize(h->stream));
        h->tm.kernel_launches += 1;
        n_new_all += hc.n_new;
        bad = hc.fail != 0;
        if (!bad && hc.n_new) {
            k_enc_distinct<<<(hc.n_new + 127) / 128, 128, 0, h->stream>>>(E, rt, perm);
            h->tm.kernel_launches += 1;
        }
        if (!bad && hc.n_direct) {
            const u64 pcap = next_pow2(std::max<u64>(1024, 2 * hc.n_direct));
            if (pcap > S->pos_cap) {
                cudaFree(S->posmap); S->posmap = nullptr; S->pos_cap = 0;
                E2CU(cudaMalloc(&S->posmap, pcap * sizeof(PosSlot)));
                S->pos_cap = pcap;
            }
            E2CU(cudaMemsetAsync(S->posmap, 0, pcap * sizeof(PosSlot), h->stream));
            E.posmap = S->posmap; E.pos_mask = pcap - 1;
            if (hc.n_direct > hc.n_long) {
                k_enc_direct_short<<<(unsigned)((hc.n_direct + 127) / 128), 128, 0, h->stream>>>(E, rt, perm);
                h->tm.kernel_launches += 1;
            }
            if (hc.n_long) {
                E2CU(cudaFuncSetAttribute(k_enc_direct_long, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ENC_LONG_MAX * 4));
                k_enc_direct_long<<<(unsigned)std::min<u64>(hc.n_direct, (u64)h->sms * 2), 256, 2 * ENC_LONG_MAX * 4, h->stream>>>(E, rt, perm);
                h->tm.kernel_launches += 1;
            }
        }
        u32 fail_bits = hc.fail;
        total = 0;
        if (!bad) {
            k_enc_count<<<ntiles, E2_THREADS, 0, h->stream>>>(E, S->part);
            k_flag_scan_parts<<<1, 1024, 0, h->stream>>>(S->part, S->excl, ntiles, S->d_total);
            h->tm.kernel_launches += 2;
            EncCtl hc2;
            E2CU(cudaMemcpyAsync(&total, S->d_total, 8, cudaMemcpyDeviceToHost, h->stream));
            E2CU(cudaMemcpyAsync(&hc2, S->ctl, sizeof(EncCtl), cudaMemcpyDeviceToHost, h->stream));
            E2CU(cudaStreamSynchronize(h->stream));
            fail_bits = hc2.fail;
            bad = fail_bits != 0;
            S->st_tmp_ids = hc2.tmp_used;
        }
        // only the per-piece area was too small (a chunk that went unresolved for that reason raised the other bit as
        // well: k_enc_count): every memo entry of this attempt is complete, so the piece can simply be done again
        if (bad && (fail_bits & E2_FAIL_TMP) && attempt == 0 && S->tmp_cap < m) {
            cudaFree(S->tmp); S->tmp = nullptr; S->tmp_cap = 0;
            E2CU(cudaMalloc(&S->tmp, m * 4));
            S->tmp_cap = m;
            S->st_retries += 1;
            continue;
        }
        break;
    }
    hc.n_new = (u32)n_new_all;
    S->st_chunks_new += hc.n_new; S->st_direct += hc.n_direct; S->st_long += hc.n_long; S->st_pieces += 1;
    if (bad) {
        // something did not fit (id pool, lists, an oversize chunk, a tag collision): start the memo afresh for
        // the pieces to come and let the caller run the general path on this one
        if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
        int rc = enc2_reset_memo(h);
        if (rc) return rc;
        S->st_fallbacks += 1;
        *fell_back = 1;
        return BPE_OK;
    }
    if (*written + total > cap) {
        if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
        return fail(h, BPE_ERR_CAPACITY, "output buffer too small");
    }
    if (total > S->ids_cap) {
        cudaFree(S->d_ids); S->d_ids = nullptr; S->ids_cap = 0;
        const u64 want = total + total / 8 + 1024;
        E2CU(cudaMalloc(&S->d_ids, want * 4));
        S->ids_cap = want;
    }
    if (total) {
        k_enc_write<<<ntiles, E2_THREADS, 0, h->stream>>>(E, S->excl, S->d_ids);
        h->tm.kernel_launches += 1;
    }
    if (e0) {
        cudaEventRecord(e1, h->stream); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); S->st_kernel_ms += ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (total) {
        E2CU(cudaGetLastError());
        E2CU(cudaMemcpyAsync(out + *written, S->d_ids, total * 4, cudaMemcpyDeviceToHost, h->stream));
        E2CU(cudaStreamSynchronize(h->stream));
        h->tm.d2h_bytes += total * 4;
    }
    *written += total;
    return BPE_OK;
}

// general path for one piece whose chunk-start flags are on the device: offsets to the host, then the round-1
// kernels on the scratch handle (thread per chunk / CTA per long chunk / stream rounds for one huge chunk)
static int enc2_piece_fallback(bpe_handle *h, const uint8_t *host_bytes, const unsigned char *d_flag, u64 m, const int32_t *merges,
                               int32_t n_merges, const uint8_t *perm, int32_t *out, u64 cap, u64 *written) {
    u64 *d_offs = nullptr;
    E2CU(cudaMalloc(&d_offs, m * 8));
    u64 k = 0;
    int rc = flags_to_offsets(h, d_flag, m, d_offs, &k, 0);
    std::vector<u64> offs;
    if (!rc) {
        offs.resize(k);
        cudaError_t e = cudaMemcpyAsync(offs.data(), d_offs, k * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_offs);
    if (rc) return rc;
    u64 got = 0;
    rc = encode_general(h, host_bytes, m, offs.data(), k, merges, n_merges, perm, out + *written, cap - *written, &got);
    if (!rc) *written += got;
    return rc;
}

// general path for one piece that contains special tokens: every part between two occurrences is encoded on its own
// (regex.py:159-163), the occurrences contribute their ids
static int enc2_piece_fallback_special(bpe_handle *h, const uint8_t *host_bytes, const unsigned char *d_flag, u64 m,
                                       const std::vector<u64> &hits, const std::vector<unsigned char> &which, const int32_t *merges,
                                       int32_t n_merges, const uint8_t *perm, int32_t *out, u64 cap, u64 *written) {
    u64 *d_offs = nullptr;
    E2CU(cudaMalloc(&d_offs, m * 8));
    u64 k = 0;
    int rc = flags_to_offsets(h, d_flag, m, d_offs, &k, 0);
    std::vector<u64> offs;
    if (!rc) {
        offs.resize(k);
        cudaError_t e = cudaMemcpyAsync(offs.data(), d_offs, k * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, BPE_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d_offs);
    if (rc) return rc;
    u64 lo = 0;
    std::vector<u64> rel;
    for (size_t j = 0; j <= hits.size(); ++j) {
        const u64 hi = j < hits.size() ? (hits[j] >> 8) : m;
        if (hi > lo) {      // ordinary part [lo, hi): its chunk starts are the flags inside it
            const u64 *f0 = std::lower_bound(offs.data(), offs.data() + k, lo), *f1 = std::lower_bound(offs.data(), offs.data() + k, hi);
            rel.assign(f0, f1);
            for (u64 &x : rel) x -= lo;
            u64 got = 0;
            if ((rc = encode_general(h, host_bytes + lo, hi - lo, rel.data(), rel.size(), merges, n_merges, perm, out + *written, cap - *written, &got))) return rc;
            *written += got;
        }
        if (j < hits.size()) {
            if (*written >= cap) return fail(h, BPE_ERR_CAPACITY, "output buffer too small");
            out[(*written)++] = h->spec->ids[which[j]];
            lo = hi + (hits[j] & 0xffu);
        }
    }
    return BPE_OK;
}

// regex.py:111-121 (and, with special tokens, regex.py:152-163) without the host in the middle: upload the text, find
// the special tokens, split with the GPT-4 pattern, encode every chunk, return the ids.  No offsets array anywhere.
static int encode_text_impl(bpe_handle *h, const uint8_t *bytes, uint64_t n, const int32_t *merges, int32_t n_merges,
                            const uint8_t *byte_perm, bool use_spec, int32_t *out_ids, uint64_t out_cap, uint64_t *out_n) {
    *out_n = 0;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0; h->tm.init_ms = 0;
    if (n == 0) return BPE_OK;
    use_spec = use_spec && h->spec && h->spec->k;
    int rc = enc2_prepare(h, merges, n_merges, byte_perm, use_spec);
    if (rc) return rc;
    EncState *S = h->enc;
    S->st_chunks_new = S->st_direct = S->st_long = S->st_pieces = S->st_fallbacks = S->st_retries = 0; S->st_kernel_ms = 0;
    const u64 piece = g_split_piece_override ? g_split_piece_override : SPLIT_PIECE_BYTES;
    u64 written = 0;
    std::vector<u64> hits;
    std::vector<unsigned char> which;
    for (u64 s = 0; s < n;) {
        const u64 e = split_piece_end(bytes, n, s, piece, use_spec ? h->spec : nullptr);
        if (!e) return fail(h, BPE_ERR_ARG, "no letter+space cut point within a piece of the text (cannot split it piecewise)");
        SplitWork W;
        if ((rc = split_run(h, bytes + s, e - s, W, nullptr, use_spec ? &hits : nullptr, use_spec ? &which : nullptr))) return rc;
        int fb = 0;
        if ((rc = enc2_piece(h, W.bytes, W.flag, e - s, out_ids, out_cap, &written, &fb))) return rc;
        if (fb) {
            if (use_spec && !hits.empty())
                rc = enc2_piece_fallback_special(h, bytes + s, W.flag, e - s, hits, which, merges, n_merges, byte_perm, out_ids, out_cap, &written);
            else
                rc = enc2_piece_fallback(h, bytes + s, W.flag, e - s, merges, n_merges, byte_perm, out_ids, out_cap, &written);
            if (rc) return rc;
        }
        s = e;
    }
    *out_n = written;
    return BPE_OK;
}

extern "C" int bpe_encode_text_gpt4(bpe_handle *h, const uint8_t *bytes, uint64_t n, const int32_t *merges, int32_t n_merges,
                                    const uint8_t *byte_perm, int32_t *out_ids, uint64_t out_cap, uint64_t *out_n) {
    if (!h || !out_n || (!bytes && n) || n_merges < 0 || (n_merges && !merges)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    return encode_text_impl(h, bytes, n, merges, n_merges, byte_perm, false, out_ids, out_cap, out_n);
}

// RegexTokenizer.encode(text, allowed_special=...) (regex.py:123-164) in one call: the special tokens (their utf-8 bytes
// back to back, k + 1 offsets, k ids; in the order of the special_tokens dict, which is the order the reference's regex
// tries them in) are found on the device and every part between them is split and encoded on its own.
extern "C" int bpe_encode_text_gpt4_special(bpe_handle *h, const uint8_t *bytes, uint64_t n, const int32_t *merges, int32_t n_merges,
                                            const uint8_t *byte_perm, const uint8_t *special_bytes, const uint32_t *special_offsets,
                                            const int32_t *special_ids, int32_t n_special, int32_t *out_ids, uint64_t out_cap,
                                            uint64_t *out_n) {
    if (!h || !out_n || (!bytes && n) || n_merges < 0 || (n_merges && !merges)) return BPE_ERR_ARG;
    CU(cudaSetDevice(h->device));
    int rc = spec_set(h, special_bytes, special_offsets, special_ids, n_special);
    if (rc) return rc;
    return encode_text_impl(h, bytes, n, merges, n_merges, byte_perm, true, out_ids, out_cap, out_n);
}

// bpe_encode with the caller's chunk offsets (any split pattern): text + offsets up, flags from the offsets
static int encode_with_offsets(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *offs, uint64_t n_chunks,
                               const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm, int32_t *out_ids, uint64_t out_cap,
                               uint64_t *out_n) {
    int rc = enc2_prepare(h, merges, n_merges, byte_perm);
    if (rc) return rc;
    EncState *S = h->enc;
    S->st_chunks_new = S->st_direct = S->st_long = S->st_pieces = S->st_fallbacks = S->st_retries = 0; S->st_kernel_ms = 0;
    const u64 piece = g_split_piece_override ? g_split_piece_override : SPLIT_PIECE_BYTES;
    u64 written = 0;
    u64 c0 = 0;
    while (c0 < n_chunks) {
        // chunks [c0, c1) : at most `piece` bytes (a single chunk beyond that goes alone)
        const u64 lo = offs[c0];
        u64 c1 = (u64)(std::upper_bound(offs + c0, offs + n_chunks, lo + piece) - offs);
        if (c1 <= c0 + 1) c1 = c0 + 1;
        const u64 hi = c1 < n_chunks ? offs[c1] : n;
        const u64 m = hi - lo, k = c1 - c0;
        if (m >= 0xfffffff0ull) {   // one chunk of 4 GiB: the general path handles it
            u64 got = 0;
            if ((rc = encode_general(h, bytes + lo, m, nullptr, 0, merges, n_merges, byte_perm, out_ids + written, out_cap - written, &got))) return rc;
            written += got; c0 = c1;
            continue;
        }
        SplitWork W;
        if ((rc = split_carve(h, m, W))) return rc;
        if (k > S->offs_cap) {
            cudaFree(S->d_offs); S->d_offs = nullptr; S->offs_cap = 0;
            E2CU(cudaMalloc(&S->d_offs, k * 8));
            S->offs_cap = k;
        }
        E2CU(cudaMemcpyAsync(W.bytes, bytes + lo, m, cudaMemcpyHostToDevice, h->stream));
        E2CU(cudaMemsetAsync(W.flag, 0, m, h->stream));
        {   // offsets are relative to the whole text: shift while copying would need a kernel; upload and subtract there
            std::vector<u64> rel(k);
            for (u64 i = 0; i < k; ++i) rel[i] = offs[c0 + i] - lo;
            E2CU(cudaMemcpyAsync(S->d_offs, rel.data(), k * 8, cudaMemcpyHostToDevice, h->stream));
            k_enc_flags_from_offsets<<<grid_for(k, 256, h->sms * 8), 256, 0, h->stream>>>(W.flag, S->d_offs, k, m);
            E2CU(cudaStreamSynchronize(h->stream));   // rel goes out of scope
        }
        h->tm.h2d_bytes += m + k * 8;
        h->tm.kernel_launches += 1;
        int fb = 0;
        if ((rc = enc2_piece(h, W.bytes, W.flag, m, out_ids, out_cap, &written, &fb))) return rc;
        if (fb) {
            u64 got = 0;
            std::vector<u64> rel(k);
            for (u64 i = 0; i < k; ++i) rel[i] = offs[c0 + i] - lo;
            if ((rc = encode_general(h, bytes + lo, m, rel.data(), k, merges, n_merges, byte_perm, out_ids + written, out_cap - written, &got))) return rc;
            written += got;
        }
        c0 = c1;
    }
    *out_n = written;
    return BPE_OK;
}

extern "C" int bpe_encode_stats(bpe_handle *h, uint64_t *out /* [10] */) {
    if (!h || !out) return BPE_ERR_ARG;
    memset(out, 0, 10 * sizeof(uint64_t));
    if (!h->enc) return BPE_OK;
    EncState *S = h->enc;
    EncCtl hc;
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(&hc, S->ctl, sizeof(EncCtl), cudaMemcpyDeviceToHost));
    out[0] = hc.memo_used; out[1] = hc.pool_used; out[2] = S->st_chunks_new; out[3] = S->st_direct; out[4] = S->st_long;
    out[5] = S->st_pieces; out[6] = S->st_fallbacks; out[7] = (uint64_t)(S->st_kernel_ms * 1000.0);
    out[8] = S->st_retries; out[9] = S->st_tmp_ids;
    return BPE_OK;
}

// regex.py:92-121 / basic.py:57-74.  Chunked input (RegexTokenizer, any pattern: the caller brings the offsets) of at
// least ENC2_MIN_BYTES goes through the memoised kernels; short inputs (a chunk, a sentence: the memo's 0.5 GB of state
// and its reset on every change of merges buy nothing there) and a single chunk (BasicTokenizer) take the general path.
#define ENC2_MIN_BYTES (1u << 16)
extern "C" int bpe_encode(bpe_handle *h, const uint8_t *bytes, uint64_t n, const uint64_t *chunk_offsets,
                          uint64_t n_chunks, const int32_t *merges, int32_t n_merges, const uint8_t *byte_perm,
                          int32_t *out_ids, uint64_t out_cap, uint64_t *out_n) {
    if (!h || !out_n) return BPE_ERR_ARG;
    if (!bytes && n) return fail(h, BPE_ERR_ARG, "bytes is NULL");
    if (n_merges < 0 || (n_merges && !merges)) return fail(h, BPE_ERR_ARG, "bad merges");
    if (n >= (1ull << 36)) return fail(h, BPE_ERR_ARG, "input too long");
    CU(cudaSetDevice(h->device));
    int rc = check_offsets(h, chunk_offsets, n_chunks, n);   // every path indexes the text through these
    if (rc) return rc;
    *out_n = 0;
    h->tm.h2d_bytes = 0; h->tm.d2h_bytes = 0; h->tm.kernel_launches = 0;
    if (n >= ENC2_MIN_BYTES && n_merges > 0 && chunk_offsets && n_chunks >= 1)
        return encode_with_offsets(h, bytes, n, chunk_offsets, n_chunks, merges, n_merges, byte_perm, out_ids, out_cap, out_n);
    return encode_general(h, bytes, n, chunk_offsets, n_chunks, merges, n_merges, byte_perm, out_ids, out_cap, out_n);
}
