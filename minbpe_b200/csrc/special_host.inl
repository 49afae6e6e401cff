// special_host.inl — host side of the special-token front end (k_special.cuh): the set of specials of the current
// encode call, the search for their occurrences in a piece of text on the device, and the left-to-right selection of
// the occurrences re.split would report (regex.py:152-154).

struct SpecSet {
    int k = 0;
    u64 hash = 0;
    std::vector<unsigned char> blob;
    std::vector<u32> off;            // [k + 1]
    std::vector<int> ids;            // [k]
    unsigned char *d_blob = nullptr; u32 *d_off = nullptr; u64 *d_first = nullptr; int *d_ids = nullptr;
    u64 *d_list = nullptr; u64 list_cap = 0; ull *d_count = nullptr;   // candidates of one piece
    u64 *d_hit = nullptr; u64 hit_cap = 0;                              // accepted occurrences of one piece
};

static void spec_free(bpe_handle *h) {
    SpecSet *S = h->spec;
    if (!S) return;
    cudaFree(S->d_blob); cudaFree(S->d_off); cudaFree(S->d_first); cudaFree(S->d_ids); cudaFree(S->d_list); cudaFree(S->d_count);
    cudaFree(S->d_hit);
    delete S;
    h->spec = nullptr;
}

static u64 spec_fnv(const void *p, size_t n, u64 h = 0xcbf29ce484222325ull) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

// install the specials of this call (k == 0: none).  Order = the order the reference's regex tries them in
// (dict order of special_tokens, regex.py:152).
static int spec_set(bpe_handle *h, const uint8_t *bytes, const uint32_t *offsets, const int32_t *ids, int32_t k) {
    if (k < 0 || (k && (!bytes || !offsets || !ids))) return fail(h, BPE_ERR_ARG, "bad special-token arguments");
    if (k > SPEC_MAX) return fail(h, BPE_ERR_ARG, "more than 64 special tokens in one call");
    if (k && offsets[0] != 0) return fail(h, BPE_ERR_ARG, "special_offsets[0] must be 0");
    for (int s = 0; s < k; ++s) {
        if (offsets[s + 1] <= offsets[s]) return fail(h, BPE_ERR_ARG, "special tokens must not be empty");
        if (offsets[s + 1] - offsets[s] > SPEC_MAX_LEN) return fail(h, BPE_ERR_ARG, "special tokens longer than 48 bytes are not handled on the device");
        if (ids[s] < 0) return fail(h, BPE_ERR_ARG, "special-token ids must be >= 0");
    }
    if (!h->spec) h->spec = new (std::nothrow) SpecSet();
    SpecSet *S = h->spec;
    if (!S) return fail(h, BPE_ERR_INTERNAL, "out of host memory");
    u64 hash = 0;
    if (k) {
        hash = spec_fnv(bytes, offsets[k]);
        hash = spec_fnv(offsets, (size_t)(k + 1) * 4, hash);
        hash = spec_fnv(ids, (size_t)k * 4, hash) | 1ull;
    }
    if (hash == S->hash && k == S->k) return BPE_OK;
    S->k = k; S->hash = 0;                              // becomes `hash` once the device copies are in place
    S->blob.assign(bytes, bytes + (k ? offsets[k] : 0));
    S->off.assign(offsets, offsets + (k ? k + 1 : 0));
    S->ids.assign(ids, ids + k);
    if (!k) return BPE_OK;                              // k == 0 <=> hash == 0
    if (!S->d_blob) {
        CU(cudaMalloc(&S->d_blob, SPEC_MAX * SPEC_MAX_LEN));
        CU(cudaMalloc(&S->d_off, (SPEC_MAX + 1) * 4));
        CU(cudaMalloc(&S->d_first, 256 * 8));
        CU(cudaMalloc(&S->d_ids, SPEC_MAX * 4));
        CU(cudaMalloc(&S->d_count, 8));
    }
    u64 first[256];
    memset(first, 0, sizeof(first));
    for (int s = 0; s < k; ++s) first[S->blob[S->off[s]]] |= 1ull << s;
    CU(cudaMemcpyAsync(S->d_blob, S->blob.data(), S->blob.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(S->d_off, S->off.data(), (size_t)(k + 1) * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(S->d_ids, S->ids.data(), (size_t)k * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(S->d_first, first, sizeof(first), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // `first` is a stack buffer
    S->hash = hash;
    return BPE_OK;
}

static SpecDev spec_dev(const SpecSet *S) { return SpecDev{S->d_blob, S->d_off, S->d_first, S->k}; }

// Occurrences of the specials in the m text bytes at d_text that re.split reports: candidates from the device, sorted
// by position, overlapped ones dropped left to right.  hits[j] = position << 8 | length, which[j] = index of the special.
static int spec_find(bpe_handle *h, const unsigned char *d_text, u64 m, std::vector<u64> &hits, std::vector<unsigned char> &which) {
    SpecSet *S = h->spec;
    hits.clear(); which.clear();
    if (!S || !S->k || !m) return BPE_OK;
    for (;;) {
        if (!S->d_list) {
            S->list_cap = std::max<u64>(S->list_cap, 1ull << 16);
            CU(cudaMalloc(&S->d_list, S->list_cap * 8));
        }
        CU(cudaMemsetAsync(S->d_count, 0, 8, h->stream));
        k_special_find<<<grid_for(m, 256, h->sms * 8), 256, 0, h->stream>>>(d_text, m, spec_dev(S), S->d_list, S->list_cap, S->d_count);
        h->tm.kernel_launches += 1;
        ull cnt = 0;
        CU(cudaMemcpyAsync(&cnt, S->d_count, 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (cnt <= S->list_cap) {
            std::vector<u64> cand(cnt);
            if (cnt) {
                CU(cudaMemcpyAsync(cand.data(), S->d_list, cnt * 8, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                h->tm.d2h_bytes += cnt * 8;
            }
            std::sort(cand.begin(), cand.end());
            u64 next_free = 0;
            for (u64 c : cand) {
                const u64 pos = c >> 8;
                const u32 s = (u32)(c & 0xffu), len = S->off[s + 1] - S->off[s];
                if (pos < next_free) continue;            // inside the previous match: the regex never looks here
                hits.push_back((pos << 8) | len);
                which.push_back((unsigned char)s);
                next_free = pos + len;
            }
            return BPE_OK;
        }
        cudaFree(S->d_list); S->d_list = nullptr;         // more candidates than room: grow and search again
        S->list_cap = cnt + cnt / 4 + 1024;
    }
}

// upload the accepted occurrences of a piece (spec_find) for k_special_meta / k_special_flags
static int spec_upload_hits(bpe_handle *h, const std::vector<u64> &hits) {
    SpecSet *S = h->spec;
    if (hits.size() > S->hit_cap) {
        cudaFree(S->d_hit); S->d_hit = nullptr; S->hit_cap = 0;
        const u64 want = hits.size() + hits.size() / 4 + 1024;
        CU(cudaMalloc(&S->d_hit, want * 8));
        S->hit_cap = want;
    }
    CU(cudaMemcpyAsync(S->d_hit, hits.data(), hits.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));                 // `hits` is the caller's vector
    h->tm.h2d_bytes += hits.size() * 8;
    return BPE_OK;
}

// does position p (0 < p < n) lie strictly inside an occurrence of some special in the host text?  (Conservative
// for piece cuts: an occurrence the regex would not report, because an earlier match overlaps it, also counts.)
static bool spec_covers(const SpecSet *S, const uint8_t *b, u64 n, u64 p) {
    if (!S || !S->k) return false;
    for (int s = 0; s < S->k; ++s) {
        const u32 lo = S->off[s], len = S->off[s + 1] - lo;
        for (u32 j = 1; j < len; ++j) {                   // the occurrence would start at p - j
            if (p < j || p - j + len > n) continue;
            if (memcmp(b + (p - j), S->blob.data() + lo, len) == 0) return true;
        }
    }
    return false;
}
