// oracle/split_harness.cpp — TEST INFRASTRUCTURE.  Runs minbpe_b200/csrc/split_logic.h (the scan
// operators and the chunk-start rule the CUDA splitter is built from) on the CPU, with the scans
// evaluated TILE BY TILE exactly as the kernels decompose them (tile aggregates, exclusive scan of the
// aggregates, in-tile scans seeded with the carries), so that tests/test_split_rules.py can pin the
// product's rule code against the `regex` module without a GPU.
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "../minbpe_b200/csrc/split_logic.h"

struct ByteAcc {
    const uint8_t *p;
    uint32_t operator()(uint64_t i) const { return p[i]; }
};

// hits[j] = position << 8 | length of a special-token occurrence (WITH_B only): its bytes get the boundary class SC_B
// before the scans, exactly as k_special_meta does, and the flags are fixed up as k_special_flags does.
template <bool WITH_B, int PATTERN = 0>   // PATTERN: 0 = GPT-4, 1 = GPT-2 (split_logic.h)
static int flags_impl(const uint8_t *bytes, uint64_t n, const uint8_t *cls_table, const uint8_t *contr, uint64_t tile,
                      const uint64_t *hits, uint64_t n_hits, uint8_t *flags) {
    if (n == 0) return 0;
    if (tile == 0) tile = n;
    const ByteAcc B{bytes};
    std::vector<uint8_t> meta(n);
    for (uint64_t i = 0; i < n; ++i) meta[i] = (uint8_t)spl_meta_of(B, i, n, cls_table);
    if (WITH_B)
        for (uint64_t j = 0; j < n_hits; ++j)
            for (uint64_t t = 0; t < (hits[j] & 0xff); ++t) { uint8_t &m = meta[(hits[j] >> 8) + t]; m = (uint8_t)((m & SM_START) | SC_B); }
    const ByteAcc M{meta.data()};
    const uint64_t ntiles = (n + tile - 1) / tile;
    // 1. tile aggregates
    std::vector<SplFwd> fagg(ntiles);
    std::vector<SplBwd> gagg(ntiles);
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
        SplFwd f = spl_fwd_identity();
        for (uint64_t i = lo; i < hi; ++i) f = spl_fwd_combine(f, spl_fwd_elem(i, meta[i], i ? meta[i - 1] : 0));
        SplBwd g = spl_bwd_identity();
        for (uint64_t i = hi; i > lo; --i) g = spl_bwd_combine(spl_bwd_elem<WITH_B>(i - 1, n, meta[i - 1], i < n ? meta[i] : 0), g);
        fagg[t] = f; gagg[t] = g;
    }
    // 2. exclusive scans of the aggregates: forward left to right, backward right to left
    std::vector<SplFwd> fcar(ntiles);
    std::vector<SplBwd> gcar(ntiles);
    SplFwd frun = spl_fwd_identity();
    for (uint64_t t = 0; t < ntiles; ++t) { fcar[t] = frun; frun = spl_fwd_combine(frun, fagg[t]); }
    SplBwd grun = spl_bwd_identity();
    for (uint64_t t = ntiles; t > 0; --t) { gcar[t - 1] = grun; grun = spl_bwd_combine(gagg[t - 1], grun); }
    // 3. in-tile scans seeded with the carries, then the rule
    std::vector<SplFwd> fv(tile);
    std::vector<SplBwd> gv(tile);
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t lo = t * tile, hi = (lo + tile < n) ? lo + tile : n;
        SplFwd f = fcar[t];
        for (uint64_t i = lo; i < hi; ++i) { f = spl_fwd_combine(f, spl_fwd_elem(i, meta[i], i ? meta[i - 1] : 0)); fv[i - lo] = f; }
        SplBwd g = gcar[t];
        for (uint64_t i = hi; i > lo; --i) { g = spl_bwd_combine(spl_bwd_elem<WITH_B>(i - 1, n, meta[i - 1], i < n ? meta[i] : 0), g); gv[i - 1 - lo] = g; }
        for (uint64_t i = lo; i < hi; ++i) {
            const SplFwd fprev = (i == lo) ? fcar[t] : fv[i - 1 - lo];
            if (PATTERN == 1) flags[i] = (meta[i] & SM_START) && spl_chunk_start_gpt2<WITH_B>(i, n, fv[i - lo], gv[i - lo], B, M) ? 1 : 0;
            else flags[i] = (meta[i] & SM_START) && spl_chunk_start<WITH_B>(i, n, fv[i - lo], fprev, gv[i - lo], B, M, contr) ? 1 : 0;
        }
    }
    if (WITH_B)
        for (uint64_t j = 0; j < n_hits; ++j) {
            const uint64_t pos = hits[j] >> 8, len = hits[j] & 0xff;
            flags[pos] = 1;
            for (uint64_t t = 1; t < len; ++t) flags[pos + t] = 0;
            if (pos + len < n) flags[pos + len] = 1;
        }
    return 0;
}

extern "C" int split_logic_flags(const uint8_t *bytes, uint64_t n, const uint8_t *cls_table, const uint8_t *contr,
                                 uint64_t tile, uint8_t *flags) {
    return flags_impl<false>(bytes, n, cls_table, contr, tile, nullptr, 0, flags);
}

// the same with special-token occurrences as text boundaries (regex.py:152-163; k_special.cuh + the WITH_B rule code)
extern "C" int split_logic_flags_special(const uint8_t *bytes, uint64_t n, const uint8_t *cls_table, const uint8_t *contr,
                                         uint64_t tile, const uint64_t *hits, uint64_t n_hits, uint8_t *flags) {
    return flags_impl<true>(bytes, n, cls_table, contr, tile, hits, n_hits, flags);
}

// GPT-2 pattern (regex.py:18), plain and with special-token boundaries
extern "C" int split_logic_flags_gpt2(const uint8_t *bytes, uint64_t n, const uint8_t *cls_table, const uint8_t *contr,
                                      uint64_t tile, const uint64_t *hits, uint64_t n_hits, uint8_t *flags) {
    if (n_hits) return flags_impl<true, 1>(bytes, n, cls_table, contr, tile, hits, n_hits, flags);
    return flags_impl<false, 1>(bytes, n, cls_table, contr, tile, nullptr, 0, flags);
}
