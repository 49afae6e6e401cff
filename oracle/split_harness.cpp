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

extern "C" int split_logic_flags(const uint8_t *bytes, uint64_t n, const uint8_t *cls_table, const uint8_t *contr,
                                 uint64_t tile, uint8_t *flags) {
    if (n == 0) return 0;
    if (tile == 0) tile = n;
    const ByteAcc B{bytes};
    std::vector<uint8_t> meta(n);
    for (uint64_t i = 0; i < n; ++i) meta[i] = (uint8_t)spl_meta_of(B, i, n, cls_table);
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
        for (uint64_t i = hi; i > lo; --i) g = spl_bwd_combine(spl_bwd_elem(i - 1, n, meta[i - 1], i < n ? meta[i] : 0), g);
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
        for (uint64_t i = hi; i > lo; --i) { g = spl_bwd_combine(spl_bwd_elem(i - 1, n, meta[i - 1], i < n ? meta[i] : 0), g); gv[i - 1 - lo] = g; }
        for (uint64_t i = lo; i < hi; ++i) {
            const SplFwd fprev = (i == lo) ? fcar[t] : fv[i - 1 - lo];
            flags[i] = (meta[i] & SM_START) && spl_chunk_start(i, n, fv[i - lo], fprev, gv[i - lo], B, M, contr) ? 1 : 0;
        }
    }
    return 0;
}
