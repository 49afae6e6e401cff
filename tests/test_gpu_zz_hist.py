"""GPU: the packed dense byte-pair histogram (k_hist_dense_packed) is cross-checked against k_hist_dense on a handle's
first use (b200bpe.cu hist_dense) and must have been adopted: bpe_timing.hist_kernel == 1.  (If this fails on a B200 the
library has fallen back to k_hist_dense on its own — results are unaffected — and the packed kernel has a bug the CPU
emulator could not see.)"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_packed_histogram_agrees_and_is_in_use():
    from minbpe_b200.engine import Engine
    from minbpe_b200.synth import generate
    eng = Engine(0)
    raw = generate(1337, 8 << 20)
    for n, merges in ((8 << 20, 6), (3 << 20, 6), (70001, 4)):
        eng.load_stream(raw[:n], None)                      # one chunk: every adjacent pair counts
        p, c, d = eng.train(merges)
        assert eng.timing()["hist_kernel"] == 1
        w = oracle.c_train(raw[:n].astype(np.int32), None, merges)
        assert d == merges and np.array_equal(p, w[0]) and np.array_equal(c, w[1])
    eng.close()
