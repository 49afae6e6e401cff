"""GPU: the segment filter of the training loop (BPE_OPT_SEG_FILTER; k_seg_filter.cuh + k_merge_seg<true>): per-segment id
signatures decide which segments a merge can touch at all, the merge pass then works through that list only.  Merges and
counts must be those of the oracle whether the filter is off, switched on when merges become sparse, or on from the start —
over enough merges to cross re-packing, pairs (a,a) and table growth."""
import os

import numpy as np
import pytest
import regex

import oracle

pytestmark = pytest.mark.gpu
SMALL = bool(os.environ.get("BPE_TEST_SMALL"))      # set by tests/test_emu.py (CPU SIMT emulator): smaller corpus, fewer merges

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


@pytest.mark.parametrize("mode", [2, 1])
def test_filtered_training_equals_the_oracle(mode):
    from minbpe_b200 import engine as E
    from minbpe_b200.synth import generate
    text = generate(1337, (1 if SMALL else 2) << 20).tobytes().decode("utf-8") + " zz zz zzzz zzzzzz aaaa aaaa"
    data, offs = oracle.split_to_stream(text, GPT4)
    # mode 1 switches the filter on only once merges have become sparse (about merge 2,700 of this corpus)
    merges = (800 if SMALL else 1500) if mode == 2 else (3000 if SMALL else 3600)
    ub, uo, uw = oracle.c_dedup_chunks(data, offs)
    wp, wc, wd = oracle.c_train(ub.astype(np.int32), uo, merges, weights=uw)
    eng = E.Engine(0)
    eng.set_option(E.OPT_SEG_FILTER, mode)
    eng.set_option(E.OPT_BATCH, 64)                 # more batch boundaries: re-packing and the switch-on happen there
    eng.load_stream(data, offs)
    p, c, d = eng.train(merges)
    tm = eng.timing()
    assert d == wd == merges and np.array_equal(p, wp) and np.array_equal(c, wc)
    assert tm["filter_segments"] > 0 and tm["filter_candidates"] < tm["filter_segments"]      # it ran, and it skipped segments
    # continuing the run in a second call (table kept, signatures rebuilt) and the table itself
    p2, c2, d2 = eng.train(100, first_idx=256 + merges)
    w2 = oracle.c_train(ub.astype(np.int32), uo, merges + 100, weights=uw)
    assert d2 == 100 and np.array_equal(p2, w2[0][merges:]) and np.array_equal(c2, w2[1][merges:])
    sp, sc = eng.get_stats()
    assert eng.debug_table() == {(int(a), int(b)): int(n) for (a, b), n in zip(sp, sc)}
    with pytest.raises(E.EngineError):
        eng.set_option(E.OPT_SEG_FILTER, 3)
    eng.close()


def test_filter_on_basic_tokenizer_stream_and_tiny_inputs():
    from minbpe_b200 import engine as E
    eng = E.Engine(0)
    eng.set_option(E.OPT_SEG_FILTER, 2)
    for text, merges in (("aaabdaaabac", 3), ("ab" * 3000 + "cd" * 2000 + "abcd" * 1500, 40), ("x", 1), ("hello world " * 700, 30)):
        raw = np.frombuffer(text.encode(), dtype=np.uint8)
        eng.load_stream(raw, None)
        p, c, d = eng.train(merges)
        w = oracle.c_train(raw.astype(np.int32), None, merges)
        assert d == w[2] and np.array_equal(p, w[0]) and np.array_equal(c, w[1]), text[:20]
    eng.close()
