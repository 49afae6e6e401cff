"""
GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI (ctypes) and through
the reference-shaped Python classes, against
  * the golden vectors generated from the unmodified reference (tests/golden/*.json),
  * the CPU oracle (oracle/) on the same seeded inputs,
  * size-independent properties at sizes the oracle cannot reach (expanding the merged stream
    through the vocab reproduces the corpus; incremental pair table == recount of the stream).
Bit-exact everywhere: this is integer / index work.
"""
import contextlib
import hashlib
import io
import os

import numpy as np
import pytest
import regex

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def ids_sha(ids):
    return sha(np.asarray(ids, dtype="<i4").tobytes())


@pytest.fixture(scope="module")
def eng():
    from minbpe_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def make(kind):
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    return BasicTokenizer() if kind == "basic" else RegexTokenizer()


# ---------------------------------------------------------------------------------------------
# reference known-answer test + golden configs through the reference-shaped classes

@pytest.mark.parametrize("kind", ["basic", "regex"])
def test_wikipedia_example(kind):
    # reference tests/test_tokenizer.py:80-107
    tok = make(kind)
    text = "aaabdaaabac"
    tok.train(text, 256 + 3)
    assert tok.merges == {(97, 97): 256, (256, 97): 257, (257, 98): 258}
    assert [tok.vocab[i] for i in (256, 257, 258)] == [b"aa", b"aaa", b"aaab"]
    ids = tok.encode(text)
    assert ids == [258, 100, 258, 97, 99]
    assert tok.decode(tok.encode(text)) == text


@pytest.mark.parametrize("kind", ["basic", "regex"])
def test_taylorswift_512(golden_train, taylorswift, tmp_path, kind):
    """BASELINE.json configs[1]: bit-exact merges/vocab/ids/.model/.vocab/verbose output."""
    g = golden_train[f"taylorswift_{kind}_512"]
    tok = make(kind)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tok.train(taylorswift, 512, verbose=True)
    assert [list(p) for p in tok.merges] == g["merges"]
    assert sha("\n".join(f"{a} {b}" for a, b in tok.merges).encode()) == g["merges_sha256"]
    assert sha(buf.getvalue().encode()) == g["verbose_sha256"]  # includes every stats[pair] count
    ids = tok.encode(taylorswift)
    assert len(ids) == g["n_ids"] and ids[:64] == g["ids_head"] and ids_sha(ids) == g["ids_sha256"]
    assert tok.decode(ids) == taylorswift
    prefix = str(tmp_path / kind)
    tok.save(prefix)
    assert sha(open(prefix + ".model", "rb").read()) == g["model_sha256"]
    assert sha(open(prefix + ".vocab", "rb").read()) == g["vocab_sha256"]


def test_save_load_specials(golden_train):
    # reference tests/test_tokenizer.py:109-132
    from minbpe_b200 import RegexTokenizer
    g = golden_train["llama_regex_320_specials"]
    text = open(os.path.join(GOLDEN, "llama_text.txt"), encoding="utf-8").read()
    tok = RegexTokenizer()
    tok.train(text, 256 + 64)
    tok.register_special_tokens(g["specials"])
    assert [list(p) for p in tok.merges] == g["merges"]
    assert tok.decode(tok.encode(text, "all")) == text
    ids = tok.encode(text, "all")
    assert ids == g["ids_all"]
    assert tok.encode(text, "none") == g["ids_none"]
    assert tok.encode(text, {"<|endoftext|>", "<|fim_suffix|>"}) == g["ids_subset"]
    tok.save("test_tokenizer_tmp")
    tok = RegexTokenizer()
    tok.load("test_tokenizer_tmp.model")
    assert tok.decode(ids) == text
    assert tok.encode(text, "all") == ids
    for f in ("test_tokenizer_tmp.model", "test_tokenizer_tmp.vocab"):
        os.remove(f)


def test_small_cases(golden_cases):
    """Runs, ties, overlaps, unicode, chunk isolation, exhaustion (ValueError)."""
    for c in golden_cases:
        kind, text, V = c["tokenizer"], c["text"], c["vocab_size"]
        tok = make(kind)
        if "raises" in c:
            with pytest.raises(ValueError):
                tok.train(text, V)
            assert tok.merges == {}  # the reference dies before assigning self.merges
            if c["n_done"]:
                tok.train(text, 256 + c["n_done"])
                assert [list(p) for p in tok.merges] == c["merges"]
            continue
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            tok.train(text, V, verbose=True)
        assert [list(p) for p in tok.merges] == c["merges"], (kind, text)
        counts = [int(x) for x in regex.findall(r"had (\d+) occurrences", buf.getvalue())]
        assert counts == c["counts"], (kind, text)
        for t, want in c["encode"].items():
            assert tok.encode(t) == want, (kind, text, t)
            assert tok.decode(want) == t


def test_primitives(golden_primitives):
    """Module-level get_stats / merge (base.py:13-41) incl. dict insertion order."""
    from minbpe_b200 import get_stats, merge
    for g in golden_primitives["get_stats"]:
        assert [[a, b, c] for (a, b), c in get_stats(g["ids"]).items()] == g["items"]
    for g in golden_primitives["merge"]:
        assert merge(g["ids"], tuple(g["pair"]), g["idx"]) == g["out"]
    acc = {}
    for L in golden_primitives["get_stats_accumulate"]["lists"]:
        get_stats(L, acc)
    assert [[a, b, c] for (a, b), c in acc.items()] == golden_primitives["get_stats_accumulate"]["items"]


@pytest.mark.parametrize("kind,V", [("basic", 300), ("regex", 320)])
def test_synth_prefix_golden(golden_train, kind, V):
    from minbpe_b200.synth import generate
    g = golden_train[f"synth1337_256k_{kind}_{V}"]
    text = generate(1337, 256 * 1024).tobytes().decode("utf-8")
    tok = make(kind)
    tok.train(text, V)
    assert [list(p) for p in tok.merges] == g["merges"]
    if g["ids_sha256"]:
        ids = tok.encode(text)
        assert len(ids) == g["n_ids"] and ids_sha(ids) == g["ids_sha256"]


# ---------------------------------------------------------------------------------------------
# C-ABI level differential tests against the oracle

def random_stream(rng, n, alphabet, p_chunk):
    ids = rng.choice(np.asarray(alphabet, dtype=np.int32), size=n)
    if p_chunk > 0 and n > 1:
        cuts = np.flatnonzero(rng.random(n) < p_chunk)
        offs = np.unique(np.concatenate([[0], cuts])).astype(np.uint64)
    else:
        offs = None
    return ids.astype(np.int32), offs


SIZES = [0, 1, 2, 3, 31, 32, 33, 127, 128, 129, 1023, 1024, 1025, 4093, 4094, 4095, 4096, 4097, 4098, 4099,
         8191, 8192, 8193, 12289, 40000, 131072 + 5]


@pytest.mark.parametrize("n", SIZES)
def test_merge_and_stats_vs_oracle(eng, n):
    """bpe_merge / bpe_get_stats on sizes around warp-row (128) and tile (4096) boundaries, with
    and without chunk marks, pairs with a != b and a == b (run-parity rule)."""
    rng = np.random.default_rng(1000 + n)
    for alphabet, p_chunk in (([7, 8], 0.0), ([7, 8], 0.05), ([7], 0.0), ([7], 0.01), ([1, 2, 3, 300], 0.2), ([5, 5, 5, 9], 0.002)):
        ids, offs = random_stream(rng, n, alphabet, p_chunk)
        for pair in ((7, 8), (7, 7), (5, 5), (300, 1), (8, 7)):
            eng.load_ids(ids, offs)
            new_len = eng.merge(pair[0], pair[1], 999)
            want = oracle.c_merge(ids, pair, 999, offs)
            got = eng.read_stream()
            assert new_len == len(want)
            assert np.array_equal(got, want), (n, alphabet, p_chunk, pair)
        eng.load_ids(ids, offs)
        gp, gc = eng.get_stats()
        wp, wc = oracle.c_get_stats(ids, offs)
        assert np.array_equal(gp, wp) and np.array_equal(gc, wc), (n, alphabet, p_chunk)


def test_long_runs_across_tiles(eng):
    """a == a runs spanning many 4096-token tiles, starting at every alignment, with chunk marks
    inside the run: exercises the block-wide backward walk and the row/tile parity carry."""
    for lead in (0, 1, 2, 3, 4093, 4095, 4096, 4097):
        for run in (1, 2, 5, 4096, 4097, 8192 + 3, 30001):
            ids = np.concatenate([np.full(lead, 3), np.full(run, 7), np.array([9, 7, 7, 7, 9, 7, 7])]).astype(np.int32)
            for offs in (None, np.array([0, lead + run // 2], dtype=np.uint64) if lead + run // 2 > 0 else None,
                         np.array([0] + list(range(max(1, lead), lead + run, 1000)), dtype=np.uint64)):
                if offs is not None:
                    offs = np.unique(offs)
                eng.load_ids(ids, offs)
                eng.merge(7, 7, 500)
                want = oracle.c_merge(ids, (7, 7), 500, offs)
                assert np.array_equal(eng.read_stream(), want), (lead, run, None if offs is None else offs[:4])


def train_both(eng, data, offs, num_merges):
    eng.load_stream(data, offs)
    gp, gc, gdone = eng.train(num_merges)
    wp, wc, wdone = oracle.c_train(np.frombuffer(bytes(data), dtype=np.uint8).astype(np.int32), offs, num_merges)
    return (gp, gc, gdone), (wp, wc, wdone)


@pytest.mark.parametrize("seed", range(6))
def test_train_random_vs_oracle(eng, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 60000))
    alphabet = [ord(c) for c in ("ab", "abc ", "aab\n", "abcdefgh ", "a", "ab")[seed % 6]]
    data = rng.choice(np.asarray(alphabet, dtype=np.uint8), size=n).astype(np.uint8)
    offs = None
    if seed % 2:
        offs = np.unique(np.concatenate([[0], np.flatnonzero(rng.random(n) < 0.1)])).astype(np.uint64)
    got, want = train_both(eng, data.tobytes(), offs, 120)
    assert got[2] == want[2]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_incremental_table_equals_recount(eng, taylorswift):
    """After k merges the incrementally maintained pair table must equal get_stats() of the
    current stream (every pair, every count) — checked against the oracle's recount."""
    from minbpe_b200.tokenizer import split_text
    data, offs = split_text(GPT4, taylorswift[:60000])
    for k in (1, 2, 3, 10, 57):
        eng.load_stream(data, offs)
        pairs, counts, done = eng.train(k)
        assert done == k
        table = eng.debug_table()
        ids = np.frombuffer(bytes(data), dtype=np.uint8).astype(np.int32)
        _, _, _, final = oracle.c_train(ids, offs, k, want_final=True)
        # chunk starts of the final stream: a chunk keeps its first token through every merge
        cur = eng.read_stream()
        assert np.array_equal(cur, final)
        # recount on the device too (full-histogram kernel, first-occurrence order)
        gp, gc = eng.get_stats()
        recount = {(int(a), int(b)): int(c) for (a, b), c in zip(gp, gc)}
        assert table == recount


def test_rescan_mode_equals_incremental(eng, taylorswift):
    from minbpe_b200 import engine as E
    data = taylorswift[:40000].encode("utf-8")
    eng.load_stream(data, None)
    p1, c1, d1 = eng.train(64)
    eng.set_option(E.OPT_RESCAN, 1)
    try:
        eng.load_stream(data, None)
        p2, c2, d2 = eng.train(64)
    finally:
        eng.set_option(E.OPT_RESCAN, 0)
    assert d1 == d2 == 64 and np.array_equal(p1, p2) and np.array_equal(c1, c2)


def test_table_growth_path(eng, golden_train, taylorswift):
    """Force a tiny pair table so that k_apply_delta hits its load limit repeatedly: the host grows
    the table and re-runs the apply; merges and counts must not change."""
    from minbpe_b200 import engine as E
    g = golden_train["taylorswift_basic_512"]
    eng.set_option(E.OPT_TABLE_LOG2, 13)
    eng.set_option(E.OPT_BATCH, 7)
    try:
        eng.load_stream(taylorswift.encode("utf-8"), None)
        pairs, counts, done = eng.train(256)
        tm = eng.timing()
    finally:
        eng.set_option(E.OPT_TABLE_LOG2, 0)
        eng.set_option(E.OPT_BATCH, 256)
    assert done == 256 and pairs.tolist() == g["merges"] and counts.tolist() == g["counts"]
    assert tm["table_slots"] > (1 << 13)  # it did grow


def test_encode_vs_oracle_random(eng):
    rng = np.random.default_rng(5)
    text = "".join(rng.choice(list("abc de\n"), size=20000))
    data, offs = oracle.split_to_stream(text, GPT4)
    eng.load_stream(data, offs)
    merges, _, done = eng.train(40)
    got = eng.encode(data, offs, merges)
    want = oracle.c_encode(data, offs, merges)
    assert np.array_equal(got, want)
    # a different text with the same merges; and a byte permutation (gpt4.py:76-77 style)
    other = "".join(rng.choice(list("abc de\n"), size=5000))
    d2, o2 = oracle.split_to_stream(other, GPT4)
    assert np.array_equal(eng.encode(d2, o2, merges), oracle.c_encode(d2, o2, merges))
    perm = rng.permutation(256).astype(np.uint8)
    assert np.array_equal(eng.encode(d2, o2, merges, perm), oracle.c_encode(d2, o2, merges, perm))
    # training state is untouched by encode
    assert eng.stream_len() > 0


def test_medium_synth_vs_oracle(eng):
    """4 MiB of the synthetic corpus, regex chunks, 48 merges: device vs oracle, all counts."""
    from minbpe_b200.synth import generate
    from minbpe_b200.tokenizer import split_text
    text = generate(1337, 4 << 20).tobytes().decode("utf-8")
    data, offs = split_text(GPT4, text)
    got, want = train_both(eng, data, offs, 48)
    assert got[2] == want[2] == 48
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def expand(ids, vocab_bytes, vocab_len):
    """ids -> bytes through the vocab, vectorised."""
    lens = vocab_len[ids]
    starts = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(lens, out=starts[1:])
    out = np.empty(int(starts[-1]), dtype=np.uint8)
    maxlen = int(vocab_len.max())
    for k in range(maxlen):
        sel = lens > k
        out[starts[:-1][sel] + k] = vocab_bytes[ids[sel], k]
    return out


def test_large_properties(eng):
    """64 MiB synthetic, Basic (one chunk) and chunked with synthetic cut points: properties that
    do not need the oracle at full size — (1) expanding the merged stream through the learned
    vocab gives back the corpus byte for byte, (2) stream length == n - sum(replacements) is
    consistent with the token accounting, (3) the incremental table equals a device recount,
    (4) the first merges equal the oracle's on a prefix-independent statistic (global top pair)."""
    from minbpe_b200.synth import generate
    n = 64 << 20
    raw = generate(1338, n)
    for offs in (None, np.unique(np.concatenate([[0], np.flatnonzero(raw == 32)])).astype(np.uint64)):
        eng.load_stream(raw, offs)
        K = 40
        pairs, counts, done = eng.train(K)
        assert done == K
        tm = eng.timing()
        cur = eng.read_stream()
        assert tm["tokens_in"] - tm["tokens_out"] == n - len(cur)
        # (1) expansion
        vocab = [bytes([i]) for i in range(256)]
        for a, b in pairs.tolist():
            vocab.append(vocab[a] + vocab[b])
        vlen = np.array([len(v) for v in vocab], dtype=np.int64)
        vb = np.zeros((len(vocab), int(vlen.max())), dtype=np.uint8)
        for i, v in enumerate(vocab):
            vb[i, : len(v)] = np.frombuffer(v, dtype=np.uint8)
        assert np.array_equal(expand(cur, vb, vlen), raw)
        # (3) table == recount
        table = eng.debug_table()
        gp, gc = eng.get_stats()
        assert table == {(int(a), int(b)): int(c) for (a, b), c in zip(gp, gc)}
        # (4) first merge = most frequent adjacent byte pair (numpy recount, ties by first position)
        w = raw.astype(np.int64)
        key = w[:-1] * 256 + w[1:]
        if offs is not None:
            valid = np.ones(n - 1, dtype=bool)
            valid[offs[1:].astype(np.int64) - 1] = False
            key = key[valid]
        hist = np.bincount(key, minlength=65536)
        top = hist.max()
        assert counts[0] == top
        tied = np.flatnonzero(hist == top)
        first = min(int(np.flatnonzero(key == t)[0]) for t in tied)
        assert pairs[0].tolist() == [int(key[first]) // 256, int(key[first]) % 256]


def test_encode_chunk_kernels_vs_oracle(eng):
    """Chunk-parallel encode: thread-per-chunk kernel (short chunks), CTA-per-chunk kernel (chunks of
    65..8192 tokens, incl. runs of one symbol = the a==a greedy rule), and the stream-round fallback
    for a chunk beyond 8192 tokens — all against the oracle, with and without a byte permutation."""
    from minbpe_b200.synth import generate
    from minbpe_b200.tokenizer import split_text
    text = generate(1337, 1 << 20).tobytes().decode("utf-8")
    data, offs = split_text(GPT4, text)
    eng.load_stream(data, offs)
    merges, _, done = eng.train(200)
    assert done == 200
    # craft an input with long chunks: long words, long symbol runs, one whitespace run
    extra = "x" * 70 + " " + "ab" * 500 + " " + "=" * 3000 + "\n" + "lyiltumdya" * 300 + " " + "z" * 8000 + " end"
    for t in (text[:200000] + extra, extra, "a", "ab", "hello world"):
        d2, o2 = split_text(GPT4, t)
        want = oracle.c_encode(d2, o2, merges)
        got = eng.encode(d2, o2, merges)
        assert np.array_equal(got, want), (len(t), len(got), len(want))
    perm = np.random.default_rng(3).permutation(256).astype(np.uint8)
    d2, o2 = split_text(GPT4, text[:50000])
    assert np.array_equal(eng.encode(d2, o2, merges, perm), oracle.c_encode(d2, o2, merges, perm))
    # a single chunk longer than ENC_LONG_MAX -> stream-round path
    long_one = ("lyiltumdya" * 1200).encode()
    assert np.array_equal(eng.encode(long_one, None, merges), oracle.c_encode(long_one, None, merges))
