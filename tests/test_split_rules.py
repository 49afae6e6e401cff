"""CPU: the data-parallel restatement of the GPT-4 split pattern (oracle/split_rules.py) against the
installed `regex` module — the pattern minbpe uses at regex.py:19,41,114 — on the reference corpus,
the synthetic corpus, hand-written edge cases and random strings over an adversarial alphabet."""
import random

import regex

from oracle.split_rules import split

GPT4 = regex.compile(
    r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*|\s*[\r\n]|\s+(?!\S)|\s+""")

CASES = [
    "", "a", " ", "  ", "\n", " \n ", "hello world", "  hello   world  ", "it's we'll I'VE you'Re 'tis 'sfoo x'llama",
    " 's 'll", "a 'll b", "!!!word !word ! word", "\t!x", " !word", "12345 1 12 123 1234567", "a1b22c333d4444",
    "x\n\n\ny", "!\n\n  x", "a \n b", "a  \n  \n  b", "\r\n\r\n a\r\nb \r\n", "a b c", "tab\t\tend\t", "''''a", "'s", "''s", "a''s",
    "日本語 テキスト ١٢٣٤ ½⅓", "é'S ʼs", "don't.stop!!!\n\nnow", " 　x", "x 　", "\n x", "\n  x", "!\nx", "! \nx", "!' s",
    "a'ſ b'K", "...'ll", "-'ll", "1's 22'll",
]


def check(text):
    want = GPT4.findall(text)
    got = split(text)
    assert got == want, (text, got, want)


def test_cases():
    for t in CASES:
        check(t)


def test_corpora(taylorswift):
    check(taylorswift)
    from minbpe_b200.synth import generate
    check(generate(1337, 1 << 20).tobytes().decode("utf-8"))


def test_random_adversarial():
    rnd = random.Random(12345)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '"]
    for _ in range(6000):
        n = rnd.randint(0, 24)
        check("".join(rnd.choice(alphabet) for _ in range(n)))
    # random code points from several planes
    pools = [range(0x20, 0x7f), range(0xa0, 0x250), range(0x370, 0x400), range(0x2000, 0x2070), range(0x3000, 0x3100),
             range(0x1f600, 0x1f650), [0x9, 0xa, 0xd, 0x20, 0x85, 0x1680, 0x2028, 0x2029, 0x202f, 0x205f]]
    for _ in range(2000):
        n = rnd.randint(0, 30)
        check("".join(chr(rnd.choice(rnd.choice(pools))) for _ in range(n)))


def test_local_formulation():
    """oracle/split_rules_local.py: the same rules from six segmented scans + neighbours at distance <= 2
    (the form the next device splitter evaluates tile by tile) against `regex`."""
    from oracle.split_rules_local import split as split_local

    def check_local(text):
        assert split_local(text) == GPT4.findall(text), text

    for t in CASES:
        check_local(t)
    rnd = random.Random(777)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '"]
    for _ in range(4000):
        check_local("".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 24))))
    pools = [range(0x20, 0x7f), range(0xa0, 0x250), range(0x2000, 0x2070), range(0x3000, 0x3100),
             [0x9, 0xa, 0xd, 0x20, 0x85, 0x1680, 0x2028, 0x2029, 0x202f, 0x205f]]
    for _ in range(1000):
        check_local("".join(chr(rnd.choice(rnd.choice(pools))) for _ in range(rnd.randint(0, 30))))


def test_local_formulation_corpus(taylorswift):
    from oracle.split_rules_local import split as split_local
    text = taylorswift[:60000]
    assert split_local(text) == GPT4.findall(text)


def test_product_rule_code_on_cpu(taylorswift):
    """minbpe_b200/csrc/split_logic.h — the scan operators and the chunk-start rule the CUDA splitter is built
    from — compiled for the CPU (oracle/split_harness.cpp) with the scans evaluated tile by tile like the
    kernels do, against `regex` (byte offsets of every chunk)."""
    import numpy as np

    import oracle
    from minbpe_b200.unicode_tables import tables
    cls, contr = tables()

    def check_bytes(text, tiles):
        data, offs = oracle.split_to_stream(text, GPT4)
        for tile in tiles:
            got = oracle.split_logic_offsets(data.tobytes(), cls, contr, tile)
            assert np.array_equal(got, offs), (text[:80], tile)

    for t in CASES:
        if t:
            check_bytes(t, (0, 1, 3, 7, 64))
    rnd = random.Random(4242)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '", "K", "\U0001f600"]
    for _ in range(4000):
        check_bytes("".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 24))), (0, rnd.randint(1, 9)))
    check_bytes(taylorswift, (0, 5, 2048, 4096))
    from minbpe_b200.synth import generate
    check_bytes(generate(1337, 2 << 20).tobytes().decode("utf-8"), (4096,))


def test_product_rule_code_with_special_token_boundaries():
    """regex.py:152-163: the text is cut at the special tokens first (re.split, leftmost, first alternative wins, no
    overlaps) and every part is split by the GPT-4 pattern ON ITS OWN.  The device does it in one pass: the bytes of every
    occurrence get the boundary class SC_B and the WITH_B instantiation of split_logic.h treats them as end / start of
    text.  Pinned here against `regex` on random adversarial texts with specials sprinkled in — including specials that
    begin or end with a space, a letter, an apostrophe or a digit, adjacent specials, and specials at either end."""
    import numpy as np

    import oracle
    from minbpe_b200.unicode_tables import tables
    cls, contr = tables()
    rnd = random.Random(777)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '", "K", "\U0001f600"]
    special_sets = [["<|endoftext|>"], ["<|a|>", "<|ab|>", "|>"], [" <s> ", "'s", "12"], ["x", "日"], ["\n<eot>\n", "  "], ["a b", "b a"]]

    def expected(text, specials):
        pat = "(" + "|".join(regex.escape(k) for k in specials) + ")"
        offs, hits, pos = [], [], 0
        for part in regex.split(pat, text):
            nb = len(part.encode("utf-8"))
            if part in specials:
                offs.append(pos)
                hits.append((pos, nb))
            else:
                p = pos
                for ch in GPT4.findall(part):
                    offs.append(p)
                    p += len(ch.encode("utf-8"))
            pos += nb
        return np.asarray(offs, dtype=np.uint64), hits

    n_hits = 0
    for it in range(5000):
        specials = rnd.choice(special_sets)
        parts = []
        for _ in range(rnd.randint(1, 5)):
            parts.append("".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 12))))
            if rnd.random() < 0.8:
                parts.append(rnd.choice(specials))
        text = "".join(parts)
        if not text:
            continue
        want, hits = expected(text, specials)
        n_hits += len(hits)
        data = text.encode("utf-8")
        for tile in (0, rnd.randint(1, 9)):
            got = oracle.split_logic_offsets_special(data, cls, contr, hits, tile)
            assert np.array_equal(got, want), (text, specials, tile, got.tolist(), want.tolist())
    assert n_hits > 5000


GPT2 = regex.compile(r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")


def test_product_rule_code_gpt2_pattern(taylorswift):
    """The GPT-2 split pattern (regex.py:18) from the same scans (spl_chunk_start_gpt2), plain and with special-token
    boundaries, against `regex`."""
    import numpy as np

    import oracle
    from minbpe_b200.unicode_tables import tables
    cls, contr = tables()

    def check_bytes(text, tiles):
        data, offs = oracle.split_to_stream(text, GPT2)
        for tile in tiles:
            got = oracle.split_logic_offsets_gpt2(data.tobytes(), cls, contr, (), tile)
            assert np.array_equal(got, offs), (text[:80], tile, got.tolist()[:20], offs.tolist()[:20])

    for t in CASES:
        if t:
            check_bytes(t, (0, 1, 3, 7, 64))
    rnd = random.Random(2024)
    alphabet = list("ab'sSdDmMtTlLvVeErR 12\t\n\r!.,' 　é日ſ½") + ["  ", "\n\n", "'ll", "'ve", " '", "K", "\U0001f600", "'re", "'LL"]
    for _ in range(6000):
        check_bytes("".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 24))), (0, rnd.randint(1, 9)))
    check_bytes(taylorswift, (0, 5, 4096))
    from minbpe_b200.synth import generate
    check_bytes(generate(1337, 1 << 20).tobytes().decode("utf-8"), (4096,))
    # with specials as boundaries
    special_sets = [["<|endoftext|>"], ["<|a|>", "<|ab|>", "|>"], [" <s> ", "'s", "12"], ["x", "日"], ["\n<eot>\n", "  "], ["a b", "b a"]]
    for it in range(3000):
        specials = rnd.choice(special_sets)
        parts = []
        for _ in range(rnd.randint(1, 5)):
            parts.append("".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 12))))
            if rnd.random() < 0.8:
                parts.append(rnd.choice(specials))
        text = "".join(parts)
        if not text:
            continue
        pat = "(" + "|".join(regex.escape(k) for k in specials) + ")"
        offs, hits, pos = [], [], 0
        for part in regex.split(pat, text):
            nb = len(part.encode("utf-8"))
            if part in specials:
                offs.append(pos)
                hits.append((pos, nb))
            else:
                p = pos
                for ch in GPT2.findall(part):
                    offs.append(p)
                    p += len(ch.encode("utf-8"))
            pos += nb
        got = oracle.split_logic_offsets_gpt2(text.encode("utf-8"), cls, contr, hits, rnd.choice((0, 3, 8)))
        assert np.array_equal(got, np.asarray(offs, dtype=np.uint64)), (text, specials)
