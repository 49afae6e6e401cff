"""
GPT4Tokenizer (reference: minbpe/gpt4.py:57-130) on the device encode path.

The reference class is a RegexTokenizer whose merges are RECOVERED from a tiktoken rank table (`cl100k_base`: token bytes
-> rank), whose single-byte tokens are permuted (rank of byte b != b: the "byte shuffle", gpt4.py:76-77) and which
refuses train / save / load.  Everything that makes it different from RegexTokenizer is data: a merges table and a
256-entry byte permutation — exactly the two extra arguments the C ABI's encode entry points take (`byte_perm`), so
encode / encode_ordinary / the special-token front end run on the GPU unchanged.

`tiktoken.get_encoding("cl100k_base")` downloads the table on first use; without network (this container, the GPU
boxes) construct the class from any rank table of the same shape: `GPT4Tokenizer(mergeable_ranks={bytes: rank, ...})`
(keyword not in the reference), e.g. one read from a local `.tiktoken` file with `load_tiktoken_file`.
"""
import base64

import numpy as np

from .tokenizer import GPT4_SPLIT_PATTERN, RegexTokenizer, render_token

# gpt4.py:49-55
GPT4_SPECIAL_TOKENS = {
    "<|endoftext|>": 100257,
    "<|fim_prefix|>": 100258,
    "<|fim_middle|>": 100259,
    "<|fim_suffix|>": 100260,
    "<|endofprompt|>": 100276,
}


def load_tiktoken_file(path):
    """A `.tiktoken` rank file (lines of `<base64 token> <rank>`) -> {bytes: rank} in file order."""
    ranks = {}
    with open(path, "rb") as f:
        for line in f:
            if line.strip():
                tok, rank = line.split()
                ranks[base64.b64decode(tok)] = int(rank)
    return ranks


def _parents(ranks, token, limit):
    """The two byte strings whose merge produced `token` (rank `limit`): BPE on the token's own bytes using only merges
    of lower rank leaves exactly them (gpt4.py:11-27 `bpe` with max_rank)."""
    pieces = [token[i:i + 1] for i in range(len(token))]
    while len(pieces) > 1:
        best_at, best = -1, limit
        for i in range(len(pieces) - 1):
            r = ranks.get(pieces[i] + pieces[i + 1])
            if r is not None and r < best:
                best_at, best = i, r
        if best_at < 0:
            break
        pieces[best_at:best_at + 2] = [pieces[best_at] + pieces[best_at + 1]]
    return pieces


def recover_merges(mergeable_ranks):
    """gpt4.py:30-46: {(rank of left parent, rank of right parent): rank} for every multi-byte token, in table order."""
    merges = {}
    for token, rank in mergeable_ranks.items():
        if len(token) < 2:
            continue
        parts = _parents(mergeable_ranks, token, rank)
        if len(parts) != 2:
            raise ValueError(f"rank table is not a BPE merge forest: token {token!r} does not split into two parents")
        merges[(mergeable_ranks[parts[0]], mergeable_ranks[parts[1]])] = rank
    return merges


def _cl100k_ranks():
    try:
        import tiktoken
        return tiktoken.get_encoding("cl100k_base")._mergeable_ranks
    except Exception as ex:  # noqa: BLE001  (no network, no cached vocabulary, tiktoken missing)
        raise RuntimeError("GPT4Tokenizer needs tiktoken's cl100k_base rank table, which could not be loaded here "
                           f"({ex!r}); pass mergeable_ranks=... (e.g. load_tiktoken_file(path))") from ex


class GPT4Tokenizer(RegexTokenizer):
    """gpt4.py:57-130.  Pretrained: train / save / load raise, as in the reference."""

    def __init__(self, *, mergeable_ranks=None, special_tokens=None, device=None):
        super().__init__(pattern=GPT4_SPLIT_PATTERN, device=device)
        ranks = _cl100k_ranks() if mergeable_ranks is None else mergeable_ranks
        singles = [ranks.get(bytes((b,))) for b in range(256)]
        if any(r is None for r in singles) or sorted(singles) != list(range(256)):
            raise ValueError("rank table must give the 256 single bytes the ranks 0..255 (in any order)")
        self.merges = recover_merges(ranks)
        # ids of single bytes are their ranks: vocab lives in the permuted byte space (gpt4.py:67-71) ...
        vocab = {i: bytes((i,)) for i in range(256)}
        for (left, right), idx in self.merges.items():
            vocab[idx] = vocab[left] + vocab[right]
        self.vocab = vocab
        # ... and text bytes are permuted on the way in, un-permuted on the way out (gpt4.py:76-77)
        self.byte_shuffle = {b: singles[b] for b in range(256)}
        self.inverse_byte_shuffle = {r: b for b, r in self.byte_shuffle.items()}
        self._byte_perm = np.asarray(singles, dtype=np.uint8)
        self._unshuffle = bytes(self.inverse_byte_shuffle[i] for i in range(256))
        self.register_special_tokens(dict(GPT4_SPECIAL_TOKENS) if special_tokens is None else special_tokens)

    def decode(self, ids):
        """gpt4.py:88-93: vocabulary bytes joined, bytes un-permuted, utf-8 with replacement.  Unknown ids raise KeyError
        (the reference indexes self.vocab directly), special tokens included — as there."""
        if self._decode_on_device(len(ids)):
            data, bad = self._device_decode(ids, self.vocab)
            if data is None:
                raise KeyError(ids[bad])
        else:
            data = b"".join(self.vocab[idx] for idx in ids)
        return data.translate(self._unshuffle).decode("utf-8", errors="replace")

    def train(self, text, vocab_size, verbose=False):
        raise NotImplementedError

    def save(self, file_prefix):
        raise NotImplementedError("GPT4Tokenizer cannot be saved.")

    def load(self, model_file):
        raise NotImplementedError("GPT4Tokenizer cannot be loaded.")

    def save_vocab(self, vocab_file):
        """gpt4.py:110-130: the .vocab rendering of the base class, with the single bytes shown un-permuted."""
        shown = {i: bytes((self.inverse_byte_shuffle[i],)) for i in range(256)}
        for (left, right), idx in self.merges.items():
            shown[idx] = shown[left] + shown[right]
        parents = {idx: pair for pair, idx in self.merges.items()}
        with open(vocab_file, "w", encoding="utf-8") as f:
            for idx, token in shown.items():
                if idx in parents:
                    left, right = parents[idx]
                    f.write(f"[{render_token(shown[left])}][{render_token(shown[right])}] -> [{render_token(token)}] {idx}\n")
                else:
                    f.write(f"[{render_token(token)}] {idx}\n")
