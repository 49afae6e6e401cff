"""GPU: resume training from a saved model (base.py:97-165 save/load either side of the hot path).
bpe_replay applies the loaded merges to the text with the training kernels (pair table maintained on the way), then
bpe_train continues the same run."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["basic", "regex"])
def test_resume_from_model_file(kind, golden_train, taylorswift, tmp_path):
    """train(512) == train(384); save; load into a NEW tokenizer; train(512, resume=True).
    The loaded merges are replayed on the text with the training kernels (bpe_replay), then bpe_train continues:
    merges, vocab and the saved .model must equal the reference's 512-vocab golden run."""
    import hashlib
    from minbpe_b200 import BasicTokenizer, RegexTokenizer
    g = golden_train[f"taylorswift_{kind}_512"]
    cls = BasicTokenizer if kind == "basic" else RegexTokenizer
    a = cls()
    a.train(taylorswift, 384)
    assert [list(p) for p in a.merges] == g["merges"][:128]
    a.save(str(tmp_path / "half"))
    b = cls()
    b.load(str(tmp_path / "half.model"))
    b.train(taylorswift, 512, resume=True)
    assert [list(p) for p in b.merges] == g["merges"]
    assert list(b.merges.values()) == list(range(256, 512))
    b.save(str(tmp_path / "full"))
    assert hashlib.sha256(open(str(tmp_path / "full.model"), "rb").read()).hexdigest() == g["model_sha256"]
    assert hashlib.sha256(open(str(tmp_path / "full.vocab"), "rb").read()).hexdigest() == g["vocab_sha256"]
    # resume with nothing left to do, and with nothing to resume from
    b.train(taylorswift, 512, resume=True)
    assert [list(p) for p in b.merges] == g["merges"]
    c = cls()
    c.train(taylorswift, 300, resume=True)
    assert [list(p) for p in c.merges] == g["merges"][:44]
    with pytest.raises(ValueError):
        b.train(taylorswift, 300, resume=True)
