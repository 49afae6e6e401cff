"""CPU: the C-ABI library loads and exports every symbol include/b200bpe.h declares, the Python
loader binds them all, and the product fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b200bpe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpe_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from minbpe_b200 import engine
    names = declared_functions()
    assert {"bpe_create", "bpe_destroy", "bpe_load_stream", "bpe_get_stats", "bpe_merge", "bpe_train",
            "bpe_encode", "bpe_last_error"} <= set(names)
    lib = ctypes.CDLL(engine.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200bpe.h but not exported"
    L = engine.load_library()  # binds argtypes for every entry point; raises on mismatch
    assert L.bpe_abi_version() == engine.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from minbpe_b200 import BasicTokenizer, engine
    with pytest.raises(engine.EngineError):
        engine.Engine()
    with pytest.raises(engine.EngineError):
        BasicTokenizer().train("aaabdaaabac", 259)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under minbpe_b200/ may reference it."""
    pkg = os.path.join(ROOT, "minbpe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
