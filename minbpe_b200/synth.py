"""
Synthetic UTF-8 corpus generator — bench/test tooling, not the BPE path.
Thin ctypes wrapper over csrc/synth.c (host C, multi-threaded, deterministic per seed).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "libbpesynth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = ctypes.CDLL(_SO)
        _lib.bpe_synth_generate.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        _lib.bpe_synth_generate.restype = ctypes.c_int
        _lib.bpe_synth_generate_at.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        _lib.bpe_synth_generate_at.restype = ctypes.c_int
    return _lib


def generate(seed, nbytes, threads=None, out=None, first_block=0):
    """Return a uint8 numpy array of exactly ``nbytes`` bytes of valid UTF-8 for ``seed``
    (cfg3: seed 1337 / 2**30 B, cfg4: 1338 / 2**34 B, cfg5: 1339 / 4e9 B).  ``first_block`` > 0
    returns the bytes a full generation would hold from offset ``first_block`` MiB on (a contiguous
    shard of the same corpus; every block is padded to exactly 1 MiB and ends on a character)."""
    if out is None:
        out = np.empty(int(nbytes), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size == nbytes and out.flags["C_CONTIGUOUS"]
    threads = threads or min(32, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    rc = _load().bpe_synth_generate_at(int(seed), int(first_block), out.ctypes.data, int(nbytes), int(threads))
    if rc != 0:
        raise RuntimeError(f"bpe_synth_generate failed: {rc}")
    return out
