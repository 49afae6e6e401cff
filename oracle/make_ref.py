#!/usr/bin/env python3
"""Vendor the UNMODIFIED pure-Python reference (karpathy/minbpe: minbpe/base.py, basic.py, regex.py) into
oracle/_ref/minbpe/ so that bench.py can time the reference's own code on the GPU box's host cores
(`cpu_baseline.python_reference`, `kind: "reference"`).  /root/reference does not exist on the GPU box; this
script runs HERE (called by __graft_entry__.build() when /root/reference is present), the output directory is
git-ignored (no reference source enters the history) but travels with the gpurun snapshot like a built .so.

Only the three files on the hot path are copied; the package __init__ written here imports just those (the
reference's own __init__ also imports gpt4.py -> tiktoken, which needs the network at construction time)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MINBPE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref", "minbpe")
FILES = ("base.py", "basic.py", "regex.py")


def make(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "minbpe")):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, "minbpe", f), os.path.join(DST, f))
    with open(os.path.join(DST, "__init__.py"), "w") as fh:
        fh.write("from .base import Tokenizer\nfrom .basic import BasicTokenizer\nfrom .regex import RegexTokenizer\n")
    if verbose:
        print("vendored", FILES, "->", DST)
    return True


def load():
    """Import the vendored reference package (None when it was never vendored)."""
    if not os.path.exists(os.path.join(DST, "regex.py")):
        return None
    import importlib
    root = os.path.join(HERE, "_ref")
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("minbpe")


if __name__ == "__main__":
    print(make(verbose=True))
