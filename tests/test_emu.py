"""CPU: the kernel SOURCES of minbpe_b200/csrc on the SIMT emulator (tests/emu/, test infrastructure).

tests/emu/build_emu.py compiles b200bpe.cu — kernels and host side — with g++ against tests/emu/cuda_emu.h (one fiber
per CUDA thread, rendezvous barriers and warp collectives, guarded allocations) into libb200bpe_emu.so, which exports
the C ABI of include/b200bpe.h.  The GPU parity tests are then run against THAT library (BPE_LIB_PATH) in
subprocesses: the same test code, the same oracle, the kernels' logic executed on the CPU.

It is a logic check for a container without a GPU, NOT a product path (nothing in minbpe_b200/ loads the emulator
build) and not a substitute for `-m gpu` on a B200: performance, the PTX paths (TMA / mbarrier are emulated as
immediate copies) and the cross-GPU memory model are out of its reach.

The four jobs start together (they are independent processes) and each test waits for its own job."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _env(lib):
    env = dict(os.environ)
    env.update(BPE_LIB_PATH=lib, BPE_TEST_SMALL="1", EMU_SMS="2", PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTEST_CURRENT_TEST", None)
    return env


def _pytest(lib, files, k=None):
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-o", "timeout=900"] + files
    if k:
        cmd += ["-k", k]
    return subprocess.Popen(cmd, cwd=ROOT, env=_env(lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


@pytest.fixture(scope="module")
def jobs():
    sys.path.insert(0, EMU)
    import build_emu
    lib = build_emu.build()
    t = lambda f: os.path.join("tests", f)  # noqa: E731
    procs = {
        # the round-2 kernels that have not run on a GPU yet: memoised chunk encode (k_encode2.cuh), bpe_replay / resume,
        # bpe_decode
        "new_kernels": _pytest(lib, [t("test_gpu_zy_encode2.py"), t("test_gpu_zz_resume.py"), t("test_gpu_decode.py")]),
        # kernels already validated on B200s, as a check of the emulator itself (golden vectors of the reference)
        "validated_kernels": _pytest(lib, [t("test_gpu_parity.py")],
                                     "wikipedia or taylorswift or small_cases or primitives or long_runs or table_growth or rescan"),
        "splitter": _pytest(lib, [t("test_gpu_split.py")], "not piecewise"),
        # the sharded loop on 2..4 emulated GPUs (threads): NCCL-style collectives and the NVLink peer-memory kernels
        "sharded": subprocess.Popen([sys.executable, os.path.join(EMU, "emu_sharded.py"), "2:collective", "2:p2p", "3:p2p", "4:p2p"],
                                    cwd=ROOT, env=_env(lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True),
    }
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


def _finish(jobs, name, timeout=1500):
    p = jobs[name]
    try:
        out, _ = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        p.kill()
        out, _ = p.communicate()
        pytest.fail(f"emulator job {name} did not finish in {timeout} s\n{out[-3000:]}")
    assert p.returncode == 0, f"emulator job {name} failed\n{out[-6000:]}"
    return out


def test_emu_new_kernels_encode2_resume_decode(jobs):
    out = _finish(jobs, "new_kernels")
    assert " passed" in out and "failed" not in out


def test_emu_validated_kernels_against_reference_goldens(jobs):
    out = _finish(jobs, "validated_kernels")
    assert " passed" in out and "failed" not in out


def test_emu_gpt4_splitter(jobs):
    out = _finish(jobs, "splitter")
    assert " passed" in out and "failed" not in out


def test_emu_sharded_loop_collective_and_p2p(jobs):
    out = _finish(jobs, "sharded")
    assert "emu sharded ok" in out
    assert out.count("bit-exact on every rank") == 12


def test_emulator_is_not_a_product_path():
    """Nothing under minbpe_b200/ may reach the emulator build or its header."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "minbpe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".c")):
                text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                assert "libb200bpe_emu" not in text and '#include "cuda_emu.h"' not in text and "emu/_build" not in text, f
