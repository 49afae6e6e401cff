"""TEST INFRASTRUCTURE: counterpart of minbpe_b200.dist.GpuStepEngine for the CPU SIMT emulator build of the library —
"device" buffers are torch CPU tensors (their data_ptr() is what the emulated kernels dereference), collectives run over
gloo.  Used by bench.py only under BPE_BENCH_EMU=1 (tests/test_emu.py)."""
import contextlib

import torch
import torch.distributed as dist


class HostStepEngine:
    stream = None

    def __init__(self, engine):
        self.e = engine

    def stream_ctx(self):
        return contextlib.nullcontext()

    def new_i64(self, n):
        return torch.zeros(n, dtype=torch.int64)

    def begin(self, dense):
        self.e.step_begin(dense.data_ptr())

    def table(self, dense, num_merges, first_idx, poll_every):
        self.e.step_table(dense.data_ptr(), num_merges, first_idx, poll_every)
        return self.e.step_delta_len()

    def select(self, cand, rank):
        self.e.step_select(cand.data_ptr(), rank)

    def merge(self, cand, delta):
        self.e.step_merge(cand.data_ptr(), delta.data_ptr())

    def apply(self, delta):
        self.e.step_apply(delta.data_ptr())

    def poll(self):
        return self.e.step_poll()

    def result(self, cap):
        return self.e.step_result(cap)

    def sync_ranks(self, group=None):
        dist.barrier(group=group)
