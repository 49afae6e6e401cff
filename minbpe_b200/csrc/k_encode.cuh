// k_encode.cuh — per-chunk encode kernels (regex.py:92-109).  Round 1: encode runs through the
// stream kernels (k_select_rank + k_merge + k_apply_delta, one round per applicable merge rank);
// the warp-per-chunk kernel for corpora with many small chunks lands here next.
#pragma once
#include "common.cuh"
