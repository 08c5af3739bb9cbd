"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

CPU restatement of the two `openai-whisper` (`whisper/timing.py`, third party,
un-vendored, un-pinned) routines the reference's hot path calls:

  * median_filter -- reference call sites timing.py:65 and timing.py:158
  * dtw           -- reference call sites timing.py:103 and timing.py:165

Only the CPU branches are restated (the reference always reaches `dtw` with a
`.cpu()` tensor, timing.py:102-103, and the CPU median is the `unfold().sort()`
form).  The upstream Triton branches are deliberately NOT the parity target:
the upstream Triton DTW breaks ties differently from the CPU recurrence.
"""
from __future__ import annotations

import numba
import numpy as np
import torch
import torch.nn.functional as F


def median_filter(x: torch.Tensor, filter_width: int):
    """Sliding median of odd width along the last dim, reflect padding.

    Identity when the last dim is not longer than the one-sided pad (the pad
    would be illegal), exactly as the published routine behaves.
    """
    half = filter_width // 2
    if x.shape[-1] <= half:
        return x
    ndim = x.ndim
    if ndim <= 2:
        x = x[None, None, :]
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    x = F.pad(x, (half, half, 0, 0), mode="reflect")
    out = x.unfold(-1, filter_width, 1).sort()[0][..., half]
    if ndim <= 2:
        out = out[0, 0]
    return out


@numba.jit(nopython=True)
def backtrace(trace: np.ndarray):
    i = trace.shape[0] - 1
    j = trace.shape[1] - 1
    # border overrides: row 0 always steps in time, column 0 always steps in text
    trace[0, :] = 2
    trace[:, 0] = 1
    out = []
    while i > 0 or j > 0:
        out.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i -= 1
            j -= 1
        elif t == 1:
            i -= 1
        elif t == 2:
            j -= 1
        else:
            raise ValueError("Unexpected trace[i, j]")
    arr = np.array(out)
    return arr[::-1, :].T


@numba.jit(nopython=True, parallel=True)
def dtw_cpu(x: np.ndarray):
    """min-of-three DTW; fp32 cost table, ties resolve to the time step (code 2)."""
    N, M = x.shape
    cost = np.ones((N + 1, M + 1), dtype=np.float32) * np.inf
    trace = -np.ones((N + 1, M + 1), dtype=np.float32)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0 = cost[i - 1, j - 1]
            c1 = cost[i - 1, j]
            c2 = cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    return backtrace(trace)


def dtw(x: torch.Tensor) -> np.ndarray:
    # CPU path only: double-precision view of the input, fp32 cost table.
    return dtw_cpu(x.double().cpu().numpy())
