"""Length bucketing of utterances into batches.

The reference aligns strictly one utterance at a time (infer_ali.py:48,57) and skips what does not fit
Whisper's context: more than 1500 frames (30 s) or more than 448 tokens (infer_ali.py:78-81).  Here many
utterances go through one forward and one launch per stage, so what shares a batch matters:

  * the decoder runs on token rows right-padded to the longest sequence of the batch -> sort by token count;
  * the capture kernel launches one grid per frame-cluster size (1-6 or 8 CTAs of 224 frames) -> utterances
    of similar duration should share a batch (text length and duration are strongly correlated in speech);
  * the maps of a batch are `4 * L * H * sum(T * F)` bytes and live in HBM at once -> cap them.

`plan_batches` returns lists of indices; every index of a kept utterance appears exactly once.
"""
from __future__ import annotations

from typing import Sequence

MAX_FRAMES = 1500  # reference infer_ali.py:25
MAX_LENGTH = 448   # reference infer_ali.py:26


def fits_context(n_tokens: int, max_frames: int) -> bool:
    """The reference's skip rule (infer_ali.py:78-81): False for utterances it prints and skips."""
    return 1 <= max_frames <= MAX_FRAMES and n_tokens <= MAX_LENGTH


def utterance_cost(n_tokens: int, max_frames: int, n_layers: int = 24, width: int = 1024, n_enc_layers: int | None = None,
                   n_ctx: int = 1500) -> float:
    """Rough FLOP-equivalent cost of one utterance, used only to balance shards and order batches: the fixed
    30 s encoder, the decoder linears (proportional to T, with the cross-attention K/V projections of the 1500
    encoder frames as a fixed part) and the map traffic (proportional to T * F, weighted as bytes * 200 FLOP/B,
    the ridge of the machine)."""
    n_enc_layers = n_layers if n_enc_layers is None else n_enc_layers
    enc = n_enc_layers * (24.0 * n_ctx * width * width + 4.0 * n_ctx * n_ctx * width)
    dec_fixed = n_layers * 4.0 * n_ctx * width * width
    dec = n_layers * n_tokens * (28.0 * width * width + 4.0 * n_ctx * width)
    maps = 8.0 * n_layers * (width // 64) * n_tokens * max_frames * 200.0
    return enc + dec_fixed + dec + maps


def plan_batches(n_tokens: Sequence[int], max_frames: Sequence[int], batch_size: int, *, n_maps: int = 384,
                 map_budget_bytes: float = 24e9):
    """Indices grouped into batches of at most `batch_size` utterances of similar length.
    n_maps = L * H (maps per utterance); a batch is closed early when its maps would exceed the budget.
    Returns (batches, skipped): skipped are the indices the reference would skip."""
    keep = [i for i in range(len(n_tokens)) if fits_context(int(n_tokens[i]), int(max_frames[i]))]
    skipped = [i for i in range(len(n_tokens)) if i not in set(keep)]
    keep.sort(key=lambda i: (int(n_tokens[i]), int(max_frames[i]), i))
    batches, cur, cur_bytes = [], [], 0.0
    for i in keep:
        b = 4.0 * n_maps * int(n_tokens[i]) * int(max_frames[i])
        if cur and (len(cur) >= batch_size or cur_bytes + b > map_budget_bytes):
            batches.append(cur)
            cur, cur_bytes = [], 0.0
        cur.append(i)
        cur_bytes += b
    if cur:
        batches.append(cur)
    return batches, skipped


def padding_waste(n_tokens: Sequence[int], batches) -> float:
    """Fraction of decoder token rows that are padding under a batch plan (0 = none)."""
    real = padded = 0
    for b in batches:
        t = [int(n_tokens[i]) for i in b]
        real += sum(t)
        padded += max(t) * len(t)
    return 1.0 - real / padded if padded else 0.0
