"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Arithmetic is torch-CPU fp32 so that it is the same arithmetic
the reference performs with DEVICE='cpu' (infer_ali.py:22); the DTW is the C
restatement in dtw_oracle.c.  Checked bit-for-bit against fixtures produced by the
reference's own code (oracle/gen_golden.py -> tests/golden/).
"""
from __future__ import annotations

import string

import numpy as np
import torch
import torch.nn.functional as F

from . import dtw as _dtw
from . import use_shim

use_shim()
from whisper.model import disable_sdpa  # noqa: E402  (restated dependency)

TOKENS_PER_SECOND = 50  # whisper.audio: 16000 / (160 * 2); used at timing.py:111


# --------------------------------------------------------------------------
# get_attentions, timing.py:45-67
# --------------------------------------------------------------------------
def capture_logits(model, mel, tokens):
    """timing.py:48-63.  Pre-softmax cross-attention logits of every decoder layer,
    stacked to (L, H, T, n_audio_ctx) fp32, plus the decoder output logits (T, V).
    The hook reads the LAST element of cross_attn's return tuple (timing.py:52)."""
    grabbed = {}
    handles = []
    for idx, blk in enumerate(model.decoder.blocks):
        handles.append(
            blk.cross_attn.register_forward_hook(lambda _m, _i, outs, idx=idx: grabbed.__setitem__(idx, outs[-1]))
        )
    try:
        with torch.no_grad(), disable_sdpa():
            out_logits = model(mel.unsqueeze(0), tokens.unsqueeze(0))[0]
    finally:
        for h in handles:
            h.remove()
    qk = torch.cat([grabbed[i] for i in range(model.dims.n_text_layer)])
    return qk, out_logits


def median_along_frames(x, width: int):
    """whisper.timing.median_filter as reached from timing.py:65: reflect-pad by
    width//2 INSIDE the trimmed window, take the middle order statistic; identity
    if the row is not longer than the pad."""
    half = width // 2
    if x.shape[-1] <= half:
        return x
    assert width > 0 and width % 2 == 1
    lead = x.shape[:-1]
    rows = x.reshape(1, -1, x.shape[-1])
    padded = F.pad(rows, (half, half), mode="reflect")[0]
    windows = padded.unfold(-1, width, 1)
    return windows.sort(dim=-1)[0][..., half].reshape(*lead, -1)


def filtered_softmax(qk, max_frames: int, medfilt_width: int, qk_scale: float):
    """timing.py:64-66: trim to max_frames, median-filter the LOGITS, scale, softmax
    over the trimmed frames."""
    w = qk[..., :max_frames]
    w = median_along_frames(w, medfilt_width)
    return (w * qk_scale).softmax(dim=-1)


def get_attentions(mel, tokens, model, tokenizer, max_frames, medfilt_width=7, qk_scale=1.0):
    """timing.py:45-67 (tokenizer is accepted and unused there too)."""
    qk, out_logits = capture_logits(model, mel, tokens)
    return filtered_softmax(qk, int(max_frames), medfilt_width, qk_scale), out_logits


# --------------------------------------------------------------------------
# filter_attention, timing.py:13-43 ; coverage_penalty, metrics.py:99-111
# --------------------------------------------------------------------------
def coverage_penalty(attn, threshold: float = 0.5):
    """metrics.py:99-111: sum_f max(sum_t a[t,f], thr) - F*thr."""
    cov = attn.sum(dim=0)
    return torch.maximum(cov, torch.full_like(cov, threshold)).sum(-1) - cov.size(-1) * threshold


def head_scores(attns, w_colnorm=1.0, w_rownorm=1.0, w_coverage=0.0):
    """timing.py:17-34: per-head score, (L, H) fp32.
    colnorm term: L2 over tokens, summed over frames (:21); rownorm term: L2 over
    frames, summed over tokens (:24); minus the weighted coverage penalty (:30-32)."""
    score = torch.zeros(attns.shape[0], attns.shape[1])
    if w_colnorm > 0:
        score += w_colnorm * attns.norm(dim=-2).sum(-1)
    if w_rownorm > 0:
        score += w_rownorm * attns.norm(dim=-1).sum(-1)
    if w_coverage > 0:
        for l in range(attns.shape[0]):
            for h in range(attns.shape[1]):
                score[l, h] -= w_coverage * coverage_penalty(attns[l, h])
    return score


def filter_attention(attns, topk=20, w_colnorm=1, w_rownorm=1, w_coverage=0):
    """timing.py:13-43: ascending sort of (score, (l, h), name) tuples, keep the last
    `topk`; the selected maps come back in that same ascending order."""
    score = head_scores(attns, w_colnorm, w_rownorm, w_coverage)
    table = [
        (score[l, h].item(), (l, h), f"sample_layer{l}_head{h}")
        for l in range(attns.shape[0])
        for h in range(attns.shape[1])
    ]
    kept = sorted(table)[-topk:]
    return [attns[l, h].unsqueeze(0) for _, (l, h), _ in kept], kept


# --------------------------------------------------------------------------
# retokenize.py:5-39
# --------------------------------------------------------------------------
def encode(text, tokenizer, aligned_unit_type="subword"):
    """retokenize.py:5-17: char units = one encode() per character, single space
    token between words."""
    assert aligned_unit_type in ["char", "subword"]
    if aligned_unit_type == "subword":
        return tokenizer.encode(text)
    space = tokenizer.encode(" ")
    pieces = text.split()
    ids = []
    for n, wd in enumerate(pieces):
        for ch in wd:
            ids += tokenizer.encode(ch)
        if n + 1 < len(pieces):
            ids += space
    return ids


def split_tokens_on_spaces(tokens, tokenizer, aligned_unit_type="subword"):
    """retokenize.py:19-39.  Char units: a new word opens on a special token, on a
    piece that IS a single space, or at the very start (note: not `startswith`)."""
    assert aligned_unit_type in ["char", "subword"]
    if aligned_unit_type == "subword":
        return tokenizer.split_to_word_tokens(tokens)
    pieces, piece_tokens = tokenizer.split_tokens_on_unicode(tokens)
    words, word_tokens = [], []
    for piece, toks in zip(pieces, piece_tokens):
        opens = toks[0] >= tokenizer.eot or piece == " " or not words
        if opens:
            words.append(piece)
            word_tokens.append(toks)
        else:
            words[-1] += piece
            word_tokens[-1].extend(toks)
    return words, word_tokens


# --------------------------------------------------------------------------
# force_align, timing.py:69-114
# --------------------------------------------------------------------------
def aggregate(ws, aggregation="mean", topk=-1, w_colnorm=1.0, w_rownorm=1.0, w_coverage=0.0):
    """timing.py:83-100: the (T, F) matrix fed to DTW, and the score table."""
    scores = None
    if aggregation == "mean":
        # :86-89 -- L2 over tokens, upper half of the layers, mean over (layer, head)
        normed = ws / ws.norm(dim=-2, keepdim=True)
        matrix = normed[ws.size(0) // 2 :].mean(axis=(0, 1))
    elif aggregation == "topk":
        assert topk > 0  # :92
        picked, scores = filter_attention(ws, topk, w_colnorm, w_rownorm, w_coverage)
        stack = torch.cat(picked, 0)
        matrix = torch.mean(stack / stack.norm(dim=-2, keepdim=True), 0)  # :95-97
    elif aggregation == "grad_norm":
        matrix = ws  # :99-100
    else:
        raise UnboundLocalError("matrix")  # what the reference raises at :102
    return matrix, scores


def boundaries_from_path(text_indices, time_indices, word_tokens):
    """timing.py:108-113."""
    wb = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / TOKENS_PER_SECOND
    return jump_times[wb[:-1]], jump_times[wb[1:]], wb


def force_align(ws, tokens, tokenizer, aligned_unit_type="subword", aggregation="mean", topk=-1,
                w_colnorm=1.0, w_rownorm=1.0, w_coverage=0.0):
    """timing.py:69-114."""
    matrix, scores = aggregate(ws, aggregation, topk, w_colnorm, w_rownorm, w_coverage)
    matrix = matrix[len(tokenizer.sot_sequence) : -1].cpu()  # :102
    text_indices, time_indices = _dtw.dtw_path((-matrix).numpy())  # :103
    words, word_tokens = split_tokens_on_spaces(tokens + [tokenizer.eot], tokenizer, aligned_unit_type)  # :105
    if len(word_tokens) <= 1:
        return [[], [], [], [], None]  # :106-107, a LIST
    start_times, end_times, _ = boundaries_from_path(text_indices, time_indices, word_tokens)
    return words, start_times, end_times, matrix, scores


# --------------------------------------------------------------------------
# default_find_alignment, timing.py:116-186 (the stock-Whisper baseline behind
# --default_whisper_timing, infer_ali.py:83-85)
# --------------------------------------------------------------------------
def default_find_alignment(model, tokenizer, text_tokens, mel, max_frames, *, medfilt_width=7, qk_scale=1.0):
    """Only `model.alignment_heads` (:156), filtered softmax (:157-159), std/mean normalisation over
    tokens (:160-161), mean over heads (:163), DTW (:165) -- on the CPU recurrence, which is the
    parity target -- and word grouping with the tokenizer's own split_to_word_tokens (:167)."""
    tokens = torch.tensor([*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot])
    qk, _ = capture_logits(model, mel, tokens)  # (L, H, T, n_ctx)
    heads = model.alignment_heads.indices().T
    weights = torch.stack([qk[l][h] for l, h in heads])[:, :, :max_frames]
    weights = median_along_frames(weights, medfilt_width)
    weights = (weights * qk_scale).softmax(dim=-1)
    std, mean = torch.std_mean(weights, dim=-2, keepdim=True, unbiased=False)
    weights = (weights - mean) / std
    matrix = weights.mean(axis=0)[len(tokenizer.sot_sequence):-1]
    text_indices, time_indices = _dtw.dtw_path((-matrix).numpy())
    words, word_tokens = tokenizer.split_to_word_tokens(list(text_tokens) + [tokenizer.eot])
    if len(word_tokens) <= 1:
        return [[], [], [], [], None]
    start_times, end_times, _ = boundaries_from_path(text_indices, time_indices, word_tokens)
    return words, start_times, end_times, weights, None
