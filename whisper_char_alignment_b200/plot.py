"""Attention heat-map with reference / predicted boundaries (reference plot.py:22-59).  matplotlib is
optional and imported lazily: the reference imports it unconditionally and cannot start without it."""
from __future__ import annotations

import os

import numpy as np

from .retokenize import split_tokens_on_spaces


def plot_attn(weights, text_tokens, tokenizer, gt_alignment, pred_alignment, fid, aligned_unit_type, path):
    import matplotlib

    matplotlib.use("Agg")
    import matplotlib.pyplot as plt

    os.makedirs(path, exist_ok=True)
    fig, ax = plt.subplots(figsize=(8, 3.5))
    ax.imshow(weights.detach().cpu().numpy(), aspect="auto")
    for e in gt_alignment or []:
        ax.axvline(int(e / 0.02), linewidth=2, color="white")
    for e in pred_alignment:
        ax.axvline(int(e / 0.02), linewidth=3, color="cyan" if aligned_unit_type == "subword" else "red", ls="dotted")
    _, word_tokens = split_tokens_on_spaces(list(text_tokens) + [tokenizer.eot], tokenizer, aligned_unit_type)
    for b in np.cumsum([len(w) for w in word_tokens[:-1]]):
        ax.axhline(b - 0.5, linewidth=1.5, color="gray", ls="--")
    ax.set_yticks(np.arange(len(weights) - 1, -1, -1))
    ax.set_yticklabels(([tokenizer.decode([t]) for t in text_tokens] + [""])[::-1], fontsize=9)
    ax.set_xticks([])
    plt.xlabel(r"${time} (\rightarrow)$", fontsize=18)
    plt.tight_layout()
    plt.savefig(os.path.join(path, f"{fid}.png"), bbox_inches="tight", dpi=400)
    plt.close(fig)
