"""Boundary-scoring helpers with the reference's names and semantics (metrics.py:
eval_n1 :22-43, eval_n1_strict :45-72, get_seg_metrics :74-86, coverage_penalty :99-111).
Host-side and tiny; the device-side coverage term lives in csrc/scores.cu."""
from __future__ import annotations

import string

import numpy as np
import torch


def eval_n1(y, yhat, tolerance=1):
    """Greedy in-order matching of two sorted boundary lists; returns (hits, hits)."""
    if len(yhat) == 0:
        return 0, 0
    hits = i = j = 0
    while i < len(y) and j < len(yhat):
        if abs(y[i] - yhat[j]) <= tolerance:
            hits += 1
            i += 1
            j += 1
        elif y[i] < yhat[j]:
            i += 1
        elif y[i] > yhat[j]:
            j += 1
        else:  # NaN never orders: stop instead of spinning
            break
    return hits, hits


def eval_n1_strict(y, y_hat, words, words_hat, tolerance=1):
    """A prediction counts when an unused reference boundary within `tolerance` carries the
    same (lower-cased, punctuation-stripped) word.  Returns (tp, fp, fn)."""
    ref_words = [w.lower().strip(string.punctuation) for w in words]
    hyp_words = [w.lower().strip(string.punctuation) for w in words_hat]
    taken = set()
    tp = 0
    for i in range(len(y_hat)):
        for j in range(len(y)):
            if j not in taken and ref_words[j] == hyp_words[i] and abs(y[j] - y_hat[i]) <= tolerance:
                taken.add(j)
                tp += 1
                break
    return tp, len(y_hat) - tp, len(y) - len(taken)


def get_seg_metrics(correct_predict, correct_retrieve, total_predict, total_gold):
    """(precision, recall, f1, r_value, over-segmentation)."""
    eps = 1e-7
    precision = correct_predict / (total_predict + eps)
    recall = correct_retrieve / (total_gold + eps)
    f1 = 2 * (precision * recall) / (precision + recall + eps)
    over_seg = recall / (precision + eps) - 1
    r1 = np.sqrt((1 - recall) ** 2 + over_seg ** 2)
    r2 = (-over_seg + recall - 1) / np.sqrt(2)
    return precision, recall, f1, 1 - (abs(r1) + abs(r2)) / 2, over_seg


def coverage_penalty(attn, threshold=0.5):
    """attn (tokens, frames): sum_f max(sum_t attn[t, f], threshold) - F * threshold."""
    covered = attn.sum(dim=0)
    return torch.clamp_min(covered, threshold).sum(-1) - covered.size(-1) * threshold
