"""Synthetic-shape generators for the BASELINE.json configurations (no corpora or
checkpoints are reachable offline).  Deterministic given the seed.

    C2  TIMIT-shaped      : 1680 utts, 2-4 s, ~40 chars
    C3  LibriSpeech-shaped: 2620 utts, 2-30 s, chars proportional to duration (T <= 448)
    C4  AMI-shaped        : 1-6 s, 5-25 subword tokens (large-v3, 128 mels)
    C5  probe sweep       : >= 18 words per utterance
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .retokenize import encode

SAMPLES_PER_FRAME = 320  # HOP_LENGTH * 2 (infer_ali.py:179)
N_MEL_FRAMES = 3000
MAX_FRAMES = 1500        # infer_ali.py:25
MAX_LENGTH = 448         # infer_ali.py:26

_WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at which but have an "
          "had they you were their one all we can her has there been if more when will would who so no out up "
          "speech model align word frame token whisper audio signal time path cost head layer").split()


@dataclass
class Utterance:
    fid: str
    mel: torch.Tensor          # (n_mels, 3000) fp32, zero beyond the utterance
    n_samples: int
    text: str
    text_tokens: list          # text tokens only (what force_align takes)
    tokens: torch.Tensor       # [*sot_sequence, no_timestamps, *text_tokens, eot]
    max_frames: int


def _sentence(rng, n_chars: int, min_words: int = 1) -> str:
    words, total = [], 0
    while total < n_chars or len(words) < min_words:
        w = _WORDS[int(rng.integers(len(_WORDS)))]
        words.append(w)
        total += len(w) + 1
    return " ".join(words)


def make_utterance(rng, tokenizer, n_mels, seconds, n_chars, unit, fid, min_words=1, max_subwords=None,
                   with_mel=True) -> Utterance:
    n_samples = int(seconds * 16000)
    max_frames = min(n_samples // SAMPLES_PER_FRAME, MAX_FRAMES)
    text = _sentence(rng, n_chars, min_words)
    text_tokens = encode(text, tokenizer, unit)
    budget = MAX_LENGTH - len(tokenizer.sot_sequence) - 2
    if max_subwords is not None:
        budget = min(budget, max_subwords)
    if len(text_tokens) > budget:
        text_tokens = text_tokens[:budget]
        text = tokenizer.decode(text_tokens)
    mel = None  # with_mel=False: shape descriptors only (the caller creates the mel, e.g. directly in HBM)
    if with_mel:
        mel = torch.zeros(n_mels, N_MEL_FRAMES)
        n_mel = min(2 * max_frames, N_MEL_FRAMES)
        noise = rng.standard_normal((n_mels, n_mel)).astype(np.float32) * 0.3
        mel[:, :n_mel] = torch.from_numpy(noise)
    tokens = torch.tensor([*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot])
    return Utterance(fid, mel, n_samples, text, text_tokens, tokens, max_frames)


def timit_shaped(n, tokenizer, n_mels=80, seed=0):
    rng = np.random.default_rng(seed)
    return [make_utterance(rng, tokenizer, n_mels, rng.uniform(2.0, 4.0), int(rng.integers(30, 51)), "char", f"timit{i:04d}")
            for i in range(n)]


def librispeech_shaped(n, tokenizer, n_mels=80, seed=0, with_mel=True):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        sec = rng.uniform(2.0, 30.0)
        out.append(make_utterance(rng, tokenizer, n_mels, sec, int(sec * 13.5), "char", f"libri{i:04d}", with_mel=with_mel))
    return out


def ami_shaped(n, tokenizer, n_mels=128, seed=0):
    rng = np.random.default_rng(seed)
    return [make_utterance(rng, tokenizer, n_mels, rng.uniform(1.0, 6.0), int(rng.integers(10, 50)), "subword",
                           f"ami{i:04d}", max_subwords=25) for i in range(n)]


def probe_shaped(n, tokenizer, n_mels=80, seed=0):
    rng = np.random.default_rng(seed)
    return [make_utterance(rng, tokenizer, n_mels, rng.uniform(4.0, 8.0), 90, "char", f"probe{i:04d}", min_words=18)
            for i in range(n)]


WORKLOADS = {"timit": timit_shaped, "librispeech": librispeech_shaped, "ami": ami_shaped, "probe": probe_shaped}
