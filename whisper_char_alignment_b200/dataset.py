"""Utterance sources for the CLI shells.  Same record as the reference's datasets yield
(dataset.py:38-50,100-122): (audio, mel, duration_in_samples, text, starts, ends, fid).

  TIMIT        scp lines "fid path/to/x.wav" (NIST SPHERE) with the word labels in x.wrd
  LibriSpeech  scp lines + `ls_alignment_<split>.txt`; needs a FLAC decoder (soundfile), absent offline
  synthetic    BASELINE.json-shaped random utterances with made-up reference boundaries (offline runs)
"""
from __future__ import annotations

import ast
import os
from glob import glob

import numpy as np
import torch

from . import audio, synthetic


class TIMIT(torch.utils.data.Dataset):
    def __init__(self, scp_file="scp/test.wav.scp", n_mels=80, device="cpu"):
        self.n_mels, self.device, self.items = n_mels, device, []
        for line in open(scp_file):
            parts = line.split()
            if len(parts) < 2:
                continue
            fid, wav = parts[0], parts[1]
            self.items.append((fid, wav, wav.split(".wav")[0] + ".wrd"))

    def __len__(self):
        return len(self.items)

    @staticmethod
    def read_words(path, sample_rate=audio.SAMPLE_RATE):
        starts, ends, words = [], [], []
        for line in open(path):
            a, b, w = line.split()[:3]
            starts.append(float(a) / sample_rate)
            ends.append(float(b) / sample_rate)
            words.append(w)
        return " ".join(words), starts, ends

    def size_hint(self, i):
        """(approximate token count, approximate frame count) without decoding the audio: for cost-balanced
        sharding and length bucketing (16-bit mono PCM behind a 1024-byte SPHERE header; ~14 characters/s)."""
        n_samples = max(os.path.getsize(self.items[i][1]) - 1024, 0) // 2
        return int(14 * n_samples / audio.SAMPLE_RATE) + 5, n_samples // audio.N_SAMPLES_PER_TOKEN

    def __getitem__(self, i):
        fid, wav, wrd = self.items[i]
        pcm, rate = audio.read_audio(wav)
        assert rate == audio.SAMPLE_RATE, f"{wav}: {rate} Hz"
        text, starts, ends = self.read_words(wrd) if os.path.exists(wrd) else ("", [], [])
        samples = torch.from_numpy(pcm.copy())
        mel = audio.log_mel_spectrogram(audio.pad_or_trim(samples), self.n_mels, device=self.device)
        return samples, mel, len(pcm), text, starts, ends, fid


class LibriSpeech(torch.utils.data.Dataset):
    def __init__(self, scp_file="scp/dev-clean.wav.scp", n_mels=80, device="cpu"):
        try:
            import soundfile  # noqa: F401
        except ImportError as exc:  # pragma: no cover - depends on the image
            raise RuntimeError("LibriSpeech needs a FLAC decoder (`soundfile`), which this image does not have") from exc
        self.n_mels, self.device = n_mels, device
        lines = [l.split() for l in open(scp_file) if l.strip()]
        split = lines[0][1].split("/")[-4]
        root = lines[0][1].split(split)[0]
        labels = {}
        for trans in sorted(glob(os.path.join(root, split, "**/*.trans.txt"), recursive=True)):
            for l in open(trans):
                fid, text = l.split(" ", 1)
                labels[fid] = text
        alignments = {}
        for l in open(f"ls_alignment_{split}.txt"):
            fid, rest = l.split(" ", 1)
            alignments[fid] = ast.literal_eval(rest)
        self.items = [(fid, path, labels[fid], alignments[fid]) for fid, path in (x[:2] for x in lines)]

    def __len__(self):
        return len(self.items)

    def size_hint(self, i):
        """From the transcript alone (FLAC is not decoded here): characters, and frames at ~14 characters/s."""
        n_chars = len(self.items[i][2])
        return n_chars + 5, int(n_chars / 14.0 * audio.TOKENS_PER_SECOND)

    def __getitem__(self, i):
        import soundfile

        fid, path, _, ali = self.items[i]
        pcm, rate = soundfile.read(path, dtype="float32")
        assert rate == audio.SAMPLE_RATE
        samples = torch.from_numpy(pcm)
        mel = audio.log_mel_spectrogram(audio.pad_or_trim(samples), self.n_mels, device=self.device)
        words = [(w, s, e) for w, s, e in ali if w != ""]
        return (samples, mel, len(pcm), " ".join(w for w, _, _ in words), [s for _, s, _ in words],
                [e for _, _, e in words], fid)


class Synthetic(torch.utils.data.Dataset):
    """`scp_file` selects the shape: timit | librispeech | ami | probe, optionally `name:count`."""

    def __init__(self, scp_file="timit:64", n_mels=80, device="cpu", tokenizer=None, seed=0):
        from .tokenizer import get_tokenizer

        name, _, count = scp_file.partition(":")
        tk = tokenizer or get_tokenizer(True, language="English")
        self.device = device
        self.utts = synthetic.WORKLOADS[name](int(count or 64), tk, n_mels=n_mels, seed=seed)
        rng = np.random.default_rng(seed + 1)
        self.labels = []
        for u in self.utts:  # made-up, monotone reference boundaries: one per word
            n_words = len(u.text.split())
            cuts = np.sort(rng.uniform(0, u.n_samples / audio.SAMPLE_RATE, n_words + 1))
            self.labels.append((cuts[:-1].tolist(), cuts[1:].tolist()))

    def __len__(self):
        return len(self.utts)

    def size_hint(self, i):
        return len(self.utts[i].tokens), self.utts[i].max_frames

    def __getitem__(self, i):
        u = self.utts[i]
        starts, ends = self.labels[i]
        return None, u.mel.to(self.device), u.n_samples, u.text, starts, ends, u.fid


DATASET = {"TIMIT": TIMIT, "LibriSpeech": LibriSpeech, "synthetic": Synthetic}
