"""ORACLE / TEST INFRASTRUCTURE ONLY.

Restated constants and front-end helpers of `openai-whisper`'s `whisper/audio.py`
(third party, absent from /root/reference).  The reference reads
TOKENS_PER_SECOND at timing.py:10/:111, HOP_LENGTH at infer_ali.py:179, and calls
pad_or_trim / log_mel_spectrogram at dataset.py:47-48.

`mel_filters.npz` is not available offline, so the Slaney-normalised filterbank is
computed from its closed form.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 mel frames
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2  # conv stride 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH  # 100
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN  # 50


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = F.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log_region = f >= 1000.0
    with np.errstate(divide="ignore", invalid="ignore"):
        logv = 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(log_region, logv, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * (200.0 / 3)
    logv = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return np.where(m >= 15.0, logv, lin)


def mel_filters(n_mels: int) -> torch.Tensor:
    """Slaney-style triangular filterbank (area-normalised), (n_mels, N_FFT//2+1)."""
    n_freqs = N_FFT // 2 + 1
    fft_f = np.linspace(0.0, SAMPLE_RATE / 2, n_freqs)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SAMPLE_RATE / 2), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - fft_f[None, :]
    fb = np.zeros((n_mels, n_freqs))
    for i in range(n_mels):
        lo = -ramps[i] / width[i]
        hi = ramps[i + 2] / width[i + 1]
        fb[i] = np.maximum(0.0, np.minimum(lo, hi))
    fb *= (2.0 / (edges[2 : n_mels + 2] - edges[:n_mels]))[:, None]
    return torch.from_numpy(fb.astype(np.float32))


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, device=None):
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.asarray(audio))
    audio = audio.float()
    if device is not None:
        audio = audio.to(device)
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT).to(audio.device)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    power = stft[..., :-1].abs() ** 2
    mel = mel_filters(n_mels).to(audio.device) @ power
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
