"""Audio front-end used by the CLI shells: NIST-SPHERE / RIFF-WAV
readers, pad_or_trim and the Whisper log-mel spectrogram (STFT on cuFFT when the tensor is on
the GPU).  Constants are those of `whisper.audio` that the reference reads (timing.py:10,
infer_ali.py:179, dataset.py:47-48).  `mel_filters.npz` is not available offline, so the
Slaney-normalised triangular filterbank is built from its closed form."""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = N_SAMPLES // HOP_LENGTH
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def read_audio(path: str):
    """(float32 mono samples in [-1, 1), sample_rate) from a NIST SPHERE file (TIMIT's `.wav`) or a
    16-bit PCM RIFF file."""
    raw = open(path, "rb").read()
    if raw[:7] == b"NIST_1A":
        header_len = int(raw[8:16].split()[0])
        fields = {}
        for line in raw[16:header_len].decode("ascii", "replace").splitlines():
            parts = line.split(None, 2)
            if len(parts) == 3:
                fields[parts[0]] = parts[2]
        if fields.get("sample_coding", "pcm") not in ("pcm", "pcm,embedded-shorten-v2.00"[:3]):
            raise ValueError(f"{path}: unsupported SPHERE coding {fields.get('sample_coding')}")
        order = "<i2" if fields.get("sample_byte_format", "01") == "01" else ">i2"
        pcm = np.frombuffer(raw[header_len:], dtype=order)
        channels = int(fields.get("channel_count", 1))
        if channels > 1:
            pcm = pcm.reshape(-1, channels).mean(axis=1)
        return pcm.astype(np.float32) / 32768.0, int(fields.get("sample_rate", SAMPLE_RATE))
    if raw[:4] == b"RIFF" and raw[8:12] == b"WAVE":
        pos, fmt, data = 12, None, None
        while pos + 8 <= len(raw):
            tag, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
            body = raw[pos + 8:pos + 8 + size]
            if tag == b"fmt ":
                fmt = struct.unpack("<HHIIHH", body[:16])
            elif tag == b"data":
                data = body
            pos += 8 + size + (size & 1)
        if fmt is None or data is None or fmt[0] != 1 or fmt[5] != 16:
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported")
        pcm = np.frombuffer(data, dtype="<i2")
        if fmt[1] > 1:
            pcm = pcm.reshape(-1, fmt[1]).mean(axis=1)
        return pcm.astype(np.float32) / 32768.0, int(fmt[2])
    raise ValueError(f"{path}: neither NIST SPHERE nor RIFF/WAVE")


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Zero-pad or cut `axis` to exactly `length` samples (30 s by default)."""
    is_tensor = torch.is_tensor(array)
    x = array if is_tensor else torch.from_numpy(np.asarray(array))
    n = x.shape[axis]
    if n > length:
        x = x.narrow(axis, 0, length)
    elif n < length:
        pad = [0, 0] * x.ndim
        pad[2 * (x.ndim - 1 - (axis % x.ndim)) + 1] = length - n
        x = F.pad(x, pad)
    return x if is_tensor else x.numpy()


def _mel_scale(hz):
    hz = np.asarray(hz, dtype=np.float64)
    lin = hz * 3.0 / 200.0
    log = 15.0 + 27.0 * np.log(np.maximum(hz, 1e-10) / 1000.0) / np.log(6.4)
    return np.where(hz >= 1000.0, log, lin)


def _mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (mel - 15.0)), mel * 200.0 / 3.0)


_FILTER_CACHE = {}


def mel_filters(n_mels: int, device=None) -> torch.Tensor:
    key = (n_mels, str(device))
    if key not in _FILTER_CACHE:
        freqs = np.linspace(0.0, SAMPLE_RATE / 2, N_FFT // 2 + 1)
        edges = _mel_to_hz(np.linspace(_mel_scale(0.0), _mel_scale(SAMPLE_RATE / 2), n_mels + 2))
        rising = (freqs[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
        falling = (edges[2:, None] - freqs[None, :]) / (edges[2:] - edges[1:-1])[:, None]
        bank = np.clip(np.minimum(rising, falling), 0.0, None) * (2.0 / (edges[2:] - edges[:-2]))[:, None]
        _FILTER_CACHE[key] = torch.from_numpy(bank.astype(np.float32)).to(device)
    return _FILTER_CACHE[key]


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, device=None) -> torch.Tensor:
    """(n_mels, n_samples // 160) log10 mel power, clamped to 8 dB below the peak, mapped to ~[-1, 1]."""
    x = audio if torch.is_tensor(audio) else torch.from_numpy(np.asarray(audio))
    x = x.float()
    if device is not None:
        x = x.to(device)
    if padding:
        x = F.pad(x, (0, padding))
    spec = torch.stft(x, N_FFT, HOP_LENGTH, window=torch.hann_window(N_FFT, device=x.device), return_complex=True)
    power = spec[..., :-1].abs().square()
    mel = mel_filters(n_mels, x.device) @ power
    log_mel = mel.clamp_min(1e-10).log10()
    log_mel = torch.maximum(log_mel, log_mel.max() - 8.0)
    return (log_mel + 4.0) / 4.0
