"""ORACLE / TEST INFRASTRUCTURE ONLY: seeded model + input builders shared by the
golden generator and the tests (weights are never committed; they are re-created
from the seed with the same torch build)."""
from __future__ import annotations

import torch

from . import use_shim

use_shim()
from whisper.model import ModelDimensions, Whisper, dims_for  # noqa: E402

MICRO = dict(n_mels=80, n_audio_ctx=256, n_audio_state=128, n_audio_head=2, n_audio_layer=2,
             n_vocab=51865, n_text_ctx=448, n_text_state=128, n_text_head=2, n_text_layer=2)
MINI = dict(n_mels=80, n_audio_ctx=512, n_audio_state=256, n_audio_head=4, n_audio_layer=2,
            n_vocab=51865, n_text_ctx=448, n_text_state=256, n_text_head=4, n_text_layer=3)


# full audio context (1500 frames) at a width the CPU reference finishes in seconds: the shape class of
# BASELINE.json configs[2] (LibriSpeech: T up to 448, F up to 1500) for reference-generated fixtures
LONG = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=256, n_audio_head=4, n_audio_layer=2,
            n_vocab=51865, n_text_ctx=448, n_text_state=256, n_text_head=4, n_text_layer=3)


def make_dims(name: str) -> ModelDimensions:
    if name == "micro":
        return ModelDimensions(**MICRO)
    if name == "mini":
        return ModelDimensions(**MINI)
    if name == "long":
        return ModelDimensions(**LONG)
    return dims_for(name)


def make_model(name: str, seed: int = 0, qk_gain: float = 4.0) -> Whisper:
    """Random-init model of the named size.  `qk_gain` multiplies the cross-attention
    query/key weights so that the maps are peaky enough for DTW paths to be decided by
    the data, not by last-bit noise ('random-init fragility')."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = Whisper(make_dims(name))
    with torch.no_grad():
        model.decoder.positional_embedding.normal_(0, 0.02)
        for blk in model.decoder.blocks:
            blk.cross_attn.query.weight.mul_(qk_gain)
            blk.cross_attn.query.bias.mul_(qk_gain)
            blk.cross_attn.key.weight.mul_(qk_gain)
    torch.random.set_rng_state(state)
    return model.eval()


def long_text(n_chars: int, seed: int) -> str:
    """Seeded lowercase pseudo-sentence of about n_chars characters (3-9 letter words)."""
    g = torch.Generator().manual_seed(seed)
    words, total = [], 0
    while total < n_chars:
        n = int(torch.randint(3, 10, (1,), generator=g))
        w = "".join(chr(97 + int(c)) for c in torch.randint(0, 26, (n,), generator=g))
        words.append(w)
        total += n + 1
    return " ".join(words)[:n_chars].rstrip()


def make_mel(n_mels: int, n_frames_total: int, n_frames_speech: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(n_mels, n_frames_total, generator=g) * 0.3
    mel[:, n_frames_speech:] = 0
    return mel
