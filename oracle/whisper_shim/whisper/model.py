"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

CPU restatement of the slice of the third-party `openai-whisper` package
(`whisper/model.py`, un-vendored and un-pinned by the reference: README.md:8
`pip3 install -U openai-whisper`; must be >= v20240930 because the reference
imports `disable_sdpa` at timing.py:8) that the reference's hot path touches:

  * timing.py:48      model.dims.n_text_layer
  * timing.py:50-55   model.decoder.blocks[i].cross_attn forward hook, reads outs[-1]
                      (the PRE-softmax fp32 `qk` of shape (1, H, T, n_audio_ctx))
  * timing.py:57-58   `with disable_sdpa(): model(mel[None], tokens[None])`
  * timing.py:156     model.alignment_heads (sparse bool, upper half of the layers)

The module tree and parameter names match the published checkpoints
({"dims": ..., "model_state_dict": ...}) so a state dict moves freely between
this oracle, the product model and a real checkpoint.

Written from the published architecture description; nothing here is used by the
product. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
import it.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


class LayerNorm(nn.LayerNorm):
    # statistics are always taken in fp32, result cast back to the input dtype
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class Linear(nn.Linear):
    def forward(self, x):
        b = None if self.bias is None else self.bias.to(x.dtype)
        return F.linear(x, self.weight.to(x.dtype), b)


class Conv1d(nn.Conv1d):
    def _conv_forward(self, x, weight, bias):
        b = None if bias is None else bias.to(x.dtype)
        return super()._conv_forward(x, weight.to(x.dtype), b)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0):
    """Fixed sinusoidal table of the audio encoder: [sin | cos] halves."""
    assert channels % 2 == 0
    half = channels // 2
    step = math.log(max_timescale) / (half - 1)
    inv = torch.exp(-step * torch.arange(half))
    ang = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


class MultiHeadAttention(nn.Module):
    # class-level switch flipped by disable_sdpa(); the reference relies on it
    # (timing.py:46-47 "make sure MultiHeadAttention.use_sdpa = False").
    use_sdpa = True

    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)  # key projection carries no bias
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)

    def forward(self, x, xa=None, mask=None, kv_cache=None):
        src = x if xa is None else xa
        q = self.query(x)
        if kv_cache is not None and xa is not None and self.key in kv_cache:
            k, v = kv_cache[self.key], kv_cache[self.value]
        else:
            k, v = self.key(src), self.value(src)
        wv, qk = self.qkv_attention(q, k, v, mask)
        return self.out(wv), qk

    def qkv_attention(self, q, k, v, mask=None):
        _, n_ctx, n_state = q.shape
        # both operands are scaled by d_head^-1/4, so q.k carries d_head^-1/2
        scale = (n_state // self.n_head) ** -0.25

        def heads(t):
            return t.view(*t.shape[:2], self.n_head, -1).permute(0, 2, 1, 3)

        q, k, v = heads(q), heads(k), heads(v)
        if MultiHeadAttention.use_sdpa:
            causal = mask is not None and n_ctx > 1
            a = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
            return a.permute(0, 2, 1, 3).flatten(start_dim=2), None

        qk = (q * scale) @ (k * scale).transpose(-1, -2)
        if mask is not None:
            qk = qk + mask[:n_ctx, :n_ctx]
        qk = qk.float()
        w = F.softmax(qk, dim=-1).to(q.dtype)
        out = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
        return out, qk.detach()


@contextlib.contextmanager
def disable_sdpa():
    prev = MultiHeadAttention.use_sdpa
    MultiHeadAttention.use_sdpa = False
    try:
        yield
    finally:
        MultiHeadAttention.use_sdpa = prev


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_state: int, n_head: int, cross_attention: bool = False):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.cross_attn = MultiHeadAttention(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state) if cross_attention else None
        self.mlp = nn.Sequential(Linear(n_state, 4 * n_state), nn.GELU(), Linear(4 * n_state, n_state))
        self.mlp_ln = LayerNorm(n_state)

    def forward(self, x, xa=None, mask=None, kv_cache=None):
        x = x + self.attn(self.attn_ln(x), mask=mask, kv_cache=kv_cache)[0]
        if self.cross_attn is not None:
            x = x + self.cross_attn(self.cross_attn_ln(x), xa, kv_cache=kv_cache)[0]
        return x + self.mlp(self.mlp_ln(x))


class AudioEncoder(nn.Module):
    def __init__(self, n_mels, n_ctx, n_state, n_head, n_layer):
        super().__init__()
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks = nn.ModuleList([ResidualAttentionBlock(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = LayerNorm(n_state)

    def forward(self, x):
        x = F.gelu(self.conv1(x))
        x = F.gelu(self.conv2(x))
        x = x.permute(0, 2, 1)
        assert x.shape[1:] == self.positional_embedding.shape, "incorrect audio shape"
        x = (x + self.positional_embedding).to(x.dtype)
        for blk in self.blocks:
            x = blk(x)
        return self.ln_post(x)


class TextDecoder(nn.Module):
    def __init__(self, n_vocab, n_ctx, n_state, n_head, n_layer):
        super().__init__()
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        self.positional_embedding = nn.Parameter(torch.empty(n_ctx, n_state))
        self.blocks = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head, cross_attention=True) for _ in range(n_layer)]
        )
        self.ln = LayerNorm(n_state)
        mask = torch.full((n_ctx, n_ctx), float("-inf")).triu_(1)
        self.register_buffer("mask", mask, persistent=False)

    def forward(self, x, xa, kv_cache=None):
        offset = next(iter(kv_cache.values())).shape[1] if kv_cache else 0
        x = self.token_embedding(x) + self.positional_embedding[offset : offset + x.shape[-1]]
        x = x.to(xa.dtype)
        for blk in self.blocks:
            x = blk(x, xa, mask=self.mask, kv_cache=kv_cache)
        x = self.ln(x)
        # tied output embedding, logits always fp32
        return (x @ self.token_embedding.weight.to(x.dtype).transpose(0, 1)).float()


class Whisper(nn.Module):
    def __init__(self, dims: ModelDimensions):
        super().__init__()
        self.dims = dims
        self.encoder = AudioEncoder(
            dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer
        )
        self.decoder = TextDecoder(
            dims.n_vocab, dims.n_text_ctx, dims.n_text_state, dims.n_text_head, dims.n_text_layer
        )
        # default alignment heads: every head of the upper half of the decoder
        heads = torch.zeros(dims.n_text_layer, dims.n_text_head, dtype=torch.bool)
        heads[dims.n_text_layer // 2 :] = True
        self.register_buffer("alignment_heads", heads.to_sparse(), persistent=False)

    def embed_audio(self, mel):
        return self.encoder(mel)

    def logits(self, tokens, audio_features):
        return self.decoder(tokens, audio_features)

    def forward(self, mel, tokens):
        return self.decoder(tokens, self.encoder(mel))

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def is_multilingual(self):
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self):
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)


# Published size table (n_mels, audio ctx/state/head/layer, vocab, text ctx/state/head/layer).
_DIMS = {
    "tiny": (80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base": (80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small": (80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium": (80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v3": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
}


def dims_for(name: str) -> ModelDimensions:
    return ModelDimensions(*_DIMS[name])
