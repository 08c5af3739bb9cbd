"""ORACLE / TEST INFRASTRUCTURE ONLY.

Import shim standing in for the third-party `openai-whisper` package, which the
reference imports (timing.py:7-10, infer_ali.py:18-20) but which is neither vendored
in /root/reference nor installable offline.  With this directory on sys.path the
reference's own timing.py / retokenize.py / metrics.py import and run unmodified.
"""
from . import audio, model, timing, tokenizer  # noqa: F401
from .audio import log_mel_spectrogram, pad_or_trim  # noqa: F401
from .model import ModelDimensions, Whisper, dims_for  # noqa: F401


def load_model(name: str, device=None, seed: int = 0):
    """No checkpoints offline: a seeded random-init model of the named size."""
    import torch

    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = Whisper(dims_for(name))
    with torch.no_grad():
        m.decoder.positional_embedding.normal_(0, 0.02)
    torch.random.set_rng_state(gen_state)
    return m if device is None else m.to(device)
