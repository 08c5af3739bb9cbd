import json
import os
import sys


# WCA_FP32_GEMM=bf16x9 runs the GPU suite with the upstream linears on cuBLAS 12.9's BF16x9-emulated
# fp32 GEMMs (what bench.py uses by default); the libraries must be mapped before torch is imported.
if os.environ.get("WCA_FP32_GEMM") == "bf16x9":
    import ctypes

    os.environ.setdefault("CUBLAS_EMULATE_SINGLE_PRECISION", "1")
    for _name in ("libcublasLt.so.12", "libcublas.so.12"):
        ctypes.CDLL(os.path.join(os.environ.get("WCA_CUBLAS_DIR", "/usr/local/cuda/lib64"), _name), mode=ctypes.RTLD_GLOBAL)

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names():
    """Fixtures of get_attentions + force_align (the default_find_alignment ones are listed separately)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("default_"))


def default_timing_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("default_") and f.endswith(".npz"))


def load_golden(name):
    """Fixture produced by oracle/gen_golden.py from the reference's own timing.py."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["case"] = json.loads(str(g["case"]))
    if "words" in g:
        g["words"] = json.loads(str(g["words"]))
    g["sentinel"] = bool(g["sentinel"])
    if "mel_pad_value" in g:  # c1: only the speech part of the mel is stored
        full = np.full((g["mel"].shape[0], 3000), g["mel_pad_value"], dtype=np.float32)
        full[:, : g["mel"].shape[1]] = g["mel"]
        g["mel"] = full
    return g


@pytest.fixture(scope="session")
def oracle_models():
    """Seeded oracle (CPU) models, shared across tests; weights are regenerated, never stored."""
    from oracle.synth import make_model

    cache = {}

    def get(name, seed=0, gain=4.0):
        key = (name, seed, gain)
        if key not in cache:
            cache[key] = make_model(name, seed, gain)
        return cache[key]

    return get


@pytest.fixture(scope="session")
def tokenizer():
    from whisper_char_alignment_b200.tokenizer import get_tokenizer

    return get_tokenizer(True, language="English")
