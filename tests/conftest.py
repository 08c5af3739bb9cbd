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


# Measured parity figures (maximum relative map error, ...) are part of the evidence: tests call
# `record_measure` and the figures are printed at the end of the run (also with -q) and written to
# gpurun_out/parity_measures[_<mode>].json when that directory exists.
MEASURES = {}
GEMM_MODE = os.environ.get("WCA_FP32_GEMM") or "native"


def record_measure(name, value):
    MEASURES[name] = max(float(value), MEASURES.get(name, 0.0))


def pytest_terminal_summary(terminalreporter):
    if not MEASURES:
        return
    terminalreporter.write_line(f"measured parity figures (fp32 GEMM mode: {GEMM_MODE})")
    for k in sorted(MEASURES):
        terminalreporter.write_line(f"  MEASURE[{GEMM_MODE}] {k} = {MEASURES[k]:.3e}")
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, f"parity_measures_{GEMM_MODE}.json")
        prev = {}
        if os.path.exists(path):
            try:
                prev = json.load(open(path))
            except ValueError:
                prev = {}
        prev.update(MEASURES)
        json.dump(prev, open(path, "w"), indent=1, sort_keys=True)


def max_rel_err(got, want, atol=1e-7, rtol=1e-4):
    """max |got - want| / max(|want|, atol / rtol): the quantity assert_allclose(rtol, atol) bounds by ~rtol."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float((np.abs(got - want) / np.maximum(np.abs(want), atol / rtol)).max()) if got.size else 0.0


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
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("default_", "large_", "aux_")))


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


def large_names():
    """Reference-generated fixtures at LibriSpeech-class shapes (T ~ 200-400, F 1100-1500): cost matrix, path,
    times and scores in full, the maps as a strided sample + digests (oracle/gen_golden.py main_large)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("large_") and f.endswith(".npz"))


def load_large(name):
    import torch

    from oracle.synth import make_dims, make_mel

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["case"] = c = json.loads(str(g["case"]))
    g["words"] = json.loads(str(g["words"]))
    dims = make_dims(c["model"])
    mel = make_mel(dims.n_mels, 2 * dims.n_audio_ctx, 2 * c["frames"], seed=c["mel_seed"])
    # the mel is re-created from its seed: same torch CPU generator, checked against the stored digest
    assert abs(mel.double().sum().item() - g["mel_digest"][0]) < 1e-6 * g["mel_digest"][1]
    assert abs(mel.abs().double().sum().item() - g["mel_digest"][1]) < 1e-9 * g["mel_digest"][1]
    g["mel"] = mel.numpy()
    g["path_text"] = g["path_text"].astype(np.int64)
    g["path_time"] = g["path_time"].astype(np.int64)
    return g


def check_large_maps(w, g, rtol, label=""):
    """w: (L, H, T, F) numpy maps against the sample and digests of a large fixture.  Returns the measured
    maximum relative difference on the sample."""
    from oracle.gen_golden_digests import SAMPLE_F, SAMPLE_T, large_digests

    assert tuple(w.shape) == tuple(g["weights_shape"])
    d = large_digests(w)
    ref = g["weights_sample"]
    err = float((np.abs(d["weights_sample"] - ref) / np.maximum(np.abs(ref), 1e-7 / rtol)).max())
    print(f"{label}: max relative map difference on the {SAMPLE_T}x{SAMPLE_F}-strided sample = {err:.3e}")
    np.testing.assert_allclose(d["weights_sample"], ref, rtol=rtol, atol=1e-7)
    np.testing.assert_allclose(d["weights_colsum"], g["weights_colsum"], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(d["weights_sumsq"], g["weights_sumsq"], rtol=2 * rtol)
    np.testing.assert_allclose(d["weights_rowmax"], g["weights_rowmax"], rtol=rtol, atol=1e-7)
    # arg-max of a row may legitimately differ where two frames tie within the tolerance
    same = d["weights_argmax"] == g["weights_argmax"]
    if not same.all():
        l, h, t = np.nonzero(~same)
        a = w[l, h, t, g["weights_argmax"][l, h, t].astype(np.int64)]
        np.testing.assert_allclose(a, g["weights_rowmax"][l, h, t], rtol=rtol, atol=1e-7)
    return err


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
