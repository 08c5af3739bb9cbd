"""ORACLE / TEST INFRASTRUCTURE ONLY -- digests of an (L, H, T, F) map tensor, shared by the golden
generator (oracle/gen_golden.py, which needs /root/reference) and the tests (which must not)."""
import numpy as np

SAMPLE_T, SAMPLE_F = 7, 11  # strides of the stored sample of the maps


def large_digests(w):
    """w: (L, H, T, F) numpy fp32 -> the digests stored for / compared on a large case."""
    w64 = w.astype(np.float64)
    return dict(
        weights_sample=w[:, :, ::SAMPLE_T, ::SAMPLE_F].copy(),
        weights_colsum=w64.sum(axis=2).astype(np.float32),       # (L, H, F)
        weights_sumsq=(w64 * w64).sum(axis=(2, 3)),              # (L, H) float64
        weights_argmax=w.argmax(axis=3).astype(np.int16),        # (L, H, T)
        weights_rowmax=w.max(axis=3),                            # (L, H, T)
    )
