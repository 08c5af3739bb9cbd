"""Maps the CUDA toolkit's cuBLAS 12.9 before torch so that CUBLAS_EMULATE_SINGLE_PRECISION works
(see bench.py::_preload_cublas_emulation).  `import _preload` first; WCA_FP32_GEMM=native disables."""
import ctypes
import os

MODE = "native"
if os.environ.get("WCA_FP32_GEMM", "bf16x9") != "native":
    try:
        os.environ.setdefault("CUBLAS_EMULATE_SINGLE_PRECISION", "1")
        for _n in ("libcublasLt.so.12", "libcublas.so.12"):
            ctypes.CDLL(os.path.join(os.environ.get("WCA_CUBLAS_DIR", "/usr/local/cuda/lib64"), _n), mode=ctypes.RTLD_GLOBAL)
        MODE = "bf16x9"
    except OSError:
        os.environ.pop("CUBLAS_EMULATE_SINGLE_PRECISION", None)
