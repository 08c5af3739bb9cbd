"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement of the reference's alignment hot path (`timing.get_attentions` +
`timing.force_align`, /root/reference/timing.py:13-114) and of the third-party
`openai-whisper` routines it calls (un-vendored, un-pinned; see whisper_shim/).

Parity status: the reference ships no tests and no golden vectors, so parity is
pinned by running the reference's OWN timing.py / retokenize.py / metrics.py,
unmodified, against `whisper_shim/` in the build container and committing the
outputs as fixtures (tests/golden/, generator: oracle/gen_golden.py).  The
restatement in ref_path.py and dtw_oracle.c is checked bit-for-bit against those
fixtures by the CPU test-suite.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (whisper_char_alignment_b200/) never does.
"""
import os
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(ORACLE_DIR, "whisper_shim")


def use_shim():
    """Make `import whisper` / `import num2words` resolve to the restated shim."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
