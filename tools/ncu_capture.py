"""One capture launch at a BASELINE.json shape, for ncu.  usage: ncu_capture.py [timit|libri|ami] [B] [simt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi
if os.environ.get("WCA_LIB"):  # A/B timing against another build of the library
    _cabi.LIB_PATH = os.path.abspath(os.environ["WCA_LIB"])

shape = sys.argv[1] if len(sys.argv) > 1 else "timit"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
simt = len(sys.argv) > 3 and sys.argv[3] == "simt"
reps = int(os.environ.get("WCA_REPS", "20"))
dev = torch.device("cuda:0")
L, H, D, n_ctx = 24, 16, 64, 1500
width = 3
rng = np.random.default_rng(0)
if shape == "timit":
    Ts = rng.integers(35, 56, B); Fs = rng.integers(100, 200, B)
elif shape == "ami":  # BASELINE configs[3]: large-v3 dims, short segments, subword tokens, medfilt 7
    L, H, width = 32, 20, 7
    Ts = rng.integers(10, 31, B); Fs = rng.integers(50, 300, B)
else:
    Fs = rng.integers(100, 1500, B); Ts = np.minimum(448, (Fs * 0.27).astype(int) + 5)
t_max = int(Ts.max())
q = [torch.randn(B, t_max, H * D, device=dev) for _ in range(L)]
k = [torch.randn(B, n_ctx, H * D, device=dev) for _ in range(L)]
recs = np.zeros(B, dtype=_cabi.UTT_DTYPE)
off = 0
for b in range(B):
    recs[b]["n_tokens"], recs[b]["n_frames"] = Ts[b], Fs[b]
    recs[b]["q_row0"], recs[b]["k_row0"], recs[b]["ws_off"] = b * t_max, b * n_ctx, off
    off += L * H * int(Ts[b]) * int(Fs[b])
flags = (_cabi.WCA_CAPTURE_FORCE_SIMT if simt else 0) | int(os.environ.get("WCA_DBG", "0"), 0)  # bits 8..: experiment switches
from whisper_char_alignment_b200.timing import _cluster_bucket
buckets = {}
for b in range(B):
    buckets.setdefault(1 if simt else _cluster_bucket(int(Fs[b])), []).append(b)
launches = [(recs[m], _cabi.upload_utts(recs[m], dev)) for _, m in sorted(buckets.items())]
ws = torch.empty(off, device=dev)
partials = None
if os.environ.get("WCA_PARTIALS", "0") == "1" and not simt:  # head-score partials written by the score warps
    poff = 0
    for b in range(B):
        recs[b]["part_off"] = poff
        poff += _cabi.capture_partials_floats(L * H, int(Ts[b]), int(Fs[b]))
    partials = torch.empty(poff, device=dev)
    launches = [(recs[m], _cabi.upload_utts(recs[m], dev)) for _, m in sorted(buckets.items())]

def capture():
    for sub, d in launches:
        _cabi.capture_attention(q, k, H, H * D, H * D, d, len(sub), int(sub["n_tokens"].max()), int(sub["n_frames"].max()),
                                width, 1.0, ws, flags, partials)

bytes_alg = 4 * off + sum(4 * L * (int(t) + int(f)) * H * D for t, f in zip(Ts, Fs))
for _ in range(10):
    capture()
torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for _ in range(reps):
    for _ in range(8):  # ~0.3 ms of queued GPU work: the capture launch (144 tensor-map encodes on the host) is enqueued before
        flush.zero_()   # the GPU reaches event a, so the events time the kernel and not the host; also flushes L2
    a.record()
    capture()
    b_.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b_))
ms = float(np.median(times))
print(f"{shape} B={B} {'simt+filter' if simt else 'tcgen05'} dbg={flags:#x} partials={partials is not None}: median {ms:.3f} (min {min(times):.3f}) ms/batch ({len(launches)} launch(es)), algorithmic {bytes_alg/1e6:.1f} MB -> {bytes_alg/ms/1e6:.1f} GB/s "
      f"({bytes_alg/ms/1e6/6548.5*100:.1f}% of measured HBM peak)")
