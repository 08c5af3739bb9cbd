"""Per-role timeline of the persistent capture kernel (CTA 0) from the WCA_CAPTURE_TRACE stamps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi

EV = ["ProdQ", "ProdK0", "ProdKLast", "SplAFree", "SplQDone", "SplK0Done", "SplKLast", "MmaAccEmpty", "MmaAReady",
      "MmaB0", "MmaIssued", "EpiAccFull", "EpiA", "EpiXMax", "EpiB", "EpiXSum", "EpiC", "SplFull", "SplLoaded", "SplStored", "SplFenced"]
shape = sys.argv[1] if len(sys.argv) > 1 else "timit"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
width = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dbg = int(sys.argv[4], 0) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
L, H, D, n_ctx = 24, 16, 64, 1500
rng = np.random.default_rng(0)
if shape == "timit":
    Ts = rng.integers(35, 56, B); Fs = rng.integers(100, 200, B)
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
d_utts = _cabi.upload_utts(recs, dev)
partials = None
if os.environ.get("WCA_PARTIALS", "0") == "1":  # head-score partials produced in the epilogue
    poff = 0
    for b in range(B):
        recs[b]["part_off"] = poff
        poff += _cabi.capture_partials_floats(L * H, int(Ts[b]), int(Fs[b]))
    partials = torch.empty(poff, device=dev)
    d_utts = _cabi.upload_utts(recs, dev)
ws = torch.empty(off, device=dev)
for _ in range(2):
    _cabi.capture_attention(q, k, H, H * D, H * D, d_utts, B, t_max, int(Fs.max()), width, 1.0, ws, 0, partials)
_cabi.capture_attention(q, k, H, H * D, H * D, d_utts, B, t_max, int(Fs.max()), width, 1.0, ws, _cabi.WCA_CAPTURE_TRACE | dbg, partials)
torch.cuda.synchronize()
tr = _cabi.capture_trace().reshape(-1, len(EV)).astype(np.float64)
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, np.nan)
np.set_printoptions(linewidth=250, suppress=True)
print("tile  " + " ".join(f"{e:>11s}" for e in EV))
for i in range(4, 16):
    print(f"{i:4d}  " + " ".join(f"{v:11.0f}" for v in tr[i]))
d = lambda a, b: np.nanmean((tr[6:36, EV.index(b)] - tr[6:36, EV.index(a)]))
print("\nmean cycles over tiles 6..35 (T=%d..%d, F=%d..%d):" % (Ts.min(), Ts.max(), Fs.min(), Fs.max()))
print(" tile period (MmaIssued[i+1]-MmaIssued[i])  :", np.nanmean(np.diff(tr[6:36, EV.index('MmaIssued')])))
print(" producer: issue(i) - MmaIssued(i-1)        :", np.nanmean(tr[7:36, EV.index('ProdQ')] - tr[6:35, EV.index('MmaIssued')]))
print(" TMA     : issue -> landed (SplFull)        :", d("ProdQ", "SplFull"))
print(" splitter: landed -> K slab lo              :", d("SplFull", "SplK0Done"))
print(" splitter: K slab lo -> Q lo + arrive       :", d("SplK0Done", "SplQDone"))
print(" mma     : AccEmpty seen -> operands ready  :", d("MmaAccEmpty", "MmaAReady"))
print(" mma     : operands ready -> 24 MMAs issued :", d("MmaAReady", "MmaIssued"))
print(" mma->epi: Issued->AccFull seen             :", d("MmaIssued", "EpiAccFull"))
print(" epilogue: sweep A                          :", d("EpiAccFull", "EpiA"))
print(" epilogue: xmax exchange                    :", d("EpiA", "EpiXMax"))
print(" epilogue: sweep B                          :", d("EpiXMax", "EpiB"))
print(" epilogue: xsum exchange                    :", d("EpiB", "EpiXSum"))
print(" epilogue: sweep C                          :", d("EpiXSum", "EpiC"))
print(" epilogue: total AccFull->C                 :", d("EpiAccFull", "EpiC"))
print(" epilogue: C(i) -> AccFull(i+2) (same WG)   :", np.nanmean(tr[8:36, EV.index('EpiAccFull')] - tr[6:34, EV.index('EpiC')]))
