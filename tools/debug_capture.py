"""Stand-alone check of wca_capture_attention (tcgen05 vs CUDA-core vs fp64 torch) on random Q/K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as Fn
from whisper_char_alignment_b200 import _cabi

dev = torch.device("cuda:0")
torch.manual_seed(0)
L, H, D = 2, 3, 64
n_ctx = 1500

def reference(q, k, T, F, width, scale):
    s = np.float32(0.35355339059327373)
    qs = (q[:, :T] * s).double(); ks = (k[:, :F] * s).double()          # (L, T, H*D)
    qh = qs.view(L, T, H, D).permute(0, 2, 1, 3); kh = ks.view(L, F, H, D).permute(0, 2, 1, 3)
    logits = qh @ kh.transpose(-1, -2)                                     # (L,H,T,F) fp64
    x = logits.float()
    half = width // 2
    if F > half and width > 1:
        xp = Fn.pad(x.reshape(1, -1, F), (half, half), mode="reflect")[0]
        x = xp.unfold(-1, width, 1).sort()[0][..., half].reshape(L, H, T, F)
    return logits, (x * scale).softmax(-1)

def run(T, F, width, scale=1.0, simt=False, B=1):
    q = [torch.randn(B, max(T, 1), H * D, device=dev) * 2 for _ in range(L)]
    k = [torch.randn(B, n_ctx, H * D, device=dev) * 2 for _ in range(L)]
    recs = np.zeros(B, dtype=_cabi.UTT_DTYPE)
    off = 0
    for b in range(B):
        recs[b]["n_tokens"], recs[b]["n_frames"] = T, F
        recs[b]["q_row0"], recs[b]["k_row0"], recs[b]["ws_off"] = b * T, b * n_ctx, off
        off += L * H * T * F
    d_utts = _cabi.upload_utts(recs, dev)
    base = _cabi.WCA_CAPTURE_FORCE_SIMT if simt else 0
    raw = torch.full((off,), float("nan"), device=dev)
    _cabi.capture_attention(q, k, H, H * D, H * D, d_utts, B, T, F, width, scale, raw, base | _cabi.WCA_CAPTURE_RAW_LOGITS)
    ws = torch.full((off,), float("nan"), device=dev)
    _cabi.capture_attention(q, k, H, H * D, H * D, d_utts, B, T, F, width, scale, ws, base)
    torch.cuda.synchronize()
    worst = (0.0, 0.0)
    for b in range(B):
        lg, pr = reference(torch.stack([x[b] for x in q]).cpu(), torch.stack([x[b] for x in k]).cpu(), T, F, width, scale)
        n = L * H * T * F
        got_raw = raw[b * n:(b + 1) * n].view(L, H, T, F).cpu()
        got = ws[b * n:(b + 1) * n].view(L, H, T, F).cpu()
        e_raw = (got_raw.double() - lg).abs().max().item()
        rel = ((got - pr).abs() / pr.clamp_min(1e-30)).max().item()
        if not np.isfinite(e_raw) or not np.isfinite(rel):
            bad = torch.nonzero(~torch.isfinite(got_raw))
            print("   non-finite raw at", bad[:5].tolist(), "count", len(bad), " maps nan:", int((~torch.isfinite(got)).sum()))
        worst = (max(worst[0], e_raw), max(worst[1], rel))
    print(f"T={T:4d} F={F:5d} w={width} B={B} {'simt' if simt else 'tc  '}: max|logit err|={worst[0]:.3e}  max rel map err={worst[1]:.3e}", flush=True)
    return worst

cases = [(5, 3, 7), (45, 150, 3), (40, 145, 3), (45, 150, 7), (128, 464, 3), (130, 97, 5), (200, 500, 3), (64, 465, 7),
         (33, 929, 1), (448, 1500, 3), (17, 1441, 7), (45, 16, 3), (45, 17, 5)]
sel = sys.argv[1:] or ["tc"]
for T, F, w in cases:
    for mode in sel:
        run(T, F, w, simt=(mode == "simt"))
run(45, 150, 3, B=7)
run(45, 150, 3, scale=0.5)
