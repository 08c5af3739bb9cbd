import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi
dev = torch.device("cuda:0")
L, H, D, n_ctx = 1, 2, 64, 256
T, F = int(sys.argv[1]), int(sys.argv[2])
q = [torch.randn(1, T, H * D, device=dev) for _ in range(L)]
k = [torch.randn(1, n_ctx, H * D, device=dev) for _ in range(L)]
recs = np.zeros(1, dtype=_cabi.UTT_DTYPE)
recs[0]["n_tokens"], recs[0]["n_frames"] = T, F
d_utts = _cabi.upload_utts(recs, dev)
ws = torch.zeros(L * H * T * F, device=dev)
_cabi.capture_attention(q, k, H, H * D, H * D, d_utts, 1, T, F, 3, 1.0, ws, _cabi.WCA_CAPTURE_RAW_LOGITS)
torch.cuda.synchronize()
ref = torch.einsum("thd,fhd->htf", (q[0][0] * np.float32(0.35355339059327373)).double().view(T, H, D), (k[0][0, :F] * np.float32(0.35355339059327373)).double().view(F, H, D))
print("max err", (ws.view(H, T, F).double() - ref).abs().max().item())
