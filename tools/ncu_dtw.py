"""One DTW launch for ncu: usage ncu_dtw.py n_prob N M"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi
n_prob, N, M = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
costs = torch.from_numpy(-np.abs(rng.standard_normal((n_prob, N, M)).astype(np.float32))).to(dev)
recs = np.zeros(n_prob, dtype=_cabi.UTT_DTYPE)
for b in range(n_prob):
    recs[b]["n_tokens"], recs[b]["n_frames"], recs[b]["row_begin"], recs[b]["row_end"] = N, M, 0, N
    recs[b]["matrix_off"], recs[b]["jump_off"] = b * N * M, b * N
d_utts = _cabi.upload_utts(recs, dev)
jumps = torch.empty(n_prob * N, dtype=torch.int32, device=dev)
nbytes = _cabi.dtw_workspace_bytes(n_prob, N, M)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None
for _ in range(3):
    _cabi.dtw_align(costs.data_ptr(), d_utts, n_prob, N, M, False, jump_frames=jumps, trace_ws=ws)
torch.cuda.synchronize()
