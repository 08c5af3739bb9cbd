"""clock64 role timeline of CTA (0,0,0) of the encoder attention kernel (debug).  usage: trace_enc_attn.py [B]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import _cabi
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = _cabi.load()
lib.wca_debug_enc_attn_buffer.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(20000, device=dev)
q, k, v = (torch.randn(B, 1500, 1024, device=dev) for _ in range(3))
for _ in range(2):
    _cabi.encoder_attention(q, k, v, 16)
torch.cuda.synchronize()
lib.wca_debug_enc_attn_buffer(dbg.data_ptr())
_cabi.encoder_attention(q, k, v, 16)
torch.cuda.synchronize()
lib.wca_debug_enc_attn_buffer(None)
names = ["tma_issue", "kv_full", "split_done", "qk_issue", "qk_issued", "s_full", "exp_done", "p_arrive", "p_ready", "pv_issued", "pv_done", "fold_done"]
st = dbg.view(torch.int32)[17000:17000 + len(names) * 32].cpu().numpy().astype(np.int64).reshape(len(names), 32)[:, :24]
t0 = st[0, 0]
rel = (st - t0) & 0xFFFFFFFF
print("block " + " ".join(f"{n:>10s}" for n in names))
for j in range(24):
    print(f"{j:5d} " + " ".join(f"{int(rel[e, j]):10d}" for e in range(len(names))))
d = np.diff(rel[names.index("s_full")])
print("s_full period: mean", d[4:].mean(), "min", d.min(), "max", d.max())
for a_, b_ in (("s_full", "exp_done"), ("exp_done", "p_arrive"), ("p_arrive", "p_ready"), ("p_ready", "pv_issued"), ("pv_issued", "pv_done"),
               ("pv_done", "fold_done"), ("qk_issue", "qk_issued"), ("qk_issued", "s_full"), ("kv_full", "split_done"), ("tma_issue", "kv_full")):
    x = rel[names.index(b_)] - rel[names.index(a_)]
    print(f"{a_:>10s} -> {b_:<10s} mean {x[4:].mean():8.0f}  min {x.min():6d} max {x.max():6d}")
