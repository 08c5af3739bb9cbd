"""Stage-by-stage dump of CTA (0,0,0) of the encoder attention kernel (debug)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisper_char_alignment_b200 import _cabi
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _cabi.load()
lib.wca_debug_enc_attn_buffer.argtypes = [ctypes.c_void_p]
dbg = torch.zeros(20000, device=dev)
lib.wca_debug_enc_attn_buffer(dbg.data_ptr())
torch.manual_seed(0)
q, k, v = (torch.randn(1, S, 64, device=dev) for _ in range(3))
out = _cabi.encoder_attention(q, k, v, 1)
torch.cuda.synchronize()
lib.wca_debug_enc_attn_buffer(None)
n = min(S, 128); nk = min(S, 64)
S0 = dbg[:8192].view(128, 64)[:n, :nk].double()
ref = (q[0, :n].double() @ k[0, :nk].double().T)
print("S0 err", (S0 - ref).abs().max().item(), "ref max", ref.abs().max().item())
if (S0 - ref).abs().max() > 1e-3:
    print("S0[0,:8]", S0[0, :8].tolist()); print("ref[0,:8]", ref[0, :8].tolist())
    # is it a permutation / transposition / partial sum?
    for name, cand in (("hi only tf32-ish", ref), ("ref.T", ref.T if n == nk else None)):
        if cand is not None:
            print(name, (S0 - cand).abs().max().item())
    # per-k-step contributions
    for ks in range(8):
        part = q[0, :n, ks*8:(ks+1)*8].double() @ k[0, :nk, ks*8:(ks+1)*8].double().T
        print("kstep", ks, "corr with S0:", torch.corrcoef(torch.stack([part.flatten(), S0.flatten()]))[0, 1].item())
    print("corr S0 vs ref:", torch.corrcoef(torch.stack([ref.flatten(), S0.flatten()]))[0, 1].item())
c = 0.125 * 1.4426950408889634
m = dbg[16512:16512 + n].double(); l = dbg[16384:16384 + n].double()
full = (q[0, :n].double() @ k[0].double().T) * c
print("m_ref vs max of block0:", (m - (ref * c).max(dim=1).values).abs().max().item())
P = torch.exp2(full - m[:, None])
print("l err", ((l - P.sum(1)) / P.sum(1)).abs().max().item())
O = dbg[8192:16384].view(128, 64)[:n].double()
Oref = P @ v[0].double()
print("O err", (O - Oref).abs().max().item(), "Oref max", Oref.abs().max().item())
if (O - Oref).abs().max() > 1e-3:
    print("O[0,:8]", O[0, :8].tolist()); print("Oref[0,:8]", Oref[0, :8].tolist())
    print("corr O vs Oref:", torch.corrcoef(torch.stack([Oref.flatten(), O.flatten()]))[0, 1].item())
    Ot = P @ v[0].double()
final = (out[0, :n].double() - (Oref / P.sum(1, keepdim=True))).abs().max().item()
print("final err", final)
