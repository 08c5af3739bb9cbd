"""Where does a step go?  torch.profiler kernel table of one batched get_attentions + force_align."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _preload  # noqa: F401
import torch
from torch.profiler import ProfilerActivity, profile
from whisper_char_alignment_b200 import synthetic, timing, whisper_model
from whisper_char_alignment_b200.tokenizer import get_tokenizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "medium"
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
tk = get_tokenizer(True)
model = whisper_model.load_model(name, dev, qk_gain=4.0)
utts = synthetic.timit_shaped(B, tk, n_mels=model.dims.n_mels)
mels = torch.stack([u.mel for u in utts]).to(dev)
toks = [u.tokens.to(dev) for u in utts]

def step():
    ws, _ = timing.get_attentions_batch(mels, toks, model, tk, [u.max_frames for u in utts], 3, 1.0)
    return timing.force_align_batch(ws, [u.text_tokens for u in utts], tk, "char", "topk", 10)

for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
import time
for fn, label in [(lambda: model.encoder(mels), "encoder"), ]:
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.no_grad():
        for _ in range(3): xa = fn()
    torch.cuda.synchronize(); print(label, (time.perf_counter()-t)/3*1000, "ms")
tokpad = torch.stack([torch.cat([t, t[-1:].expand(max(len(x) for x in toks)-len(t))]) for t in toks])
torch.cuda.synchronize(); t=time.perf_counter()
with torch.no_grad():
    for _ in range(3): model.decoder(tokpad, xa)
torch.cuda.synchronize(); print("decoder", (time.perf_counter()-t)/3*1000, "ms")
print(torch.backends.cuda.preferred_blas_library(), torch.version.cuda)
import ctypes
try:
    print("cublas version", torch.cuda.current_blas_handle())
except Exception as e: print(e)
