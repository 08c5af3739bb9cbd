"""One LibriSpeech-shaped batch through the API with per-stage error checks (debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisper_char_alignment_b200 import synthetic, timing, whisper_model, _cabi
from whisper_char_alignment_b200.tokenizer import get_tokenizer
dev = torch.device("cuda:0")
tk = get_tokenizer(True, language="English")
model = whisper_model.load_model("random:" + (sys.argv[1] if len(sys.argv) > 1 else "mini"), dev, qk_gain=4.0)
utts = synthetic.librispeech_shaped(int(sys.argv[2]) if len(sys.argv) > 2 else 12, tk, n_mels=model.dims.n_mels, seed=3)
utts.sort(key=lambda u: len(u.tokens))
print([(len(u.tokens), u.max_frames, timing._cluster_bucket(u.max_frames)) for u in utts])
mels = torch.stack([u.mel for u in utts]).to(dev)
ws, _ = timing.get_attentions_batch(mels, [u.tokens.to(dev) for u in utts], model, tk, [u.max_frames for u in utts], 3, 1.0)
torch.cuda.synchronize(); print("capture ok")
res = timing.force_align_batch(ws, [u.text_tokens for u in utts], tk, "char", "topk", 10)
torch.cuda.synchronize(); print("force_align ok", len(res))
