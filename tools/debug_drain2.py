import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from whisper_char_alignment_b200 import synthetic, timing, whisper_model, _cabi
from whisper_char_alignment_b200.tokenizer import get_tokenizer
dev = torch.device("cuda:0")
tk = get_tokenizer(True, language="English")
model = whisper_model.load_model("random:medium", dev, qk_gain=4.0)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
utts = [synthetic.make_utterance(rng, tk, 80, 29.9, 600, "char", f"u{i}") for i in range(n)]
print([(len(u.tokens), u.max_frames) for u in utts])
mels = torch.stack([u.mel for u in utts]).to(dev)
ws, _ = timing.get_attentions_batch(mels, [u.tokens.to(dev) for u in utts], model, tk, [u.max_frames for u in utts], 3, 1.0)
torch.cuda.synchronize(); print("capture ok")
try:
    res = timing.force_align_batch(ws, [u.text_tokens for u in utts], tk, "char", "topk", 10)
    torch.cuda.synchronize(); print("force_align ok", len(res))
except Exception as e:
    print("ERR", e)
    plan = timing._Plan(ws, 3, [10] * n, [5] * n)
    print(plan.max_tokens, plan.max_frames, plan.totals)
