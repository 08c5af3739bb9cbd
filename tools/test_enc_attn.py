"""Encoder attention kernel: accuracy vs an fp64 reference and speed vs torch SDPA (fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from whisper_char_alignment_b200 import _cabi

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def ref64(q, k, v, H):
    B, S, W = q.shape
    sp = lambda t: t.double().view(B, S, H, 64).transpose(1, 2)
    o = F.softmax(sp(q) @ sp(k).transpose(-1, -2) / 8.0, dim=-1) @ sp(v)
    return o.transpose(1, 2).reshape(B, S, W)


def sdpa32(q, k, v, H):
    B, S, W = q.shape
    sp = lambda t: t.view(B, S, H, 64).transpose(1, 2)
    return F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, S, W)


def check(B, S, H, gain=1.0, seed=0, fused=False):
    g = torch.Generator(device=dev).manual_seed(seed)
    if fused:
        buf = torch.randn(B, S, 3 * H * 64, device=dev, generator=g)
        q, k, v = buf[..., : H * 64], buf[..., H * 64: 2 * H * 64], buf[..., 2 * H * 64:]
    else:
        q, k, v = (torch.randn(B, S, H * 64, device=dev, generator=g) for _ in range(3))
    q = q * gain
    if fused:
        q = q.contiguous()
    out = _cabi.encoder_attention(q, k, v, H)
    torch.cuda.synchronize()
    r = ref64(q, k, v, H)
    err = (out.double() - r).abs().max().item()
    err32 = (sdpa32(q.contiguous(), k.contiguous(), v.contiguous(), H).double() - r).abs().max().item()
    scale = r.abs().max().item()
    print(f"B={B} S={S} H={H} gain={gain} fused={fused}: max|err|={err:.3e} (torch fp32 sdpa {err32:.3e}), max|ref|={scale:.3f}", flush=True)
    return err, err32


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "check"):
        bad = 0
        for args in [(1, 64, 1), (1, 128, 1), (1, 200, 3), (2, 1500, 2), (1, 1500, 16, 1.0, 1, True), (2, 448, 4, 6.0), (1, 1500, 2, 12.0)]:
            e, e32 = check(*args)
            bad += not (e < max(4 * e32, 2e-6))
        # rows with an increasing key magnitude force the lazy-rescale path
        B, S, H = 1, 1500, 1
        q = torch.ones(B, S, 64, device=dev)
        k = (torch.arange(S, device=dev).float()[None, :, None] / 8.0).expand(B, S, 64).contiguous()
        v = torch.randn(B, S, 64, device=dev)
        out = _cabi.encoder_attention(q, k, v, H)
        r = ref64(q, k, v, H)
        e = (out.double() - r).abs().max().item()
        print(f"monotone logits (rescale path): max|err|={e:.3e}")
        bad += not (e < 1e-5)
        print("FAIL" if bad else "PASS")
    if what in ("all", "time"):
        B, S, H = 16, 1500, 16
        q, k, v = (torch.randn(B, S, H * 64, device=dev) for _ in range(3))
        for name, fn in (("wca_encoder_attention", lambda: _cabi.encoder_attention(q, k, v, H)), ("torch sdpa fp32", lambda: sdpa32(q, k, v, H))):
            for _ in range(2):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(5):
                fn()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            flops = 4 * B * H * S * S * 64
            print(f"{name}: {ms:.3f} ms per call (B={B} H={H} S={S}), {flops / ms / 1e9:.1f} useful TFLOP/s")
