"""ORACLE / TEST INFRASTRUCTURE ONLY -- golden-vector generator.

Runs the reference's OWN, UNMODIFIED hot-path code (/root/reference/timing.py,
retokenize.py, metrics.py) on CPU against the restated `whisper` dependency
(oracle/whisper_shim) and writes inputs + outputs to tests/golden/*.npz.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden
The fixtures are committed; nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import use_shim  # noqa: E402

use_shim()
sys.path.insert(0, REF)

import retokenize as ref_retok  # noqa: E402  (the reference's own module)
import timing as ref_timing  # noqa: E402  (the reference's own module)
from whisper.audio import log_mel_spectrogram, pad_or_trim  # noqa: E402
from whisper.tokenizer import get_tokenizer  # noqa: E402

from oracle.synth import long_text, make_mel, make_model  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, model, seed, gain, text, unit, aggr, topk, width, qk_scale, frames, extra
    dict(name="micro_char_topk", model="micro", text="hello big world", unit="char", aggr="topk", topk=2, width=3, qk_scale=1.0, frames=97),
    dict(name="micro_char_mean", model="micro", text="hello big world", unit="char", aggr="mean", topk=-1, width=7, qk_scale=1.0, frames=97),
    dict(name="micro_sub_topk_all", model="micro", text="a quick brown fox jumps", unit="subword", aggr="topk", topk=50, width=5, qk_scale=0.5, frames=64),
    dict(name="micro_width1", model="micro", text="no filter here", unit="char", aggr="topk", topk=3, width=1, qk_scale=1.0, frames=50),
    dict(name="micro_short_frames", model="micro", text="hi yo", unit="char", aggr="topk", topk=2, width=7, qk_scale=1.0, frames=3),
    dict(name="micro_coverage", model="micro", text="coverage penalty on", unit="char", aggr="topk", topk=2, width=3, qk_scale=1.0, frames=80, w_coverage=0.7, w_colnorm=0.5, w_rownorm=2.0),
    dict(name="micro_eot_only", model="micro", text="", unit="char", aggr="topk", topk=2, width=3, qk_scale=1.0, frames=40),
    dict(name="micro_full_ctx", model="micro", text="the whole context window is used for this one", unit="char", aggr="topk", topk=3, width=7, qk_scale=1.0, frames=256),
    dict(name="mini_char_topk", model="mini", text="whisper has an internal word aligner", unit="char", aggr="topk", topk=5, width=3, qk_scale=1.0, frames=211),
    dict(name="mini_sub_mean", model="mini", text="whisper has an internal word aligner", unit="subword", aggr="mean", topk=-1, width=7, qk_scale=1.0, frames=211),
]


# Large shapes (BASELINE.json configs[2] class: T ~ 400, F up to 1500): the full (L,H,T,F) maps would be tens of MB,
# so the fixture keeps the cost matrix, the path, the times and the scores in full and the maps as a strided sample
# plus per-head digests (column sums, row arg-max, sum of squares).  The mel is re-created from its seed.
LARGE_CASES = [
    dict(name="long_char_topk", model="long", n_chars=398, text_seed=41, unit="char", aggr="topk", topk=5, width=3,
         qk_scale=1.0, frames=1500, mel_seed=501),
    dict(name="long_char_mean_w7", model="long", n_chars=190, text_seed=42, unit="char", aggr="mean", topk=-1, width=7,
         qk_scale=1.0, frames=1100, mel_seed=502),
]
from oracle.gen_golden_digests import large_digests  # noqa: E402


def run_large_case(c, model, tk):
    c = dict(c)
    c["text"] = long_text(c["n_chars"], c["text_seed"])
    mel = make_mel(model.dims.n_mels, 2 * model.dims.n_audio_ctx, 2 * c["frames"], seed=c["mel_seed"])
    text_tokens = ref_retok.encode(c["text"], tk, c["unit"])
    tokens = torch.tensor([*tk.sot_sequence, tk.no_timestamps, *text_tokens, tk.eot])
    assert len(tokens) <= 448
    w, logits = ref_timing.get_attentions(mel, tokens, model, tk, c["frames"], c["width"], c["qk_scale"])
    words, st, en, matrix, scores = ref_timing.force_align(w, text_tokens, tk, c["unit"], c["aggr"], c["topk"])
    from whisper.timing import dtw
    ti, tj = dtw(-matrix)
    out = dict(
        tokens=tokens.numpy(), text_tokens=np.array(text_tokens, dtype=np.int64),
        mel_digest=np.array([mel.double().sum().item(), mel.abs().double().sum().item()]),
        logits_digest=np.array([logits.double().sum().item(), logits.abs().double().sum().item()]),
        weights_shape=np.array(w.shape), sentinel=np.array(False),
        start_times=st, end_times=en, matrix=matrix.numpy(), words=np.array(json.dumps(words)),
        path_text=ti.astype(np.int32), path_time=tj.astype(np.int32), case=np.array(json.dumps(c)),
        **large_digests(w.numpy()),
    )
    if scores is not None:
        out.update(score_values=np.array([s[0] for s in scores], dtype=np.float64),
                   score_heads=np.array([s[1] for s in scores], dtype=np.int64))
        # every head's score, not only the selected ones (pins the ranking)
        _, all_scores = ref_timing.filter_attention(w, 10 ** 6)
        out.update(all_score_values=np.array([s[0] for s in all_scores], dtype=np.float64),
                   all_score_heads=np.array([s[1] for s in all_scores], dtype=np.int64))
    return out


def main_large():
    os.makedirs(OUT, exist_ok=True)
    tk = get_tokenizer(True, language="English")
    for c in LARGE_CASES:
        model = make_model(c["model"], 0, 4.0)
        out = run_large_case(c, model, tk)
        np.savez_compressed(os.path.join(OUT, "large_" + c["name"] + ".npz"), **out)
        print("large_" + c["name"], out["weights_shape"], out["matrix"].shape, out["start_times"][:5], len(out["path_text"]))


def sphere_pcm(path):
    """NIST SPHERE: ASCII header 'NIST_1A\\n   1024\\n', then int16 LE samples."""
    raw = open(path, "rb").read()
    assert raw[:7] == b"NIST_1A"
    hdr = int(raw[8:16].split()[0])
    return np.frombuffer(raw[hdr:], dtype="<i2").astype(np.float32) / 32768.0


def run_case(c, model, tk, mel):
    text_tokens = ref_retok.encode(c["text"], tk, c["unit"])
    tokens = torch.tensor([*tk.sot_sequence, tk.no_timestamps, *text_tokens, tk.eot])
    kw = {k: c[k] for k in ("w_colnorm", "w_rownorm", "w_coverage") if k in c}
    w, logits = ref_timing.get_attentions(mel, tokens, model, tk, c["frames"], c["width"], c["qk_scale"])
    res = ref_timing.force_align(w, text_tokens, tk, c["unit"], c["aggr"], c["topk"], **kw)
    out = dict(
        mel=mel.numpy(), tokens=tokens.numpy(), text_tokens=np.array(text_tokens, dtype=np.int64),
        weights=w.numpy(), logits_digest=np.array([logits.double().sum().item(), logits.abs().double().sum().item()]),
        sentinel=np.array(isinstance(res, list) and res[4] is None and len(res[0]) == 0 and c["text"] == ""),
    )
    if not out["sentinel"]:
        words, st, en, matrix, scores = res
        out.update(start_times=st, end_times=en, matrix=matrix.numpy(), words=np.array(json.dumps(words)))
        if scores is not None:
            out.update(score_values=np.array([s[0] for s in scores], dtype=np.float64),
                       score_heads=np.array([s[1] for s in scores], dtype=np.int64))
        # also pin the DTW path itself (what whisper.timing.dtw returns on -matrix)
        from whisper.timing import dtw
        ti, tj = dtw(-matrix)
        out.update(path_text=ti.astype(np.int64), path_time=tj.astype(np.int64))
    out["case"] = np.array(json.dumps(c))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    tk = get_tokenizer(True, language="English")
    models = {}
    for n, c in enumerate(CASES):
        c = dict(c)
        c.setdefault("seed", 0)
        c.setdefault("gain", 4.0)
        key = (c["model"], c["seed"], c["gain"])
        if key not in models:
            models[key] = make_model(c["model"], c["seed"], c["gain"])
        model = models[key]
        mel = make_mel(model.dims.n_mels, 2 * model.dims.n_audio_ctx, 2 * c["frames"], seed=100 + n)
        out = run_case(c, model, tk, mel)
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **out)
        print(c["name"], out["weights"].shape, "sentinel" if out["sentinel"] else (out["start_times"], out["end_times"]))

    # default_find_alignment (timing.py:116-186): stock-Whisper baseline on the alignment heads
    for n, c in enumerate([
        dict(name="default_micro_sub", model="micro", text="stock whisper timing path", unit="subword", frames=120, width=7, qk_scale=1.0),
        dict(name="default_mini_char", model="mini", text="alignment heads only", unit="char", frames=160, width=7, qk_scale=1.0),
    ]):
        model = make_model(c["model"], 0, 4.0)
        mel = make_mel(model.dims.n_mels, 2 * model.dims.n_audio_ctx, 2 * c["frames"], seed=300 + n)
        text_tokens = ref_retok.encode(c["text"], tk, c["unit"])
        words, st, en, weights, _ = ref_timing.default_find_alignment(model, tk, text_tokens, mel, c["frames"],
                                                                     medfilt_width=c["width"], qk_scale=c["qk_scale"])
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), mel=mel.numpy(),
                            text_tokens=np.array(text_tokens, dtype=np.int64), weights=weights.numpy(), start_times=st,
                            end_times=en, words=np.array(json.dumps(words)), case=np.array(json.dumps(c)),
                            sentinel=np.array(False))
        print(c["name"], weights.shape, st, en)

    # C1: sample/test.wav (NIST SPHERE, 46592 samples -> 145 frames), Whisper-base dims,
    # README.md:76-140 settings: char units, topk=10, medfilt_width=3.
    pcm = sphere_pcm(os.path.join(REF, "sample", "test.wav"))
    duration = len(pcm)
    mel = log_mel_spectrogram(pad_or_trim(torch.from_numpy(pcm.copy())), 80)
    c = dict(name="c1_base_sample", model="base", seed=0, gain=4.0, text="Artificial intelligence is for real",
             unit="char", aggr="topk", topk=10, width=3, qk_scale=1.0, frames=duration // 320)
    model = make_model("base", 0, 4.0)
    out = run_case(c, model, tk, mel)
    # keep the fixture small: only the speech part of the mel plus the constant pad value
    n_keep = 2 * c["frames"] + 8
    assert torch.all(mel[:, n_keep:] == mel[0, -1])
    out["mel"] = mel[:, :n_keep].numpy()
    out["mel_pad_value"] = np.array(mel[0, -1].item(), dtype=np.float32)
    out["n_samples"] = np.array(duration)
    np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **out)
    print(c["name"], out["weights"].shape, out["start_times"], out["end_times"])


def main_pcm():
    """The PCM of the C1 case (sample/test.wav, NIST SPHERE, 46592 int16 samples) so that the audio front-end of
    the product (audio.read_audio is tested on synthetic SPHERE files; audio.log_mel_spectrogram here) can be checked
    against the mel the C1 fixture was generated from, without /root/reference at test time."""
    raw = open(os.path.join(REF, "sample", "test.wav"), "rb").read()
    hdr = int(raw[8:16].split()[0])
    pcm16 = np.frombuffer(raw[hdr:], dtype="<i2").copy()
    np.savez_compressed(os.path.join(OUT, "aux_c1_pcm.npz"), pcm16=pcm16)
    print("aux_c1_pcm", pcm16.shape)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "pcm"):
        main_pcm()
    if which in ("all", "small"):
        main()
    if which in ("all", "large"):
        main_large()
