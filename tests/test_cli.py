"""CLI shells: flag compatibility and host-side pieces on CPU, end-to-end runs on the GPU."""
import glob
import json
import os
import struct

import numpy as np
import pytest
import torch

from whisper_char_alignment_b200 import audio, dataset
from whisper_char_alignment_b200.cli import eval_ali, infer_ali, probe_oracle

# flags and defaults of the reference CLIs (infer_ali.py:151-173, probe_oracle.py:141-160, eval_ali.py:56-61)
INFER_FLAGS = dict(model="medium", dataset="TIMIT", scp="scp/test.wav.scp", n_mels=80, medfilt_width=7, aggr="mean",
                   topk=15, aligned_unit_type="subword", tolerance=0.02, w_colnorm=1.0, w_rownorm=1.0, w_coverage=0.0,
                   plot=False, strict=False, save_prediction=False, default_whisper_timing=False)
PROBE_FLAGS = dict(model="medium", dataset="TIMIT", scp="scp/test.wav.scp", n_mels=80, medfilt_width=7, hit_within=10,
                   aggr="mean", topk=15, aligned_unit_type="subword", tolerance=0.02, plot=False, strict=False)


def test_cli_flags_match_the_reference():
    a = vars(infer_ali.parse_args(["--output_dir", "x"]))
    for k, v in INFER_FLAGS.items():
        assert a[k] == v, k
    b = vars(probe_oracle.parse_args(["--output_dir", "x"]))
    for k, v in PROBE_FLAGS.items():
        assert b[k] == v, k
    c = vars(eval_ali.parse_args(["--pred", "p.pkl"]))
    assert c == {"pred": "p.pkl", "tolerance": 0.05}
    with pytest.raises(SystemExit):
        infer_ali.parse_args([])  # --output_dir is required, as in the reference


def write_sphere(path, pcm16):
    header = b"NIST_1A\n   1024\n" + (f"channel_count -i 1\nsample_rate -i 16000\nsample_count -i {len(pcm16)}\n"
                                      "sample_n_bytes -i 2\nsample_byte_format -s2 01\nsample_coding -s3 pcm\nend_head\n").encode()
    open(path, "wb").write(header.ljust(1024, b"\n") + pcm16.astype("<i2").tobytes())


def test_timit_reader_and_wav_reader(tmp_path):
    rng = np.random.default_rng(0)
    pcm = (rng.standard_normal(16000 * 2) * 3000).astype(np.int16)
    wav = tmp_path / "sx1.wav"
    write_sphere(wav, pcm)
    (tmp_path / "sx1.wrd").write_text("0 8000 hello\n8000 30000 world\n")
    scp = tmp_path / "test.scp"
    scp.write_text(f"dr1-sx1 {wav}\n")
    ds = dataset.TIMIT(str(scp), n_mels=80)
    samples, mel, duration, text, starts, ends, fid = ds[0]
    assert duration == 32000 and text == "hello world" and fid == "dr1-sx1"
    assert starts == [0.0, 0.5] and ends == [0.5, 1.875]
    assert mel.shape == (80, 3000) and torch.isfinite(mel).all()
    np.testing.assert_allclose(samples.numpy(), pcm / 32768.0, rtol=0, atol=1e-7)
    # RIFF/WAVE carries the same samples
    riff = tmp_path / "a.wav"
    data = pcm.astype("<i2").tobytes()
    riff.write_bytes(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                     struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + b"data" + struct.pack("<I", len(data)) + data)
    got, rate = audio.read_audio(str(riff))
    assert rate == 16000
    np.testing.assert_array_equal(got, samples.numpy())


def test_eval_ali_on_a_handmade_prediction_file(tmp_path):
    import joblib
    from collections import defaultdict

    preds = defaultdict(int)
    preds[0] = dict(starts=[0.0, 0.5], ends=[0.5, 1.0], texts=["Hello", "world"], starts_hat=np.array([0.0, 0.52]),
                    ends_hat=np.array([0.52, 1.3]), predwords=["Hello", " world", "<|endoftext|>"], fids="eval_a")
    preds[1] = 0  # a skipped utterance, as infer_ali.py stores it
    preds[2] = dict(starts=[0.0], ends=[0.4], texts=["yes"], starts_hat=np.array([0.0]), ends_hat=np.array([0.41]),
                    predwords=["yes", "<|endoftext|>"], fids="eval_b")
    path = tmp_path / "x-predictions.pkl"
    joblib.dump(preds, path)
    res = eval_ali.run_eval(eval_ali.parse_args(["--pred", str(path), "--tolerance", "0.05"]))
    # hits: "hello"@0.52 and "yes"@0.41 out of 3 reference and 3 predicted boundaries
    assert abs(res["recall"] - 2 / 3) < 1e-6 and abs(res["precision"] - 2 / 3) < 1e-6


@pytest.mark.gpu
def test_infer_eval_probe_end_to_end_on_synthetic_data(tmp_path):
    out = tmp_path / "run"
    res = infer_ali.main(["--model", "random:micro", "--dataset", "synthetic", "--scp", "timit:10", "--output_dir", str(out),
                          "--aggr", "topk", "--topk", "2", "--aligned_unit_type", "char", "--medfilt_width", "3",
                          "--tolerance", "0.5", "--save_prediction", "--batch_size", "4", "--strict"])
    assert set(res) == {"precision", "recall", "f1", "r_value"} and 0.0 <= res["f1"] <= 1.0
    dumped = json.load(open(glob.glob(str(out / "*.json"))[0]))
    assert dumped["aggr"] == "topk" and "f1" in dumped
    assert dumped["model_source"].startswith("random-init:micro") and "ground-truth" in dumped["transcript_source"]
    pkl = glob.glob(str(out / "*-predictions.pkl"))[0]
    again = eval_ali.main(["--pred", pkl, "--tolerance", "0.5"])
    assert abs(again["f1"] - res["f1"]) < 1e-9  # same counts at the same tolerance
    base = infer_ali.main(["--model", "random:micro", "--dataset", "synthetic", "--scp", "timit:4", "--output_dir", str(out),
                           "--default_whisper_timing", "--tolerance", "0.5"])
    assert 0.0 <= base["recall"] <= 1.0
    probe = probe_oracle.main(["--model", "random:micro", "--dataset", "synthetic", "--scp", "probe:3", "--output_dir", str(out),
                               "--aligned_unit_type", "char", "--medfilt_width", "3", "--tolerance", "0.5", "--hit_within", "2"])
    assert probe["utterances_probed"] == 3 and 0.0 <= probe["hit_rate"] <= 1.0
