"""CPU suite, part 2: host-side logic of the product (tokenisation, word grouping, metrics,
argument conventions) and the C-ABI surface (library loads, exports every declared symbol,
descriptor layout agrees with the header).  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import ref_path
from whisper_char_alignment_b200 import _cabi, metrics, retokenize
from whisper_char_alignment_b200.tokenizer import get_tokenizer

TEXTS = ["Artificial intelligence is for real", "hello  big   world ", "a", "", "don't stop, it's 7 o'clock",
         "naïve café déjà vu", "x y z"]


@pytest.mark.parametrize("unit", ["char", "subword"])
@pytest.mark.parametrize("text", TEXTS)
def test_encode_and_word_split_match_the_restated_reference(text, unit, tokenizer):
    ids = retokenize.encode(text, tokenizer, unit)
    assert ids == ref_path.encode(text, tokenizer, unit)
    got = retokenize.split_tokens_on_spaces(ids + [tokenizer.eot], tokenizer, unit)
    want = ref_path.split_tokens_on_spaces(ids + [tokenizer.eot], tokenizer, unit)
    assert got == want
    assert got[0][-1] == "<|endoftext|>"
    assert sum(len(t) for t in got[1]) == len(ids) + 1
    if unit == "char":
        assert tokenizer.decode(ids) == " ".join(text.split())
    else:
        assert tokenizer.decode(ids) == text


def test_tokenizer_matches_oracle_shim_tokenizer():
    from oracle import use_shim

    use_shim()
    from whisper.tokenizer import get_tokenizer as shim_get

    a, b = get_tokenizer(True, language="English"), shim_get(True, language="English")
    assert (a.sot_sequence, a.eot, a.no_timestamps) == (b.sot_sequence, b.eot, b.no_timestamps)
    for text in TEXTS:
        assert a.encode(text) == b.encode(text)
        ids = a.encode(text) + [a.eot]
        assert a.split_to_word_tokens(ids) == b.split_to_word_tokens(ids)
        assert a.split_tokens_on_unicode(ids) == b.split_tokens_on_unicode(ids)
    en = get_tokenizer(False)
    assert len(en.sot_sequence) == 1 and en.eot == 50256


def test_unit_assertion_mirrors_reference():
    tk = get_tokenizer(True)
    with pytest.raises(AssertionError):
        retokenize.encode("a", tk, "phone")
    with pytest.raises(AssertionError):
        retokenize.split_tokens_on_spaces([1, 2], tk, "word")


def test_remove_punctuation():
    assert retokenize.remove_punctuation("Hello, world! It's 42.") == "Hello world It's fortytwo"  # the final pass strips the hyphen, as in the reference
    assert retokenize.remove_punctuation("...") == ""


def test_metrics_known_answers():
    assert metrics.eval_n1([0.5, 1.0, 2.0], [0.51, 1.5, 2.01], 0.02) == (2, 2)
    assert metrics.eval_n1([0.5], [], 0.02) == (0, 0)
    tp, fp, fn = metrics.eval_n1_strict([0.5, 1.0], [0.5, 1.0, 1.4], ["Hi,", "there"], ["hi", "THERE", "x"], 0.02)
    assert (tp, fp, fn) == (2, 1, 0)
    p, r, f1, rval, os_ = metrics.get_seg_metrics(8, 8, 10, 16)
    assert abs(p - 0.8) < 1e-6 and abs(r - 0.5) < 1e-6 and abs(f1 - 2 * 0.4 / 1.3) < 1e-6
    a = torch.tensor([[0.2, 0.9, 0.0], [0.1, 0.3, 0.0]])
    torch.testing.assert_close(metrics.coverage_penalty(a), ref_path.coverage_penalty(a))
    assert abs(metrics.coverage_penalty(a).item() - (0.5 + 1.2 + 0.5 - 1.5)) < 1e-6


# ------------------------------------------------------------------ C-ABI surface
HEADER = os.path.join(ROOT, "include", "wca_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"WCA_API\s+[\w\s\*]+?\b(wca_\w+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from whisper_char_alignment_b200 import build

    build.build()
    lib = _cabi.load()
    names = declared_symbols()
    assert len(names) >= 10
    assert sorted(_cabi.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.wca_abi_version() == _cabi.ABI_VERSION == 5
    assert lib.wca_capture_partials_floats(384, 45, 150) == 384 * 1 * 4 * 151 and lib.wca_capture_writes_partials(1500, 3, 0) == 1
    assert lib.wca_capture_writes_partials(1500, 9, 0) == 0 and lib.wca_capture_writes_partials(150, 3, _cabi.WCA_CAPTURE_FORCE_SIMT) == 0
    assert lib.wca_dtw_workspace_bytes(4, 445, 1500) == 0  # largest legal Whisper problem fits in smem
    assert lib.wca_dtw_workspace_bytes(2, 1000, 4000) > 0


def test_descriptor_layout_matches_header(tmp_path):
    prog = tmp_path / "layout.c"
    fields = [name for name, _ in _cabi.UttDesc._fields_]
    body = "\n".join(f'    printf("{f} %zu\\n", offsetof(wca_utt_t, {f}));' for f in fields)
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "wca_b200.h"\n'
        'int main(void) {\n    printf("size %zu\\n", sizeof(wca_utt_t));\n' + body + "\n    return 0;\n}\n"
    )
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    lines = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(lines["size"]) == ctypes.sizeof(_cabi.UttDesc) == _cabi.UTT_DTYPE.itemsize
    for f in fields:
        assert int(lines[f]) == getattr(_cabi.UttDesc, f).offset == _cabi.UTT_DTYPE.fields[f][1], f


def test_product_refuses_cpu_tensors_instead_of_falling_back(tokenizer):
    from whisper_char_alignment_b200 import timing

    w = torch.rand(2, 2, 8, 16)
    with pytest.raises(_cabi.WcaError):
        timing.force_align(w, [104, 105], tokenizer, "char", "mean")
    with pytest.raises(_cabi.WcaError):
        timing.dtw(torch.rand(4, 5))
    with pytest.raises(_cabi.WcaError):
        timing.median_filter_softmax(torch.rand(3, 20), 10)


def test_force_align_argument_errors_mirror_reference(tokenizer):
    from whisper_char_alignment_b200 import timing

    w = torch.rand(2, 2, 8, 16)
    with pytest.raises(AssertionError):  # timing.py:92
        timing.force_align(w, [104], tokenizer, "char", "topk", topk=-1)
    with pytest.raises(UnboundLocalError):  # timing.py:102 with an unknown aggregation
        timing.force_align(w, [104], tokenizer, "char", "median")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "whisper_char_alignment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    code = "import sys; import whisper_char_alignment_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_greedy_decode_runs_on_cpu_and_stops_at_eot(tokenizer):
    """whisper.decode stand-in (infer_ali.py:60): text tokens only, batch == single, eot never returned."""
    import torch

    from whisper_char_alignment_b200 import whisper_model

    model = whisper_model.random_init(whisper_model.ModelDimensions(80, 32, 64, 1, 1, 51865, 32, 64, 1, 1))
    g = torch.Generator().manual_seed(0)
    mel = torch.randn(2, 80, 64, generator=g)
    both = whisper_model.greedy_decode(model, mel, tokenizer, max_tokens=6)
    one = whisper_model.greedy_decode(model, mel[0], tokenizer, max_tokens=6)
    assert len(both) == 2 and both[0] == one
    for toks in both:
        assert len(toks) <= 6 and all(0 <= t < tokenizer.eot for t in toks)


# ---------------------------------------------------------------- warp-wide backtrace (csrc/dtw.cu), modelled on the host
def _walk_point_by_point(trace):
    """upstream whisper.timing.backtrace on a (N+1, M+1) trace with its border rule; returns (jump frames, path length)."""
    n, m = trace.shape[0] - 1, trace.shape[1] - 1
    jump = np.full(n, -1, dtype=np.int64)
    i, j, length = n, m, 0
    while i > 0 or j > 0:
        length += 1
        code = 1 if j == 0 else (2 if i == 0 else int(trace[i, j]))
        if code != 2 and i >= 1:
            jump[i - 1] = j - 1
        if code == 0:
            i, j = i - 1, j - 1
        elif code == 1:
            i -= 1
        else:
            j -= 1
    return jump, length


def _walk_row_by_row(trace, lanes=32):
    """What the kernel does on the jump-only path: per text row, look at `lanes` cells to the left, take the nearest one
    whose step leaves the row (a ballot + find-first-set), move up from there; path length = N + M - diagonal steps."""
    n, m = trace.shape[0] - 1, trace.shape[1] - 1
    jump = np.full(n, -1, dtype=np.int64)
    row, col, n_diag = n - 1, m - 1, 0
    while row >= 0 and col >= 0:
        window = [int(trace[row + 1, col - k + 1]) if col - k >= 0 else 2 for k in range(lanes)]
        leaves = [k for k, c in enumerate(window) if c != 2]
        if not leaves:
            col -= lanes
            continue
        k = leaves[0]
        jump[row] = col - k
        n_diag += window[k] == 0
        col -= k + (window[k] == 0)
        row -= 1
    return jump, n + m - n_diag


@pytest.mark.parametrize("seed", range(6))
def test_row_by_row_backtrace_equals_the_point_by_point_walk(seed):
    rng = np.random.default_rng(seed)
    for n, m in [(1, 1), (1, 40), (9, 1), (5, 3), (41, 150), (36, 145), (7, 400), (64, 17)]:
        # long runs of time steps (code 2) so that windows without a row change occur, and all three codes at the borders
        p2 = [0.34, 0.8, 0.97][seed % 3]
        trace = rng.choice(3, size=(n + 1, m + 1), p=[(1 - p2) / 2, (1 - p2) / 2, p2])
        want_jump, want_len = _walk_point_by_point(trace)
        got_jump, got_len = _walk_row_by_row(trace)
        np.testing.assert_array_equal(got_jump, want_jump, err_msg=f"{n}x{m}")
        assert got_len == want_len, (n, m)


# ------------------------------------------------------------------ round-2 host logic
def test_large_v3_special_tokens_follow_the_language_count():
    """large-v3 has 100 language tokens: everything after the language block moves up by one."""
    v2, v3 = get_tokenizer(True, num_languages=99), get_tokenizer(True, num_languages=100)
    assert (v2.sot_sequence, v2.no_timestamps, v2.timestamp_begin) == ((50258, 50259, 50359), 50363, 50364)
    assert (v3.sot_sequence, v3.no_timestamps, v3.timestamp_begin) == ((50258, 50259, 50360), 50364, 50365)
    assert v3.eot == v2.eot == 50257


def test_random_weights_must_be_asked_for_by_name(tmp_path):
    from whisper_char_alignment_b200 import whisper_model

    with pytest.raises(FileNotFoundError, match="random:medium"):
        whisper_model.load_model("medium")
    with pytest.raises(ValueError):
        whisper_model.load_model("random:large-v2")
    m = whisper_model.load_model("random:micro", seed=3)
    assert m.model_source == "random-init:micro:seed3:qk_gain1" and m.dims.n_audio_ctx == 1500
    path = str(tmp_path / "ckpt.pt")
    whisper_model.save_checkpoint(m, path)
    again = whisper_model.load_model(path)
    assert again.model_source.startswith("checkpoint:")
    for a, b in zip(m.state_dict().values(), again.state_dict().values()):
        assert torch.equal(a.to_dense() if a.is_sparse else a, b.to_dense() if b.is_sparse else b)


def test_empty_transcription_stays_empty_like_the_reference(tokenizer):
    """infer_ali.py:65 `len(transcription) == ''` never fires, so an empty transcript reaches force_align as
    [eot] only and yields the sentinel; the CLI must not turn it into a space token."""
    from whisper_char_alignment_b200.cli import common

    record = (None, torch.zeros(80, 3000), 32000, "", [], [], "empty")
    item = common.prepare(record, tokenizer, "subword", "cpu", None, None)
    assert item["text_tokens"] == [] and item["tokens"].tolist()[-2:] == [tokenizer.no_timestamps, tokenizer.eot]


# ------------------------------------------------------------------ audio front-end (row f3)
def _c1_pcm():
    from conftest import GOLDEN_DIR

    return np.load(os.path.join(GOLDEN_DIR, "aux_c1_pcm.npz"))["pcm16"].astype(np.float32) / 32768.0


def test_log_mel_reproduces_the_mel_of_the_c1_reference_fixture():
    """dataset.py:46-48 of the reference: log_mel_spectrogram(pad_or_trim(audio), n_mels).  The C1 fixture's mel came
    from the restated upstream front-end on sample/test.wav; the product's front-end on the same PCM gives it back."""
    from conftest import load_golden
    from whisper_char_alignment_b200 import audio

    g = load_golden("c1_base_sample")
    pcm = _c1_pcm()
    assert len(pcm) == int(g["n_samples"]) == 46592 and len(pcm) // audio.N_SAMPLES_PER_TOKEN == g["case"]["frames"]
    mel = audio.log_mel_spectrogram(audio.pad_or_trim(torch.from_numpy(pcm)), 80)
    assert mel.shape == (80, 3000)
    np.testing.assert_allclose(mel.numpy(), g["mel"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel_agrees_with_the_hf_feature_extractor(n_mels):
    """An independent implementation of the Whisper front-end (transformers' numpy WhisperFeatureExtractor: its own
    Slaney filterbank and STFT) on the same PCM: 80 mels (base/medium) and 128 mels (large-v3)."""
    transformers = pytest.importorskip("transformers")
    from whisper_char_alignment_b200 import audio

    pcm = _c1_pcm()
    fe = transformers.WhisperFeatureExtractor(feature_size=n_mels)
    want = fe(pcm, sampling_rate=16000, return_tensors="np")["input_features"][0]
    got = audio.log_mel_spectrogram(audio.pad_or_trim(torch.from_numpy(pcm)), n_mels).numpy()
    assert got.shape == want.shape == (n_mels, 3000)
    # log10 of a float32 power spectrum computed by two different FFTs: 1e-4 of the [-1, 1] range on the speech part
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
    np.testing.assert_allclose(audio.mel_filters(n_mels).numpy(), fe.mel_filters.T, rtol=0, atol=1e-6)
