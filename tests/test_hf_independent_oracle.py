"""CPU suite: the restated `whisper` dependency (oracle/whisper_shim) against an INDEPENDENT
implementation -- Hugging Face transformers' Whisper with eager attention (SURVEY.md section 8(c)
item 4).  The fixtures under tests/golden/ were produced by the reference's own timing.py running
on top of the shim; these tests show that the shim's arithmetic (and therefore every fixture, and
the product model that shares the parameter layout) agrees with a forward nobody here wrote.

Tolerances: the comparison goes through log(softmax) of HF's probabilities, so it carries the
rounding of one extra log/exp pair: maps within 1e-4 relative (measured: 0.4-1.5e-5) (+1e-7 absolute for entries that are
denormal-small in probability space), vocabulary logits within 2e-4 of their scale.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

transformers = pytest.importorskip("transformers")

from oracle import hf_crosscheck, ref_path  # noqa: E402

CASES = ["micro_char_topk", "micro_char_mean", "micro_sub_topk_all", "micro_full_ctx", "mini_char_topk", "mini_sub_mean"]


@pytest.fixture(scope="module")
def hf_models(oracle_models):
    cache = {}

    def get(name, seed=0, gain=4.0):
        key = (name, seed, gain)
        if key not in cache:
            cache[key] = hf_crosscheck.hf_model_from(oracle_models(name, seed, gain))
        return cache[key]

    return get


def test_every_parameter_of_the_published_layout_has_an_hf_counterpart(oracle_models):
    model = oracle_models("micro")
    hf = hf_crosscheck.hf_model_from(model)
    n_src = sum(v.numel() for k, v in model.state_dict().items() if k != "alignment_heads" and not k.endswith("mask"))
    n_dst = sum(v.numel() for k, v in hf.state_dict().items() if k != "proj_out.weight")
    assert n_src == n_dst


@pytest.mark.parametrize("name", CASES)
def test_reference_fixture_maps_agree_with_hf_eager_attention(name, hf_models):
    g = load_golden(name)
    c = g["case"]
    hf = hf_models(c["model"], c.get("seed", 0), c.get("gain", 4.0))
    probs, logits = hf_crosscheck.hf_forward(hf, torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"]))
    assert probs.shape[:3] == g["weights"].shape[:3]
    maps = hf_crosscheck.maps_from_probabilities(probs, c["frames"], c["width"], c["qk_scale"]).numpy()
    err = np.abs(maps - g["weights"]) / np.maximum(np.abs(g["weights"]), 1e-7 / 1e-4)
    print(f"{name}: max relative map difference HF vs reference-on-shim = {err.max():.3e}")
    np.testing.assert_allclose(maps, g["weights"], rtol=1e-4, atol=1e-7)
    # decoder output (timing.py:58 returns it as `logits`)
    scale = np.abs(g["logits_digest"][1]) / logits.numel()
    assert abs(logits.double().sum().item() - g["logits_digest"][0]) <= 2e-4 * scale * logits.numel()
    np.testing.assert_allclose(logits.abs().double().sum().item(), g["logits_digest"][1], rtol=2e-4)


def test_pre_softmax_logits_agree_with_hf_up_to_the_row_constant(oracle_models, hf_models, tokenizer):
    """The captured `qk` itself (timing.py:52): qk - logsumexp(qk) against log p_HF, full context."""
    g = load_golden("mini_char_topk")
    model = oracle_models("mini")
    mel, tokens = torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"])
    qk, out_logits = ref_path.capture_logits(model, mel, tokens)
    probs, logits = hf_crosscheck.hf_forward(hf_models("mini"), mel, tokens)
    logp = qk.double().log_softmax(-1)
    np.testing.assert_allclose(probs.double().log().numpy(), logp.numpy(), rtol=0, atol=5e-5)
    np.testing.assert_allclose(logits.numpy(), out_logits.numpy(), rtol=0, atol=2e-4 * float(out_logits.abs().max()))


def test_alignment_on_hf_maps_gives_the_reference_boundaries(hf_models, tokenizer):
    """End of the chain: boundaries computed from the independent forward equal the fixture's."""
    for name in ("micro_char_topk", "mini_char_topk", "mini_sub_mean"):
        g = load_golden(name)
        c = g["case"]
        probs, _ = hf_crosscheck.hf_forward(hf_models(c["model"]), torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"]))
        maps = hf_crosscheck.maps_from_probabilities(probs, c["frames"], c["width"], c["qk_scale"])
        words, st, en, _, scores = ref_path.force_align(maps, g["text_tokens"].tolist(), tokenizer, c["unit"], c["aggr"], c["topk"])
        assert words == g["words"]
        np.testing.assert_array_equal(st, g["start_times"])
        np.testing.assert_array_equal(en, g["end_times"])
        if scores is not None:
            assert [list(s[1]) for s in scores] == g["score_heads"].tolist()


def test_product_greedy_decode_agrees_with_hf_greedy(oracle_models, hf_models, tokenizer):
    """Row f4: the transcript step (infer_ali.py:60).  The product's greedy_decode on CPU against a plain
    greedy loop over the HF model with the same weights."""
    from dataclasses import asdict

    from whisper_char_alignment_b200 import whisper_model

    om = oracle_models("micro")
    pm = whisper_model.Whisper(whisper_model.ModelDimensions(**asdict(om.dims))).eval()
    pm.load_state_dict(om.state_dict())
    g = load_golden("micro_char_topk")
    mel = torch.from_numpy(g["mel"])
    got = whisper_model.greedy_decode(pm, mel, tokenizer, max_tokens=12)
    want = hf_crosscheck.hf_greedy(hf_models("micro"), mel, [*tokenizer.sot_sequence, tokenizer.no_timestamps],
                                   tokenizer.eot, 12)
    assert got == want


def test_large_reference_fixture_agrees_with_hf(hf_models):
    """Same statement at the LibriSpeech-class shape (T = 195, F = 1100 of 1500): sample + digests of the maps the
    reference produced on the shim, against the independent forward."""
    from conftest import check_large_maps, load_large

    g = load_large("large_long_char_mean_w7")
    c = g["case"]
    probs, _ = hf_crosscheck.hf_forward(hf_models(c["model"]), torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"]))
    maps = hf_crosscheck.maps_from_probabilities(probs, c["frames"], c["width"], c["qk_scale"]).numpy()
    check_large_maps(maps, g, 1e-4, "HF vs reference-on-shim, large_long_char_mean_w7")
