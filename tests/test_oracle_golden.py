"""CPU suite, part 1: the oracle restatement against the fixtures that the reference's own
timing.py produced (oracle/gen_golden.py).  Float tensors are compared with a tight
tolerance because the fixtures may have been generated on another CPU model; everything
downstream of a stored cost matrix is integer work and must match exactly."""
import numpy as np
import pytest
import torch

from conftest import default_timing_names, golden_names, load_golden
from oracle import dtw as odtw
from oracle import ref_path

NAMES = golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_restatement_reproduces_reference_fixture(name, oracle_models, tokenizer):
    g = load_golden(name)
    c = g["case"]
    model = oracle_models(c["model"], c.get("seed", 0), c.get("gain", 4.0))
    mel, tokens = torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"])
    w, logits = ref_path.get_attentions(mel, tokens, model, tokenizer, c["frames"], c["width"], c["qk_scale"])
    np.testing.assert_allclose(w.numpy(), g["weights"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(logits.double().sum().item(), g["logits_digest"][0], rtol=1e-5)
    kw = {k: c[k] for k in ("w_colnorm", "w_rownorm", "w_coverage") if k in c}
    text_tokens = g["text_tokens"].tolist()
    # feed the STORED maps so that everything below is decided by identical inputs
    res = ref_path.force_align(torch.from_numpy(g["weights"]), text_tokens, tokenizer, c["unit"], c["aggr"], c["topk"], **kw)
    if g["sentinel"]:
        assert isinstance(res, list) and res == [[], [], [], [], None]
        return
    words, st, en, matrix, scores = res
    assert words == g["words"]
    np.testing.assert_allclose(matrix.numpy(), g["matrix"], rtol=1e-6, atol=1e-9)
    if scores is not None:
        assert [list(s[1]) for s in scores] == g["score_heads"].tolist()
        np.testing.assert_allclose([s[0] for s in scores], g["score_values"], rtol=1e-6)
    np.testing.assert_array_equal(st, g["start_times"])
    np.testing.assert_array_equal(en, g["end_times"])


@pytest.mark.parametrize("name", [n for n in NAMES if "eot_only" not in n])
def test_c_dtw_matches_fixture_path_bit_exact(name):
    g = load_golden(name)
    ti, tj = odtw.dtw_path(-g["matrix"])
    np.testing.assert_array_equal(ti, g["path_text"])
    np.testing.assert_array_equal(tj, g["path_time"])
    jumps = odtw.jump_frames(ti, tj)
    assert len(jumps) == g["matrix"].shape[0]
    # boundaries are the jump frames gathered at the word boundaries, / 50 (timing.py:108-113)
    assert set(np.round(g["end_times"] * 50).astype(int)).issubset(set(jumps.tolist()))


def test_c_dtw_matches_numba_restatement_on_ties_and_planted_paths():
    from oracle import use_shim

    use_shim()
    from whisper.timing import dtw as numba_dtw

    rng = np.random.default_rng(7)
    for trial in range(60):
        n, m = int(rng.integers(1, 70)), int(rng.integers(1, 110))
        x = rng.standard_normal((n, m)).astype(np.float32)
        if trial % 2:
            x = np.round(x * 2) / 2  # heavy ties
        a = numba_dtw(torch.from_numpy(x))
        b = odtw.dtw_path(x)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    # planted monotone path (the idea of upstream's own DTW test): random costs in [0,1),
    # the planted cells lowered by one, right-angle corners cut into diagonal steps so the
    # planted path is the unique optimum even under the recurrence's tie rule
    for n, m in [(10, 20), (32, 16), (123, 1500), (234, 189)]:
        want = plant_path(rng, n, m)
        x = rng.random((n, m)).astype(np.float32)
        x[want[0], want[1]] -= 1
        ti, tj = odtw.dtw_path(x)
        np.testing.assert_array_equal(ti, want[0])
        np.testing.assert_array_equal(tj, want[1])


def plant_path(rng, n, m):
    moves = np.concatenate([np.zeros(n - 1, int), np.ones(m - 1, int)])
    rng.shuffle(moves)
    i = j = k = 0
    pts = [(0, 0)]
    while k < len(moves):
        if k + 1 < len(moves) and moves[k] != moves[k + 1]:
            i, j, k = i + 1, j + 1, k + 2
        elif moves[k] == 0:
            i, k = i + 1, k + 1
        else:
            j, k = j + 1, k + 1
        pts.append((i, j))
    return np.array(pts).T


def test_median_restatement_matches_scipy():
    from scipy.ndimage import median_filter as sp_median

    g = torch.Generator().manual_seed(3)
    for shape, width in [((10,), 3), ((1, 15), 5), ((4, 5, 345), 7), ((2, 3, 24, 51), 13), ((3, 2), 7)]:
        x = torch.randn(*shape, generator=g)
        got = ref_path.median_along_frames(x, width).numpy()
        if shape[-1] <= width // 2:
            np.testing.assert_array_equal(got, x.numpy())
            continue
        size = [1] * (x.ndim - 1) + [width]
        want = sp_median(x.numpy(), size=size, mode="mirror")  # scipy 'mirror' == torch 'reflect'
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name", default_timing_names())
def test_default_find_alignment_restatement(name, oracle_models, tokenizer):
    g = load_golden(name)
    c = g["case"]
    model = oracle_models(c["model"])
    words, st, en, weights, none = ref_path.default_find_alignment(
        model, tokenizer, g["text_tokens"].tolist(), torch.from_numpy(g["mel"]), c["frames"],
        medfilt_width=c["width"], qk_scale=c["qk_scale"])
    assert none is None and words == g["words"]
    np.testing.assert_allclose(weights.numpy(), g["weights"], rtol=2e-4, atol=2e-5)
    np.testing.assert_array_equal(st, g["start_times"])
    np.testing.assert_array_equal(en, g["end_times"])


# ------------------------------------------------------------------ large, reference-generated shapes
from conftest import check_large_maps, large_names, load_large  # noqa: E402


@pytest.mark.parametrize("name", large_names())
def test_restatement_reproduces_large_reference_fixture(name, oracle_models, tokenizer):
    """T ~ 200-400, F 1100-1500 (BASELINE.json configs[2] class) through the reference's own timing.py: the port
    regenerates the maps (checked on the stored sample + digests), and from them the same matrix, scores, path
    and boundaries."""
    g = load_large(name)
    c = g["case"]
    model = oracle_models(c["model"])
    w, logits = ref_path.get_attentions(torch.from_numpy(g["mel"]), torch.from_numpy(g["tokens"]), model, tokenizer,
                                        c["frames"], c["width"], c["qk_scale"])
    check_large_maps(w.numpy(), g, 2e-5, name)
    np.testing.assert_allclose(logits.double().sum().item(), g["logits_digest"][0], rtol=1e-5)
    words, st, en, matrix, scores = ref_path.force_align(w, g["text_tokens"].tolist(), tokenizer, c["unit"], c["aggr"], c["topk"])
    assert words == g["words"]
    np.testing.assert_allclose(matrix.numpy(), g["matrix"], rtol=2e-5, atol=1e-9)
    if scores is not None:
        assert [list(s[1]) for s in scores] == g["score_heads"].tolist()
        np.testing.assert_allclose([s[0] for s in scores], g["score_values"], rtol=1e-5)
    # integer work on the STORED matrix: exact
    ti, tj = odtw.dtw_path(-g["matrix"])
    np.testing.assert_array_equal(ti, g["path_text"])
    np.testing.assert_array_equal(tj, g["path_time"])
    _, wt = ref_path.split_tokens_on_spaces(g["text_tokens"].tolist() + [tokenizer.eot], tokenizer, c["unit"])
    st2, en2, _ = ref_path.boundaries_from_path(ti, tj, wt)
    np.testing.assert_array_equal(st2, g["start_times"])
    np.testing.assert_array_equal(en2, g["end_times"])
