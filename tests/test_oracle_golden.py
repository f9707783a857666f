"""Pins the oracle (oracle/generator_oracle.py and the plain-C restatement) to the
golden vectors produced by the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from _util import load_golden, tiny_config, tiny_state_dict
from dissc_b200 import synthetic as syn
from oracle import c_oracle
from oracle import generator_oracle as go


def test_oracle_matches_reference_tiny():
    g = load_golden("gen_tiny.npz")
    sd = tiny_state_dict(g)
    y = go.code_generator_forward(sd, tiny_config(), torch.from_numpy(g["code"]), torch.from_numpy(g["f0"]),
                                  torch.from_numpy(g["spkr"]))
    assert y.shape == g["y"].shape
    assert np.abs(y.numpy() - g["y"]).max() < 5e-6


def test_oracle_matches_reference_f0_feats():
    """`f0_feats` config: the extra `f0_stats` keyword becomes two more input channels (sr/models.py:216-221)."""
    g = load_golden("gen_f0feats.npz")
    cfg = dict(tiny_config(), model_in_dim=tiny_config()["model_in_dim"] + 2, f0_feats=True)
    y = go.code_generator_forward(tiny_state_dict(g), cfg, torch.from_numpy(g["code"]), torch.from_numpy(g["f0"]),
                                  torch.from_numpy(g["spkr"]), f0_stats=torch.from_numpy(g["f0_stats"]))
    assert y.shape == g["y"].shape
    assert np.abs(y.numpy() - g["y"]).max() < 5e-6


def test_c_oracle_matches_reference_tiny():
    g = load_golden("gen_tiny.npz")
    folded = {k: v.numpy() for k, v in go.folded_state_dict(tiny_state_dict(g)).items()}
    y = c_oracle.generator_forward(folded, tiny_config(), g["code"], g["f0"], g["spkr"])
    assert np.abs(y - g["y"]).max() < 5e-6


def test_oracle_matches_reference_vctk_T50():
    """BASELINE config 1 (B=1, 50 units).  Weights are regenerated from the seed; the
    checksum guards against RNG drift between torch versions."""
    g = load_golden("gen_vctk_T50.npz")
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    assert syn.state_dict_checksum(sd) == pytest.approx(float(g["sd_checksum"]), rel=1e-12)
    y = go.code_generator_forward(sd, syn.VCTK_CONFIG, torch.from_numpy(g["code"]), torch.from_numpy(g["f0"]),
                                  torch.from_numpy(g["spkr"]))
    assert np.abs(y.numpy() - g["y"]).max() < 2e-5  # 1-thread vs n-thread oneDNN noise is ~1.6e-5


def test_oracle_ragged_is_per_utterance():
    """H4: the golden rows are B=1 reference runs on unpadded utterances."""
    g = load_golden("gen_vctk_ragged.npz")
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    code, f0, spkr = (torch.from_numpy(g[k]) for k in ("code", "f0", "spkr"))
    b, n = 1, int(g["lengths"][1])
    y = go.code_generator_forward(sd, syn.VCTK_CONFIG, code[b:b + 1, :n], f0[b:b + 1, :, :n], spkr[b:b + 1])
    assert np.abs(y.numpy()[0, 0] - g["y"][b, 0, :320 * n]).max() < 2e-5
    assert np.all(g["y"][b, 0, 320 * n:] == 0)


def test_fold_weight_norm_matches_definition():
    gen = torch.Generator().manual_seed(0)
    v = torch.randn(6, 5, 3, generator=gen)
    gg = torch.rand(6, 1, 1, generator=gen) + 0.5
    w = go.fold_weight_norm(gg, v)
    for i in range(6):
        assert torch.allclose(w[i], gg[i] * v[i] / v[i].norm(), atol=1e-6)


def test_c_conv_primitives_match_aten():
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 41, generator=gen)
    w = torch.randn(4, 6, 7, generator=gen)
    b = torch.randn(4, generator=gen)
    for d in (1, 3, 5):
        ref = F.conv1d(x, w, b, padding=(7 * d - d) // 2, dilation=d).numpy()
        got = c_oracle.conv1d(x.numpy(), w.numpy(), b.numpy(), dilation=d, padding=(7 * d - d) // 2)
        assert np.abs(ref - got).max() < 1e-5
    for k, u in ((11, 5), (8, 4), (4, 2), (16, 8)):
        wt = torch.randn(6, 3, k, generator=gen)
        bt = torch.randn(3, generator=gen)
        ref = F.conv_transpose1d(x, wt, bt, stride=u, padding=(k - u) // 2).numpy()
        got = c_oracle.conv_transpose1d(x.numpy(), wt.numpy(), bt.numpy(), u, (k - u) // 2)
        assert ref.shape == got.shape and np.abs(ref - got).max() < 1e-5


def test_len_carryover_oracle_matches_reference():
    g = load_golden("len_carryover.npz")
    i = 0
    while f"in{i}" in g:
        got = c_oracle.len_carryover(g[f"in{i}"])
        assert np.array_equal(got, g[f"out{i}"].astype(np.int32)), i
        i += 1
    assert i == 6


def test_predictor_oracle_matches_reference():
    """oracle/predictors_oracle.py vs outputs of the real LenPredictor / PitchPredictor / PitchPredictorBase."""
    from oracle import predictors_oracle as po
    g = load_golden("predictors.npz")
    sd = syn.synthetic_len_predictor_state_dict(100, 108, seed=21)
    assert syn.state_dict_checksum({k: v for k, v in sd.items()}) == pytest.approx(float(g["len.checksum"]), rel=1e-12)
    y = po.len_predictor_forward(sd, torch.from_numpy(g["len.seq"]), torch.from_numpy(g["len.spk"]), torch.tensor(2.5),
                                 torch.tensor(1.5))
    assert np.abs(y.numpy() - g["len.y"]).max() < 1e-5
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    seq, spk = torch.from_numpy(g["pitch.seq"]), torch.from_numpy(g["pitch.spk"])
    for tag, kind in (("pitch_new", "new"), ("pitch_base", "base")):
        sd = syn.synthetic_pitch_predictor_state_dict(kind, 100, 108, seed=23)
        assert syn.state_dict_checksum(sd) == pytest.approx(float(g[f"{tag}.checksum"]), rel=1e-12)
        c, r = po.pitch_predictor_forward(sd, seq, spk, kind)
        assert np.abs(c.numpy() - g[f"{tag}.class"]).max() < 1e-5
        assert np.abs(r.numpy() - g[f"{tag}.reg"]).max() < 1e-5
        assert np.abs(po.calc_freq(c, r, spk, norm=True).numpy() - g[f"{tag}.freq_norm"]).max() < 1e-5
        assert np.abs(po.calc_freq(c, r, spk, mean, std, norm=False).numpy() - g[f"{tag}.freq_hz"]).max() < 2e-3


def test_len_carryover_numpy_oracle_matches_reference():
    from oracle import predictors_oracle as po
    g = load_golden("len_carryover.npz")
    i = 0
    while f"in{i}" in g:
        assert np.array_equal(po.len_carryover_correction(g[f"in{i}"]), g[f"out{i}"].astype(np.int64)), i
        i += 1
    assert i == 6
