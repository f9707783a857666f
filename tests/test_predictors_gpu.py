"""GPU parity of the prosody predictors and infer.py glue (through the C ABI) against the oracle and the
reference-generated goldens (tests/golden/predictors.npz, len_carryover.npz)."""
import json

import numpy as np
import pytest
import torch

from _util import load_golden
from dissc_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _models(dev):
    from dissc_b200.predictors import LenPredictor, PitchPredictor, PitchPredictorBase
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    lm = LenPredictor(100, 108).to(dev)
    lm.load_state_dict(syn.synthetic_len_predictor_state_dict(100, 108, seed=21), strict=True)
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(1.5)
    pn = PitchPredictor(100, 108, id2pitch_mean=mean.to(dev), id2pitch_std=std.to(dev)).to(dev)
    pn.load_state_dict(syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=23), strict=True)
    pb = PitchPredictorBase(100, 108, id2pitch_mean=mean.to(dev), id2pitch_std=std.to(dev)).to(dev)
    pb.load_state_dict(syn.synthetic_pitch_predictor_state_dict("base", 100, 108, seed=23), strict=True)
    return lm.eval(), pn.eval(), pb.eval(), mean, std


def test_len_predictor_golden(cuda_device):
    g = load_golden("predictors.npz")
    lm, _, _, _, _ = _models(cuda_device)
    y = lm(torch.from_numpy(g["len.seq"]).to(cuda_device), torch.from_numpy(g["len.spk"]).to(cuda_device))
    assert y.shape == g["len.y"].shape
    assert np.abs(y.cpu().numpy() - g["len.y"]).max() < 1e-4


@pytest.mark.parametrize("tag", ["pitch_new", "pitch_base"])
def test_pitch_predictor_golden(cuda_device, tag):
    g = load_golden("predictors.npz")
    _, pn, pb, _, _ = _models(cuda_device)
    m = pn if tag == "pitch_new" else pb
    seq, spk = torch.from_numpy(g["pitch.seq"]).to(cuda_device), torch.from_numpy(g["pitch.spk"]).to(cuda_device)
    c, r = m(seq, spk)
    assert np.abs(c.cpu().numpy() - g[f"{tag}.class"]).max() < 1e-4
    assert np.abs(r.cpu().numpy() - g[f"{tag}.reg"]).max() < 1e-4
    fn = m.infer_freq(seq, spk, True).cpu().numpy()
    fh = m.infer_freq(seq, spk, False).cpu().numpy()
    # voicing decisions (class logit > 0) must agree wherever the golden logit is not within 1e-4 of zero
    sure = np.abs(g[f"{tag}.class"]) > 1e-4
    assert np.array_equal((fn != 0)[sure], (g[f"{tag}.freq_norm"] != 0)[sure])
    assert np.abs(fn - g[f"{tag}.freq_norm"])[sure].max() < 1e-4
    assert np.abs(fh - g[f"{tag}.freq_hz"])[sure].max() < 5e-3   # Hz scale (std ~30): 1e-4 relative


def test_len_carryover_golden_bit_exact(cuda_device):
    from dissc_b200.infer import len_carryover_correction
    g = load_golden("len_carryover.npz")
    i = 0
    while f"in{i}" in g:
        got = len_carryover_correction(torch.from_numpy(g[f"in{i}"]).to(cuda_device))
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), g[f"out{i}"]), i
        i += 1
    assert i == 6


def test_len_carryover_random_vs_oracle(cuda_device):
    from dissc_b200.infer import len_carryover_correction
    from oracle import predictors_oracle as po
    gen = torch.Generator().manual_seed(0)
    B, L = 37, 211
    lens = (2.5 + 1.5 * torch.randn(B, L, generator=gen)).float()
    lengths = torch.randint(1, L + 1, (B,), generator=gen, dtype=torch.int32)
    got, totals = len_carryover_correction(lens.to(cuda_device), lengths.to(cuda_device), return_totals=True)
    got, totals = got.cpu().numpy(), totals.cpu().numpy()
    for b in range(B):
        n = int(lengths[b])
        want = po.len_carryover_correction(lens[b:b + 1, :n].numpy())
        assert np.array_equal(got[b, :n], want), b
        assert np.all(got[b, n:] == 0) and totals[b] == want.sum()


def test_dedup_and_repeat_interleave(cuda_device):
    from dissc_b200.infer import dedup_units, repeat_interleave
    from oracle import predictors_oracle as po
    gen = torch.Generator().manual_seed(1)
    B, L, pad = 9, 157, 100
    seq = torch.full((B, L), pad, dtype=torch.int64)
    lens = [157, 1, 2, 50, 99, 100, 3, 77, 156]
    for b, n in enumerate(lens):
        runs = torch.randint(0, 100, (n,), generator=gen)
        seq[b, :n] = torch.repeat_interleave(runs, torch.randint(1, 4, (n,), generator=gen))[:n]
    seq[5, 10] = pad   # a pad token in the middle is dropped (seqs != n_tokens), merging nothing across it
    dd, counts, dd_len = dedup_units(seq.to(cuda_device), pad)
    for b in range(B):
        s = [int(u) for u in seq[b].tolist() if u != pad]
        vals, cnt = po.dedup_seq(s)
        n = int(dd_len[b])
        assert dd[b, :n].tolist() == vals and counts[b, :n].tolist() == cnt
        assert torch.all(dd[b, n:] == pad)
    Lout = int(counts.sum(1).max())
    out, out_len = repeat_interleave(dd, counts, dd_len, pad, Lout)
    for b in range(B):
        s = [int(u) for u in seq[b].tolist() if u != pad]
        assert out[b, :int(out_len[b])].tolist() == s
        assert torch.all(out[b, int(out_len[b]):] == pad)


def test_convert_batch_matches_per_utterance_oracle(cuda_device, tmp_path):
    """A padded batch through dedup -> LenPredictor -> carry-over -> repeat_interleave -> PitchPredictor equals the
    oracle's B=1 ``_infer_sample`` per utterance (units exactly; f0 within tolerance)."""
    from dissc_b200.infer import convert_batch, infer_sample
    from oracle import predictors_oracle as po
    lm, pn, pb, mean, std = _models(cuda_device)
    len_sd = syn.synthetic_len_predictor_state_dict(100, 108, seed=21)
    gen = torch.Generator().manual_seed(5)
    B, pad = 6, 100
    lens = [120, 7, 64, 1, 33, 90]
    L = max(lens)
    seqs = torch.full((B, L), pad, dtype=torch.int64)
    for b, n in enumerate(lens):
        runs = torch.randint(0, 100, (n,), generator=gen)
        seqs[b, :n] = torch.repeat_interleave(runs, torch.randint(1, 5, (n,), generator=gen))[:n]
    spk = torch.randint(0, 108, (B, 1), generator=gen)
    for kind, pm in (("new", pn), ("base", pb)):
        psd = syn.synthetic_pitch_predictor_state_dict(kind, 100, 108, seed=23)
        out_seq, f0, out_len = convert_batch(seqs.to(cuda_device), spk.to(cuda_device), pad, lm, pm, norm_pitch=True)
        for b in range(B):
            want_u, want_f = po.infer_sample(seqs[b].tolist(), int(spk[b]), pad, len_sd, (torch.tensor(2.5), torch.tensor(1.5)),
                                             psd, kind, mean, std, norm_pitch=True)
            n = int(out_len[b])
            assert out_seq[b, :n].tolist() == want_u.tolist(), (kind, b)
            got_f = f0[b, :n].cpu().numpy()
            ok = np.isclose(got_f, want_f.numpy(), atol=1e-4) | (np.abs(got_f - want_f.numpy()) < 1e-4)
            assert ok.mean() > 0.99, (kind, b)   # a voicing logit within 1e-5 of 0 may flip
    # B=1 signature-compatible path writes the reference's JSON line
    p = tmp_path / "out.txt"
    o = infer_sample(seqs[2].to(cuda_device), None, spk[2].to(cuda_device), "p225_001.wav", str(p), lm, pn, True, pad)
    line = json.loads(p.read_text().strip())
    assert line == o and set(line) == {"units", "f0", "audio"}


def test_pitch_new_rejects_more_than_850_units(cuda_device):
    from dissc_b200 import _lib
    _, pn, _, _, _ = _models(cuda_device)
    seq = torch.zeros((1, 851), dtype=torch.int64, device=cuda_device)
    with pytest.raises(_lib.DisscError):
        pn(seq, torch.zeros((1, 1), dtype=torch.int64, device=cuda_device))


def test_out_of_range_ids_raise(cuda_device):
    """nn.Embedding raises IndexError on a token / speaker id outside its table (model/len_predictor.py:15-16,
    model/pitch_predictor.py:51-52); here the gather stays in bounds and the error surfaces at check_indices()."""
    from dissc_b200 import synthetic as syn
    from dissc_b200.predictors import LenPredictor, PitchPredictor
    lm = LenPredictor(100, 108).to(cuda_device)
    lm.load_state_dict(syn.synthetic_len_predictor_state_dict(100, 108, seed=1))
    mean, std = syn.synthetic_pitch_stats(108, seed=2)
    pm = PitchPredictor(100, 108, id2pitch_mean=mean.to(cuda_device), id2pitch_std=std.to(cuda_device)).to(cuda_device)
    pm.load_state_dict(syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=3))
    seq = torch.randint(0, 100, (2, 12), device=cuda_device)
    spk = torch.tensor([[3], [7]], device=cuda_device)
    lm(seq, spk)
    lm.check_indices()
    bad = seq.clone()
    bad[1, 4] = 101                                  # table has n_tokens + 1 = 101 rows
    out = lm(bad, spk)
    assert torch.isfinite(out).all()
    with pytest.raises(IndexError, match="unit"):
        lm.check_indices()
    lm(seq, torch.tensor([[3], [108]], device=cuda_device))   # LenPredictor's speaker table has exactly n_speakers rows
    with pytest.raises(IndexError, match="speaker"):
        lm.check_indices()
    pm(seq, spk)
    pm.check_indices()
    pm(seq, torch.tensor([[-1], [7]], device=cuda_device))
    with pytest.raises(IndexError, match="speaker"):
        pm.check_indices()
    # calc_freq with a speaker id outside the statistics tables: that row is NaN, the others are untouched
    cls, reg = pm(seq, spk)
    pm.check_indices()
    good = pm.calc_freq(cls, reg, spk, norm=False)
    odd = pm.calc_freq(cls, reg, torch.tensor([[3], [108]], device=cuda_device), norm=False)
    assert torch.equal(odd[0], good[0]) and torch.isnan(odd[1]).all()
