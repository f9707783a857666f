"""Generates the committed golden vectors by running the REAL reference
(/root/reference, imported unmodified with import shims) in the build container.

    python tests/golden/make_golden.py

The reference ships no tests / vectors of its own, so these
outputs of the reference itself are what pins the oracle and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refimport  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)  # bit-stable fixtures

TINY_CONFIG = dict(syn.VCTK_CONFIG, upsample_initial_channel=32, embedding_dim=8, model_in_dim=17)


def run_reference(cfg, sd, code, f0, spkr):
    models, AttrDict = _refimport.sr_models()
    g = models.CodeGenerator(AttrDict(cfg))
    g.load_state_dict(sd)                    # sr/inference.py:120
    g.eval()
    g.remove_weight_norm()                   # sr/inference.py:162-163
    with torch.no_grad():
        return g(code=code, f0=f0, spkr=spkr)  # sr/inference.py:69


def gen_tiny():
    sd = syn.synthetic_generator_state_dict(TINY_CONFIG, seed=7)
    code, f0, spkr = syn.synthetic_inputs(2, 37, seed=11)
    y = run_reference(TINY_CONFIG, sd, code, f0, spkr)
    out = {"sd." + k: v.numpy() for k, v in sd.items()}
    out.update(code=code.numpy(), f0=f0.numpy(), spkr=spkr.numpy(), y=y.numpy())
    np.savez_compressed(os.path.join(HERE, "gen_tiny.npz"), **out)
    print("gen_tiny", y.shape, float(y.std()))


def gen_f0feats():
    """`f0_feats` config: CodeGenerator.forward appends the extra `f0_stats` keyword (B, 2) as two channels
    (sr/models.py:216-221; sr/inference.py:237-245 passes the target speaker's [mean, std] in Hz)."""
    cfg = dict(TINY_CONFIG, model_in_dim=TINY_CONFIG["model_in_dim"] + 2, f0_feats=True)
    sd = syn.synthetic_generator_state_dict(cfg, seed=17)
    # a trained model's weights on the two Hz-valued channels are small; keep their contribution O(1) so tanh does not saturate
    sd["conv_pre.weight_v"][:, -2:, :] *= 0.004
    code, f0, spkr = syn.synthetic_inputs(3, 29, seed=19)
    g = torch.Generator().manual_seed(23)
    f0_stats = torch.stack([150 + 60 * torch.rand(3, generator=g), 20 + 15 * torch.rand(3, generator=g)], dim=1)
    models, AttrDict = _refimport.sr_models()
    m = models.CodeGenerator(AttrDict(cfg))
    m.load_state_dict(sd)
    m.eval()
    m.remove_weight_norm()
    with torch.no_grad():
        y = m(code=code, f0=f0, spkr=spkr, f0_stats=f0_stats)
    out = {"sd." + k: v.numpy() for k, v in sd.items()}
    out.update(code=code.numpy(), f0=f0.numpy(), spkr=spkr.numpy(), f0_stats=f0_stats.numpy(), y=y.numpy())
    np.savez_compressed(os.path.join(HERE, "gen_f0feats.npz"), **out)
    print("gen_f0feats", y.shape, float(y.std()))


def gen_vctk():
    cfg = syn.VCTK_CONFIG
    sd = syn.synthetic_generator_state_dict(cfg, seed=0)
    code, f0, spkr = syn.synthetic_inputs(1, 50, seed=1234)   # BASELINE config 1
    y = run_reference(cfg, sd, code, f0, spkr)
    np.savez_compressed(os.path.join(HERE, "gen_vctk_T50.npz"), code=code.numpy(), f0=f0.numpy(), spkr=spkr.numpy(),
                        y=y.numpy(), sd_checksum=np.float64(syn.state_dict_checksum(sd)))
    print("gen_vctk_T50", y.shape, float(y.std()))
    # a B=3 ragged batch: each row is the reference run at B=1 on the unpadded utterance (H4)
    lens = [23, 9, 16]
    code, f0, spkr = syn.synthetic_inputs(3, max(lens), seed=99)
    ys = np.zeros((3, 1, 320 * max(lens)), np.float32)
    for b, n in enumerate(lens):
        yb = run_reference(cfg, sd, code[b:b + 1, :n], f0[b:b + 1, :, :n], spkr[b:b + 1])
        ys[b, 0, :320 * n] = yb[0, 0].numpy()
    np.savez_compressed(os.path.join(HERE, "gen_vctk_ragged.npz"), code=code.numpy(), f0=f0.numpy(),
                        spkr=spkr.numpy(), lengths=np.array(lens, np.int32), y=ys)
    print("gen_vctk_ragged", ys.shape)


def gen_predictors():
    """Seeded state dicts (dissc_b200.synthetic) are loaded into the REAL reference classes
    (strict=True pins the key set); only inputs/outputs/checksums are stored."""
    lp, pp = _refimport.predictors()
    g = torch.Generator().manual_seed(5)
    out = {}
    n_spk = 108
    seq = torch.randint(0, 100, (1, 61), generator=g)
    spk = torch.tensor([[17]])
    with torch.no_grad():
        sd = syn.synthetic_len_predictor_state_dict(100, n_spk, seed=21)
        m = lp.LenPredictor(100, n_spk).eval()
        m.load_state_dict(sd, strict=True)
        m.norm_mean, m.norm_std = torch.tensor(2.5), torch.tensor(1.5)
        out["len.checksum"] = np.float64(syn.state_dict_checksum(sd))
        out["len.seq"], out["len.spk"] = seq.numpy(), spk.numpy()
        out["len.y"] = m(seq, spk).numpy()
        mean, std = syn.synthetic_pitch_stats(n_spk, seed=22)
        seq2 = torch.randint(0, 100, (1, 143), generator=g)
        for tag, kind, cls in (("pitch_new", "new", pp.PitchPredictor), ("pitch_base", "base", pp.PitchPredictorBase)):
            sd = syn.synthetic_pitch_predictor_state_dict(kind, 100, n_spk, seed=23)
            m = cls(100, n_spk, id2pitch_mean=mean, id2pitch_std=std).eval()
            m.load_state_dict(sd, strict=True)
            out[f"{tag}.checksum"] = np.float64(syn.state_dict_checksum(sd))
            c, r = m(seq2, spk)
            out[f"{tag}.class"], out[f"{tag}.reg"] = c.numpy(), r.numpy()
            out[f"{tag}.freq_norm"] = m.infer_freq(seq2, spk, True).numpy()
            out[f"{tag}.freq_hz"] = m.infer_freq(seq2, spk, False).numpy()
        out["pitch.seq"], out["pitch.spk"] = seq2.numpy(), spk.numpy()
    np.savez_compressed(os.path.join(HERE, "predictors.npz"), **out)
    print("predictors", {k: getattr(v, "shape", v) for k, v in out.items()})


def gen_carryover():
    inf = _refimport.infer_module()
    g = torch.Generator().manual_seed(3)
    cases = []
    for L in (1, 2, 17, 200):
        lens = (2.5 + 1.5 * torch.randn(1, L, generator=g)).float()
        cases.append(lens)
    cases.append(torch.tensor([[0.5, 1.5, 2.5, 3.5, 0.49, 0.51, -3.0, 1.0, 1.0]]))      # half-to-even + clamp
    cases.append(torch.tensor([[1.4, 1.4, 1.4, 1.4, 1.4, 1.6, 1.6, 1.6, 1.6, 1.6]]))    # carry both ways
    out = {}
    for i, lens in enumerate(cases):
        r = inf.len_carryover_correction(lens)
        out[f"in{i}"], out[f"out{i}"] = lens.numpy(), r.numpy()
    np.savez_compressed(os.path.join(HERE, "len_carryover.npz"), **out)
    print("carryover", len(cases))


def gen_morph():
    """utils.morph_seq_len (the no-pitch-prediction path of infer.py:39-40) on seeded run-length sequences."""
    inf = _refimport.infer_module()          # `from utils import seed_everything, morph_seq_len` (infer.py:20)
    rng = np.random.RandomState(5)
    out = {}
    for i in range(6):
        n_runs = [1, 2, 5, 17, 40, 3][i]
        runs = rng.randint(1, 6, size=n_runs)
        toks = []
        prev = -1
        for _ in range(n_runs):
            t = rng.randint(0, 100)
            while t == prev:
                t = rng.randint(0, 100)
            toks.append(t)
            prev = t
        units = np.repeat(np.array(toks), runs)
        pitch = (100 + 50 * rng.rand(len(units))).astype(np.float32)
        pitch[rng.rand(len(units)) < 0.3] = 0.0
        t_lens = rng.randint(1, 8, size=n_runs)
        res = inf.morph_seq_len(units, pitch, t_lens)
        out[f"units{i}"], out[f"pitch{i}"], out[f"lens{i}"], out[f"out{i}"] = units, pitch, t_lens, np.asarray(res)
    np.savez_compressed(os.path.join(HERE, "morph_seq_len.npz"), **out)
    print("morph", 6)


if __name__ == "__main__":
    assert _refimport.available(), "reference tree not found"
    gen_tiny()
    gen_vctk()
    gen_f0feats()
    gen_predictors()
    gen_carryover()
    gen_morph()
