"""End-to-end CLI runs on the GPU: dissc_b200.inference (sr/inference.py surface) and dissc_b200.infer (infer.py
surface) against the oracle, on a synthetic checkpoint directory laid out like the reference's."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from _util import tiny_config
from dissc_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_vocoder_cli_matches_oracle(cuda_device, tmp_path):
    from scipy.io import wavfile
    from dissc_b200 import inference as inf
    from oracle import generator_oracle as go
    ck = tmp_path / "ckpt"
    data = tmp_path / "data"
    ck.mkdir()
    data.mkdir()
    spk = ["p225", "p226", "p227"]
    stats = {"p225": {"mean": 110.0, "std": 20.0}, "p226": {"mean": 200.0, "std": 35.0},
             "p227": {"mean": 150.0, "std": 25.0}}
    pickle.dump(spk, open(data / "id_to_spkr.pkl", "wb"))
    pickle.dump(stats, open(data / "f0_stats.pkl", "wb"))
    cfg = dict(tiny_config(), f0_normalize=True, f0_stats=str(data / "f0_stats.pkl"),
               input_training_file=str(data / "train.txt"), test_base_path=str(data / "wav"))
    json.dump(cfg, open(ck / "config.json", "w"))
    sd = syn.synthetic_generator_state_dict(cfg, seed=3)
    torch.save({"generator": sd}, ck / "g_00000005")
    torch.save({"generator": {k: torch.zeros_like(v) for k, v in sd.items()}}, ck / "g_00000001")  # older: must be ignored
    gen = torch.Generator().manual_seed(0)
    lines, utts = [], []
    for i, (s, n) in enumerate([("p226", 41), ("p225", 9), ("p227", 23), ("p226", 40)]):
        units = torch.randint(0, 100, (n,), generator=gen).tolist()
        f0 = (stats[s]["mean"] + stats[s]["std"] * torch.randn(n, generator=gen)).tolist()
        for j in range(0, n, 3):
            f0[j] = 0.0
        name = f"{s}_{i:03d}_mic2.wav"
        utts.append((s, units, f0, name))
        lines.append(str({"units": units, "f0": f0, "audio": "x/" + name}))
    (data / "val.txt").write_text("\n".join(lines) + "\n")
    out = tmp_path / "gen"
    inf.main(["--input_code_file", str(data / "val.txt"), "--checkpoint_file", str(ck), "--output_dir", str(out),
              "--vc", "--target-speakers", "p227", "--batch", "3"])
    for s, units, f0, name in utts:
        f0n = inf.normalize_f0(np.asarray(f0), stats[s]["mean"], stats[s]["std"])
        for tgt, suffix in ((s, "_gen.wav"), ("p227", "_2_gen.wav")):
            y = go.code_generator_forward(sd, cfg, torch.tensor([units]), torch.from_numpy(f0n).view(1, 1, -1),
                                          torch.tensor([[spk.index(tgt)]]))
            want = inf.peak_normalize(go.generate_int16(y))
            rate, got = wavfile.read(out / (name[:-4] + suffix))
            assert rate == 16000 and got.dtype == np.float32 and got.shape == want.shape
            # int16 quantisation of a 1e-5-accurate waveform, then peak normalisation: allow 2 LSB
            assert np.abs(got - want).max() < 2.5 / max(1.0, np.abs(go.generate_int16(y)).max())


def test_prosody_cli_matches_oracle(cuda_device, tmp_path):
    from dissc_b200 import infer as pinf
    from oracle import predictors_oracle as po
    data = tmp_path / "data"
    lenm = tmp_path / "len"
    f0m = tmp_path / "pitch"
    outd = tmp_path / "pred"
    for d in (data, lenm, f0m):
        d.mkdir()
    spk = [f"p{225 + i}" for i in range(108)]
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    pickle.dump(spk, open(data / "id_to_spkr.pkl", "wb"))
    pickle.dump({s: {"mean": float(mean[i]), "std": float(std[i])} for i, s in enumerate(spk)},
                open(data / "f0_stats.pkl", "wb"))
    len_sd = syn.synthetic_len_predictor_state_dict(100, 108, seed=21)
    psd = syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=23)
    torch.save(len_sd, lenm / "best_model.pth")
    torch.save((torch.tensor(2.5), torch.tensor(1.5)), lenm / "len_norm_stats.pth")
    torch.save(psd, f0m / "best_model.pth")
    gen = torch.Generator().manual_seed(2)
    rows = []
    for i, n in enumerate([30, 77, 5]):
        runs = torch.randint(0, 100, (n,), generator=gen)
        units = torch.repeat_interleave(runs, torch.randint(1, 4, (n,), generator=gen)).tolist()
        rows.append({"units": units, "f0": [0.0] * len(units), "audio": f"{spk[i]}_{i:03d}.wav"})
    (data / "val.txt").write_text("\n".join(str(r) for r in rows) + "\n")
    pinf.main(["--input_path", str(data / "val.txt"), "--out_path", str(outd), "--pred_len", "--pred_pitch",
               "--len_model", str(lenm) + "/", "--f0_model", str(f0m) + "/", "--f0_path", str(data / "f0_stats.pkl"),
               "--vc", "--target_speakers", "p300", "--device", "cuda:0"])
    for fname, tgt in (("val.txt", None), ("p300_val.txt", "p300")):
        got = [json.loads(l) for l in open(outd / fname)]
        assert len(got) == len(rows)
        for r, g in zip(rows, got):
            sid = spk.index(tgt if tgt else r["audio"].split("_")[0])
            wu, wf = po.infer_sample(r["units"], sid, 100, len_sd, (torch.tensor(2.5), torch.tensor(1.5)), psd, "new",
                                     mean, std, norm_pitch=True)
            assert g["audio"] == r["audio"] and g["units"] == wu.tolist()
            assert np.abs(np.asarray(g["f0"]) - wf.numpy()).max() < 1e-3


def test_vocoder_cli_f0_stats_sample_df_and_gt(cuda_device, tmp_path):
    """--f0-stats re-scaling (sr/inference.py:220-235), --sample_df target selection (:97-100,:214-216) and the
    <stem>_gt.wav copy (:253-256) against the oracle fed with the reference's own arithmetic."""
    import pandas as pd
    from scipy.io import wavfile
    from dissc_b200 import inference as inf
    from oracle import generator_oracle as go
    ck, data, wav = tmp_path / "ckpt", tmp_path / "data", tmp_path / "data" / "wav"
    ck.mkdir()
    wav.mkdir(parents=True)
    spk = ["p225", "p226", "p227"]
    pickle.dump(spk, open(data / "id_to_spkr.pkl", "wb"))
    cfg = dict(tiny_config(), f0_normalize=False, input_training_file=str(data / "train.txt"), test_base_path=str(wav))
    json.dump(cfg, open(ck / "config.json", "w"))
    sd = syn.synthetic_generator_state_dict(cfg, seed=4)
    torch.save({"generator": sd}, ck / "g_00000001")
    tstats = {1: {"f0_mean": 0.3, "f0_std": 1.2}, "f0_mean": -0.1, "f0_std": 0.8}     # speaker 2 falls back to the global entry
    torch.save(tstats, tmp_path / "tgt_f0.pt")
    gen = torch.Generator().manual_seed(1)
    lines, utts = [], []
    for i, (s, n) in enumerate([("p225", 30), ("p226", 12)]):
        units = torch.randint(0, 100, (n,), generator=gen).tolist()
        f0 = torch.randn(n, generator=gen).tolist()
        for j in range(0, n, 4):
            f0[j] = 0.0
        name = f"{s}_{i:03d}_mic2.wav"
        utts.append((s, units, f0, name))
        lines.append(str({"units": units, "f0": f0, "audio": name}))
        wavfile.write(str(wav / name), 16000, (3000 * torch.randn(n * 320, generator=gen)).to(torch.int16).numpy())
    (data / "val.txt").write_text("\n".join(lines) + "\n")

    def want_vc(units, f0, k):
        f = torch.tensor(f0).view(1, 1, -1)
        m, s_ = inf.target_f0_stats(tstats, k)
        f = torch.from_numpy(inf.rescale_f0(f.view(-1).numpy(), m, s_)).view(1, 1, -1)
        y = go.code_generator_forward(sd, cfg, torch.tensor([units]), f, torch.tensor([[k]]))
        return go.generate_int16(y)

    # 1. --vc with --f0-stats for two targets + resynthesis + gt copies
    out = tmp_path / "gen"
    inf.main(["--input_code_file", str(data / "val.txt"), "--checkpoint_file", str(ck), "--output_dir", str(out), "--vc",
              "--target-speakers", "p226", "--f0-stats", str(tmp_path / "tgt_f0.pt")])
    for s, units, f0, name in utts:
        raw = want_vc(units, f0, 1)
        rate, got = wavfile.read(out / (name[:-4] + "_1_gen.wav"))
        assert np.abs(got - inf.peak_normalize(raw)).max() < 2.5 / max(1.0, np.abs(raw).max())
        assert (out / (name[:-4] + "_gen.wav")).exists()
        rate, gt = wavfile.read(out / (name[:-4] + "_gt.wav"))
        _, src = wavfile.read(wav / name)
        assert rate == 16000 and gt.dtype == np.float32 and np.allclose(gt, src / np.abs(src).max(), atol=1e-6)
    # 2. --sample_df: only the listed (sample, target) conversions, no resynthesis, no gt
    df = pd.DataFrame({"syn_sample": ["p225_000", "p226_001", "p226_001"], "syn_trgt": ["p227", "p225", "p227"]})
    df.to_csv(tmp_path / "samples.csv")
    out2 = tmp_path / "gen2"
    inf.main(["--input_code_file", str(data / "val.txt"), "--checkpoint_file", str(ck), "--output_dir", str(out2), "--vc",
              "--sample_df", str(tmp_path / "samples.csv")])
    assert sorted(os.listdir(out2)) == ["p225_000_mic2_2_gen.wav", "p226_001_mic2_0_gen.wav", "p226_001_mic2_2_gen.wav"]


def test_prosody_cli_morph_mode_and_sample_df(cuda_device, tmp_path):
    """--pred_len without --pred_pitch: the original contour is morphed per run (utils.morph_seq_len, infer.py:39-40),
    normalised with the SOURCE speaker's statistics; --sample_df restricts the conversions per sample (:118-119)."""
    import pandas as pd
    from dissc_b200 import infer as pinf
    from dissc_b200.predictors import LenPredictor
    data, lenm, outd = tmp_path / "data", tmp_path / "len", tmp_path / "pred"
    for d in (data, lenm):
        d.mkdir()
    spk = [f"p{225 + i}" for i in range(108)]
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    pickle.dump(spk, open(data / "id_to_spkr.pkl", "wb"))
    pickle.dump({s: {"mean": float(mean[i]), "std": float(std[i])} for i, s in enumerate(spk)},
                open(data / "f0_stats.pkl", "wb"))
    len_sd = syn.synthetic_len_predictor_state_dict(100, 108, seed=21)
    torch.save(len_sd, lenm / "best_model.pth")
    torch.save((torch.tensor(2.5), torch.tensor(1.5)), lenm / "len_norm_stats.pth")
    gen = torch.Generator().manual_seed(7)
    rows = []
    for i, n in enumerate([20, 6]):
        runs = torch.randint(0, 100, (n,), generator=gen)
        units = torch.repeat_interleave(runs, torch.randint(1, 4, (n,), generator=gen)).tolist()
        f0 = (float(mean[i]) + float(std[i]) * torch.randn(len(units), generator=gen)).tolist()
        f0[0] = 0.0
        rows.append({"units": units, "f0": f0, "audio": f"{spk[i]}_{i:03d}_mic2.wav"})
    (data / "val.txt").write_text("\n".join(str(r) for r in rows) + "\n")
    df = pd.DataFrame({"syn_sample": [f"{spk[0]}_000", f"{spk[1]}_001"], "syn_trgt": ["p300", "p301"]})
    df.to_csv(tmp_path / "samples.csv")
    pinf.main(["--input_path", str(data / "val.txt"), "--out_path", str(outd), "--pred_len", "--len_model", str(lenm) + "/",
               "--f0_path", str(data / "f0_stats.pkl"), "--vc", "--target_speakers", "p300", "p301", "--device", "cuda:0",
               "--sample_df", str(tmp_path / "samples.csv")])
    assert sorted(os.listdir(outd)) == ["p300_val.txt", "p301_val.txt"]
    lm = LenPredictor(100, 108).to("cuda:0")
    lm.load_state_dict(len_sd)
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(1.5)
    for r, tgt in zip(rows, ("p300", "p301")):
        got = [json.loads(l) for l in open(outd / f"{tgt}_val.txt")]
        assert len(got) == 1 and got[0]["audio"] == r["audio"]
        seq = torch.tensor([r["units"]], device="cuda:0")
        sid = torch.tensor([[spk.index(tgt)]], device="cuda:0")
        out_seq, _, out_len, counts, dd_len = pinf.convert_batch(seq, sid, 100, lm, None, True, return_lens=True)
        assert got[0]["units"] == out_seq[0, :int(out_len[0])].tolist()
        src = spk.index(r["audio"].split("_")[0])
        p = torch.tensor(r["f0"])
        ii = p != 0
        p[ii] = (p[ii] - mean[src]) / std[src]
        want = pinf.morph_seq_len(r["units"], p.numpy(), counts[0, :int(dd_len[0])].cpu().numpy())
        assert len(got[0]["f0"]) == len(got[0]["units"]) and np.allclose(got[0]["f0"], want, atol=1e-6)
