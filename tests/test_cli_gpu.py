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
