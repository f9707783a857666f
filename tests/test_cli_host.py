"""CPU tests of the host-side CLI logic (dissc_b200/inference.py, dissc_b200/infer.py): manifest parsing,
F0 normalisation, checkpoint discovery, batching.  No GPU, no compute calls."""
import json
import os

import numpy as np
import pytest

from dissc_b200 import inference as inf
from dissc_b200 import infer as pinf


def test_parse_manifest_dict_and_json_lines(tmp_path):
    m = tmp_path / "val.txt"
    m.write_text("{'units': [1, 1, 2], 'f0': [0.0, 110.5, 0], 'audio': 'some/dir/p225_001_mic2.wav'}\n"
                 + json.dumps({"units": [5], "f0": [99.0], "audio": "p226_002.wav"}) + "\n")
    files, codes, pitch = inf.parse_manifest(str(m), "/base")
    assert [str(f) for f in files] == ["/base/p225_001_mic2.wav", "/base/p226_002.wav"]
    assert codes[0].tolist() == [1, 1, 2] and codes[0].dtype == np.int64
    assert pitch[0].tolist() == [0.0, 110.5, 0.0]
    assert pinf.parse_line("{'units': [3], 'audio': 'a'}") == {"units": [3], "audio": "a"}


def test_normalize_f0_only_touches_voiced_frames():
    f0 = np.array([0, 100, 0, 200], dtype=np.float64)
    out = inf.normalize_f0(f0, 150.0, 50.0)
    assert out.dtype == np.float32 and out.tolist() == [0.0, -1.0, 0.0, 1.0]


def test_prepare_items_speaker_ids_and_stats():
    h = inf.AttrDict(f0=True, multispkr="_", f0_normalize=True)
    files = ["/b/p226_001.wav", "/b/p999_002.wav"]
    codes = [np.array([1, 2]), np.array([3])]
    pitch = [np.array([0.0, 120.0]), np.array([90.0])]
    stats = {"p226": {"mean": 100.0, "std": 10.0}, "f0_mean": 80.0, "f0_std": 5.0}
    items = inf.prepare_items(h, files, codes, pitch, ["p225", "p226", "p999"], stats)
    assert items[0]["spkr"] == 1 and items[1]["spkr"] == 2
    assert items[0]["f0"].tolist() == [0.0, 2.0]
    assert items[1]["f0"].tolist() == [2.0]          # unknown speaker -> global f0_mean / f0_std
    with pytest.raises(NotImplementedError):
        inf.prepare_items(inf.AttrDict(f0=True, multispkr="_"), files, codes, [], ["p226", "p999"], None)


def test_scan_checkpoint_and_config(tmp_path):
    for n in ("g_00000010", "g_00000200", "do_00000200"):
        (tmp_path / n).write_text("x")
    (tmp_path / "config.json").write_text(json.dumps({"sampling_rate": 16000}))
    assert os.path.basename(inf.scan_checkpoint(str(tmp_path), "g_")) == "g_00000200"
    assert inf.scan_checkpoint(str(tmp_path), "zz_") == ""
    assert inf.load_config(str(tmp_path)).sampling_rate == 16000
    assert inf.load_config(str(tmp_path / "g_00000010")).sampling_rate == 16000


def test_batches_by_length_and_peak_normalize():
    lens = [10, 300, 299, 5, 120]
    b = inf.batches_by_length(lens, max_batch=2, max_frames=10_000)
    assert b == [[1, 2], [4, 0], [3]]
    b = inf.batches_by_length(lens, max_batch=64, max_frames=600)
    assert b[0] == [1, 2] and sorted(sum(b, [])) == [0, 1, 2, 3, 4]
    x = np.array([0, -16384, 8192], dtype=np.int16)
    assert inf.peak_normalize(x).tolist() == [0.0, -1.0, 0.5]
    assert inf.peak_normalize(np.zeros(4, np.int16)).tolist() == [0.0] * 4


def test_cli_parsers_keep_reference_flags():
    a = inf.build_parser().parse_args(["--checkpoint_file", "ck", "--vc", "--target-speakers", "p231", "p239"])
    assert a.eval_mode is True and a.n == 2508 and a.target_speakers == ["p231", "p239"]
    p = pinf.build_parser().parse_args(["--pred_len", "--pred_pitch", "--vc", "--target_speakers", "p231"])
    assert p.norm_pitch is True and p.n == 10 and p.f0_model_type == "new" and p.n_tokens == 100


# ---------------------------------------------------------------------------------------------
# flags ported later: --f0-stats, --sample_df, --code_file, _gt.wav (sr/inference.py), morph_seq_len / --sample_df (infer.py)
# ---------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_morph_seq_len_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "morph_seq_len.npz"))
    n = len([k for k in g.files if k.startswith("out")])
    assert n >= 6
    for i in range(n):
        got = pinf.morph_seq_len(g[f"units{i}"], g[f"pitch{i}"], g[f"lens{i}"])
        assert got.shape == g[f"out{i}"].shape
        assert np.array_equal(np.asarray(got, dtype=np.float64), np.asarray(g[f"out{i}"], dtype=np.float64))


def test_rescale_f0_follows_reference_arithmetic():
    # sr/inference.py:220-235 restated with torch ops on a (1,1,T) tensor, as the reference holds it
    import torch
    f0 = torch.tensor([[[0.0, 110.0, 0.0, 130.0, 90.0, 0.0, 100.0]]])
    ref = f0.clone()
    ii = ref != 0
    mean_, std_ = ref[ii].mean(), ref[ii].std()
    ref[ii] -= mean_
    ref[ii] /= std_
    ref[ii] *= 21.0
    ref[ii] += 185.0
    got = inf.rescale_f0(f0.view(-1).numpy(), 185.0, 21.0)
    assert np.array_equal(got, ref.view(-1).numpy())
    stats = {3: {"f0_mean": 200.0, "f0_std": 30.0}, "f0_mean": 150.0, "f0_std": 25.0}
    assert inf.target_f0_stats(stats, 3) == (200.0, 30.0) and inf.target_f0_stats(stats, 7) == (150.0, 25.0)


def test_code_file_sample_df_and_gt_audio(tmp_path):
    cf = tmp_path / "codes.txt"
    cf.write_text("utt_a|1 2 3 3\nutt_b|7\n\n")
    items = inf.parse_code_file(str(cf))
    assert [it["name"] for it in items] == ["utt_a", "utt_b"] and items[0]["code"].tolist() == [1, 2, 3, 3]
    import pandas as pd
    df = pd.DataFrame({"syn_sample": ["p225_001", "p225_001", "p226_002"], "syn_trgt": ["p231", "p239", "p231"]})
    spk = {"p225": 0, "p231": 1, "p239": 2}
    assert inf.sample_df_targets(df, "p225_001_mic2", spk) == [1, 2]
    assert inf.sample_df_targets(df, "p999_001_mic2", spk) == []
    from scipy.io import wavfile
    x = (np.array([0, 1000, -2000, 500], dtype=np.int16))
    wavfile.write(str(tmp_path / "a.wav"), 16000, x)
    gt = inf.load_gt_audio(tmp_path / "a.wav", 16000, pad=8)
    assert gt.shape == (8,) and gt.dtype == np.float32                  # padded to a multiple of `pad`
    assert np.allclose(gt[:4], 0.95 * x / 2000.0) and np.all(gt[4:] == 0)
    assert inf.load_gt_audio(tmp_path / "a.wav", 22050) is None          # other rate: skipped (resampy absent)
    assert inf.load_gt_audio(tmp_path / "missing.wav", 16000) is None
    assert np.allclose(inf.peak_normalize_f32(gt)[:4], x / 2000.0)


def test_wav_writer_background_threads(tmp_path):
    from scipy.io import wavfile
    w = inf.WavWriter(workers=3)
    want = {}
    for i in range(7):
        x = (np.arange(50, dtype=np.int32) * (i + 1) - 100).astype(np.int16)
        w.submit(str(tmp_path / f"a{i}.wav"), 16000, x)
        want[f"a{i}.wav"] = inf.peak_normalize(x)
    g = np.linspace(-0.3, 0.6, 40).astype(np.float32)
    w.submit(str(tmp_path / "gt.wav"), 16000, g)          # float input: normalised as a float signal
    w.close()
    for name, y in want.items():
        rate, got = wavfile.read(tmp_path / name)
        assert rate == 16000 and got.dtype == np.float32 and np.array_equal(got, y)
    assert np.allclose(wavfile.read(tmp_path / "gt.wav")[1], g / 0.6)
    bad = inf.WavWriter()
    bad.submit(str(tmp_path / "no_such_dir" / "x.wav"), 16000, np.zeros(4, np.int16))
    with pytest.raises(OSError):
        bad.close()


def test_select_items_follows_reference_shuffle_and_n():
    """sr/inference.py:340-359: --debug walks the manifest until i > n (n + 2 entries); the pool path shuffles with the
    RNG state CodeDataset.__init__ leaves behind (random.seed(1234), sr/dataset.py:157) and awaits n + 1 results."""
    import random
    from dissc_b200.inference import select_items, vc_targets_for_item
    assert select_items(10, 3, debug=True) == [0, 1, 2, 3, 4]
    assert select_items(3, 7, debug=True) == [0, 1, 2]
    assert select_items(4, -1, debug=True) == [0, 1, 2, 3]
    random.seed(1234)                      # what the reference's global RNG holds when main() shuffles
    want = list(range(50))
    random.shuffle(want)
    assert select_items(50, -1, debug=False) == want
    assert select_items(50, 9, debug=False) == want[:10]
    # VC targets: five distinct speakers per utterance, a function of the manifest index only
    a, b = vc_targets_for_item(7, 108), vc_targets_for_item(7, 108)
    assert a == b and len(set(a)) == 5 and all(0 <= k < 108 for k in a)
    assert vc_targets_for_item(8, 108) != a
    assert len(vc_targets_for_item(0, 3)) == 3


def test_f0_median_and_f0_feats_items():
    """sr/dataset.py:297-315: f0_median fills the unvoiced frames with the voiced median before normalising;
    f0_feats adds the source speaker's [mean, std] as the `f0_stats` feature."""
    from dissc_b200.inference import normalize_f0, prepare_items
    from dissc_b200 import AttrDict
    f0 = np.array([0.0, 100.0, 0.0, 140.0, 120.0], dtype=np.float32)
    got = normalize_f0(f0, 110.0, 20.0, f0_median=True)
    want = (np.array([120.0, 100.0, 120.0, 140.0, 120.0]) - 110.0) / 20.0
    assert np.allclose(got, want)
    assert np.allclose(normalize_f0(f0, 110.0, 20.0), [0, -0.5, 0, 1.5, 0.5])
    h = AttrDict(f0=True, f0_normalize=True, f0_median=True, f0_feats=True, multispkr="_")
    stats = {"p225": {"mean": 110.0, "std": 20.0}, "f0_mean": 150.0, "f0_std": 30.0}
    items = prepare_items(h, ["/x/p225_001.wav", "/x/p999_001.wav"], [[1, 2, 3, 4, 5]] * 2, [f0, f0], ["p225", "p999"], stats)
    assert np.allclose(items[0]["f0"], want) and np.allclose(items[0]["f0_stats"], [110.0, 20.0])
    assert np.allclose(items[1]["f0_stats"], [150.0, 30.0])     # unknown speaker: the global statistics
