"""GPU parity of the HuBERT-base unit encoder (through the C ABI) against the CPU oracle, against transformers'
HubertModel directly and against its committed rows (tests/golden/hubert_hf_small.npz).  The oracle restates the published
fairseq graph and is cross-checked against torchaudio's and transformers' independent implementations
(tests/test_hubert_oracle.py); the reference's own third-party encoder is not available offline."""
import numpy as np
import pytest
import torch

from oracle import hubert_oracle as ho

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    torchaudio = pytest.importorskip("torchaudio")
    torch.manual_seed(0)
    model = torchaudio.models.hubert_base().eval()
    sd = ho.from_torchaudio(model, 6)
    g = torch.Generator().manual_seed(7)
    lens = [16000, 9000, 12345, 400]
    waves = [0.1 * torch.randn(n, generator=g) for n in lens]
    feats = [ho.extract_features(sd, w.view(1, -1), 6)[0] for w in waves]          # each clip alone, unpadded
    allf = torch.cat(feats, 0)
    idx = torch.randperm(allf.shape[0], generator=g)[:100]
    cent = allf[idx] + 0.02 * torch.randn(100, 768, generator=g)
    return sd, lens, waves, feats, cent


def test_features_and_units_varlen_batch(cuda_device, setup):
    from dissc_b200.hubert import SpeechEncoder
    sd, lens, waves, feats, cent = setup
    enc = SpeechEncoder.from_state_dict(sd, cent).to(cuda_device)
    N = max(lens)
    wave = torch.zeros(len(lens), N)
    for b, w in enumerate(waves):
        wave[b, :len(w)] = w
        wave[b, len(w):] = 7.0          # garbage past the valid samples must not matter
    units, n_frames, dense = enc.encode_batch(wave.to(cuda_device), torch.tensor(lens, dtype=torch.int32))
    units, n_frames, dense = units.cpu(), n_frames.cpu(), dense.cpu()
    agree = total = 0
    worst = 0.0
    for b, f in enumerate(feats):
        T = ho.num_frames(lens[b])
        assert int(n_frames[b]) == T == f.shape[0]
        err = (dense[b, :T] - f).abs().max().item()
        worst = max(worst, err)
        assert torch.all(units[b, T:] == -1) and torch.all(dense[b, T:] == 0)
        d = ho.kmeans_distances(f.double(), cent.double())
        want = d.argmin(-1)
        top2 = d.topk(2, dim=-1, largest=False).values
        margin = (top2[:, 1] - top2[:, 0]) / top2[:, 1].clamp(min=1e-12)
        sure = margin > 1e-3                         # near-ties may legitimately flip under ~1e-4 feature noise
        assert torch.equal(units[b, :T][sure], want[sure]), f"clip {b}: unit mismatch outside near-ties"
        agree += int((units[b, :T] == want).sum())
        total += T
    assert worst < 2e-4, f"layer-6 feature max-abs error {worst}"
    assert agree / total > 0.98
    print(f"hubert: feature max-abs err {worst:.2e}; units agree {agree}/{total}")


def _unit_checks(units, dense_gpu, feat64, cent, tag):
    """units: what the GPU assigned; dense_gpu: ITS OWN layer-6 features; feat64: the fp64 oracle's features.
    (1) the quantiser is exact on the features it was given, up to the rounding of the form it evaluates: the assignment
        computes |c|^2 - 2 x.c (the fairseq / sklearn form) as a split-fp16 GEMM with fp32 accumulation, so the fp64
        distance of the chosen centroid may exceed the fp64 minimum by at most 1e-6 of the distance plus 2e-6 of
        |x|^2 + max|c|^2 -- i.e. it IS the fp64 argmin except between centroids that tie at fp32 resolution;
    (2) against the oracle's units: a frame may differ only if the feature noise can provably flip it, i.e. the GPU's
        choice is within 2 * (|x - c_a| + |x - c_b|) * |delta| + |delta|^2 of the oracle's minimum distance, with delta
        the measured feature difference of that frame."""
    c64 = cent.double()
    d_own = ho.kmeans_distances(dense_gpu.double(), c64)
    best = d_own.min(-1).values
    chosen = d_own.gather(1, units.view(-1, 1)).view(-1)
    scale = dense_gpu.double().pow(2).sum(-1) + c64.pow(2).sum(-1).max()
    excess = chosen - best
    assert torch.all(excess <= 1e-6 * best + 2e-6 * scale), \
        f"{tag}: k-means assign is not the argmin of its input (excess {excess.max():.3e}, scale {scale.max():.3e})"
    assert (units == d_own.argmin(-1)).double().mean() > 0.9, f"{tag}: too many fp32-resolution ties for a meaningful check"
    d_ref = ho.kmeans_distances(feat64, c64)
    want = d_ref.argmin(-1)
    diff = units != want
    if diff.any():
        delta = (dense_gpu.double() - feat64).norm(dim=-1)[diff]
        da, db = d_ref[diff, units[diff]], d_ref[diff, want[diff]]
        slack = 2 * (da.sqrt() + db.sqrt()) * delta + delta * delta
        assert torch.all(da - db <= slack), f"{tag}: unit differs from the oracle beyond what the feature noise explains"
    return int((~diff).sum()), int(diff.numel())


def test_config4_clip_lengths_vs_transformers(cuda_device):
    """BASELINE configs[3] shapes: a 96 000-sample clip, a 160 000-sample clip (varlen upper end) and a short one in one
    padded batch, weights = a seeded transformers.HubertModel.  Features <= 2e-4 max-abs from the fp64 oracle, from
    transformers' own forward and from its committed rows; units per _unit_checks."""
    pytest.importorskip("transformers")
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_hubert as mg
    from _util import load_golden
    from dissc_b200.hubert import SpeechEncoder
    model = mg.build_model()
    sd = ho.from_transformers(model, 6)
    gold = load_golden("hubert_hf_small.npz")
    gold_ok = abs(mg.weights_checksum(model) - float(gold["checksum"])) <= 1e-6 * float(gold["checksum"])
    ws = mg.waves()
    g = torch.Generator().manual_seed(3)
    clips = [ws[96000][0], ws[160000][0], 0.1 * torch.randn(32000, generator=g)]
    lens = [len(c) for c in clips]
    feats64 = [ho.extract_features(sd, c.view(1, -1), 6, dtype=torch.float64)[0] for c in clips]
    allf = torch.cat(feats64, 0).float()
    cent = allf[torch.randperm(allf.shape[0], generator=g)[:100]] + 0.02 * torch.randn(100, 768, generator=g)
    enc = SpeechEncoder.from_state_dict(sd, cent).to(cuda_device)
    wave = torch.full((len(clips), max(lens)), 3.0)          # garbage past the valid samples must not matter
    for b, c in enumerate(clips):
        wave[b, :len(c)] = c
    units, n_frames, dense = enc.encode_batch(wave.to(cuda_device), torch.tensor(lens, dtype=torch.int32))
    units, n_frames, dense = units.cpu(), n_frames.cpu(), dense.cpu()
    worst = agree = total = 0
    for b, f64 in enumerate(feats64):
        T = ho.num_frames(lens[b])
        assert int(n_frames[b]) == T == f64.shape[0]
        worst = max(worst, (dense[b, :T].double() - f64).abs().max().item())
        a, n = _unit_checks(units[b, :T], dense[b, :T], f64, cent, f"clip {b}")
        agree, total = agree + a, total + n
    assert worst < 2e-4, f"layer-6 feature max-abs error vs the fp64 oracle {worst}"
    with torch.no_grad():
        hf = model(clips[0].view(1, -1), output_hidden_states=True).hidden_states[6][0]
    assert (dense[0, :299] - hf).abs().max().item() < 2e-4
    if gold_ok:
        for b, n in ((0, 96000), (1, 160000)):
            assert np.abs(dense[b][gold[f"rows_{n}"]].numpy() - gold[f"feat_{n}"]).max() < 2e-4
    assert agree / total > 0.99
    print(f"hubert 96k/160k/32k: feature max-abs err {worst:.2e}; units agree with the fp64 oracle {agree}/{total}")


def test_attention_tensor_core_path_at_config4_length(cuda_device):
    """96 000-sample clips (299 frames: BASELINE configs[3]) take the tcgen05 attention kernel (<= 320 keys per CTA): a
    ragged batch vs the fp64 oracle, and vs the fp32 CUDA-core attention on the same inputs (both within the feature
    tolerance of the oracle; units per _unit_checks)."""
    pytest.importorskip("transformers")
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_hubert as mg
    from dissc_b200 import _lib
    from dissc_b200.hubert import SpeechEncoder
    model = mg.build_model()
    sd = ho.from_transformers(model, 6)
    g = torch.Generator().manual_seed(11)
    lens = [96000, 70001, 96000, 400, 33333]
    clips = [0.1 * torch.randn(n, generator=g) for n in lens]
    feats64 = [ho.extract_features(sd, c.view(1, -1), 6, dtype=torch.float64)[0] for c in clips]
    allf = torch.cat(feats64, 0).float()
    cent = allf[torch.randperm(allf.shape[0], generator=g)[:100]] + 0.02 * torch.randn(100, 768, generator=g)
    wave = torch.full((len(clips), max(lens)), -2.0)
    for b, c in enumerate(clips):
        wave[b, :len(c)] = c
    out = {}
    try:
        for mode in (1, 0):
            _lib.check(_lib.lib().dissc_tc_set_tuning(4, mode))
            enc = SpeechEncoder.from_state_dict(sd, cent).to(cuda_device)
            units, n_frames, dense = enc.encode_batch(wave.to(cuda_device), torch.tensor(lens, dtype=torch.int32))
            out[mode] = (units.cpu(), n_frames.cpu(), dense.cpu())
    finally:
        _lib.lib().dissc_tc_set_tuning(4, 1)
    for mode, (units, n_frames, dense) in out.items():
        worst = 0.0
        for b, f64 in enumerate(feats64):
            T = ho.num_frames(lens[b])
            assert int(n_frames[b]) == T
            worst = max(worst, (dense[b, :T].double() - f64).abs().max().item())
            assert torch.all(units[b, T:] == -1) and torch.all(dense[b, T:] == 0)
            _unit_checks(units[b, :T], dense[b, :T], f64, cent, f"mode {mode} clip {b}")
        print(f"attention {'tcgen05' if mode else 'fp32 CUDA-core'}: feature max-abs err vs fp64 oracle {worst:.2e}")
        assert worst < 2e-4, (mode, worst)
    assert (out[1][2] - out[0][2]).abs().max().item() < 2e-4


def test_call_surface_matches_data_encode(cuda_device, setup):
    from dissc_b200.hubert import SpeechEncoder
    sd, lens, waves, feats, cent = setup
    enc = SpeechEncoder.from_state_dict(sd, cent, deduplicate=False).to(cuda_device)
    out = enc(waves[1].view(1, -1).to(cuda_device))            # data/encode.py:32
    assert set(out) == {"units", "durations", "dense"}
    T = ho.num_frames(lens[1])
    assert out["units"].shape == (T,) and out["units"].dtype == torch.int64
    assert torch.all(out["durations"] == 1) and out["dense"].shape == (T, 768)
    dd = SpeechEncoder.from_state_dict(sd, cent, deduplicate=True).to(cuda_device)(waves[1].to(cuda_device))
    assert int(dd["durations"].sum()) == T and torch.equal(torch.repeat_interleave(dd["units"], dd["durations"]), out["units"])
    # B=1 equals the same clip inside a padded batch
    wave = torch.zeros(2, lens[0])
    wave[0], wave[1, :lens[1]] = waves[0], waves[1]
    u, nf, _ = enc.encode_batch(wave.to(cuda_device), torch.tensor([lens[0], lens[1]], dtype=torch.int32))
    assert torch.equal(u[1, :T], out["units"])


def test_encoder_kmeans_gemm_and_large_codebook(cuda_device, setup):
    """The encoder's two assignment paths: K <= 128 (one more GEMM + row argmin) and K > 128 (CUDA-core kernel)."""
    from dissc_b200.hubert import SpeechEncoder
    sd, lens, waves, feats, cent = setup
    g = torch.Generator().manual_seed(11)
    wave = waves[0].view(1, -1).to(cuda_device)
    # duplicate centroid: identical columns give identical distances, the lower index must win
    dup = cent.clone()
    dup[37] = dup[5]
    u, _, dense = SpeechEncoder.from_state_dict(sd, dup).to(cuda_device).encode_batch(wave)
    assert not bool((u == 37).any())
    _unit_checks(u[0].cpu(), dense[0].cpu(), feats[0].double(), dup, "K=100 with a duplicate")
    # 200 centroids: beyond the 128 GEMM columns
    allf = torch.cat(feats, 0)
    big = torch.cat([cent, allf[torch.randperm(allf.shape[0], generator=g)[:100]] + 0.02 * torch.randn(100, 768, generator=g)])
    u2, _, dense2 = SpeechEncoder.from_state_dict(sd, big).to(cuda_device).encode_batch(wave)
    same, n = _unit_checks(u2[0].cpu(), dense2[0].cpu(), feats[0].double(), big, "K=200")
    assert same / n > 0.98 and bool((u2 >= 128).any())
    assert torch.equal(dense.cpu(), dense2.cpu())


def test_kmeans_assign_exact(cuda_device):
    from dissc_b200.hubert import kmeans_assign
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 768, generator=g)
    c = torch.randn(100, 768, generator=g)
    c[37] = c[5]                                               # exact tie -> lowest index must win
    x[0] = c[5]
    got = kmeans_assign(x.to(cuda_device), c).cpu()
    d = ho.kmeans_distances(x.double(), c.double())
    top2 = d.topk(2, dim=-1, largest=False).values
    sure = (top2[:, 1] - top2[:, 0]) > 1e-3
    assert got[0] == 5
    assert torch.equal(got[sure], d.argmin(-1)[sure])
    # small integer-valued data: distances are exact in fp32 -> bit-exact argmin incl. ties
    xi = torch.randint(-3, 4, (500, 64), generator=g).float()
    ci = torch.randint(-3, 4, (50, 64), generator=g).float()
    assert torch.equal(kmeans_assign(xi.to(cuda_device), ci).cpu(), ho.kmeans_assign(xi, ci))


def test_short_clip_rejected(cuda_device, setup):
    from dissc_b200.hubert import SpeechEncoder
    sd, _, _, _, cent = setup
    enc = SpeechEncoder.from_state_dict(sd, cent).to(cuda_device)
    with pytest.raises(ValueError):
        enc.encode_batch(torch.zeros(1, 399, device=cuda_device))


def test_encode_cli_end_to_end(cuda_device, setup, tmp_path):
    """dissc_b200.encode (data/encode.py surface): wav files -> JSON lines, from fairseq-style checkpoint files."""
    import json
    from scipy.io import wavfile
    from dissc_b200 import encode as enc_cli
    sd, lens, waves, feats, cent = setup
    wav_dir = tmp_path / "wav"
    wav_dir.mkdir()
    names = []
    for i, w in enumerate(waves[:3]):
        pcm = (w.clamp(-1, 1) * 32767).round().to(torch.int16)
        wavfile.write(wav_dir / f"p225_{i:03d}.wav", 16000, pcm.numpy())
        names.append(f"p225_{i:03d}.wav")
    torch.save({"cfg": {"model": "hubert"}, "model": sd}, tmp_path / "hubert.pt")
    np.save(tmp_path / "km.npy", cent.numpy())
    out = tmp_path / "hubert100" / "train.txt"
    enc_cli.main(["--base_dir", str(wav_dir), "--out_file", str(out), "--device", "cuda:0",
                  "--hubert_checkpoint", str(tmp_path / "hubert.pt"), "--kmeans_path", str(tmp_path / "km.npy")])
    rows = [json.loads(l) for l in open(out)]
    assert [r["audio"] for r in rows] == names
    for r, n in zip(rows, lens[:3]):
        T = ho.num_frames(n)
        assert len(r["units"]) == T and r["durations"] == [1] * T and set(r) == {"units", "durations", "audio"}
        pcm = torch.from_numpy(wavfile.read(wav_dir / r["audio"])[1].astype(np.float32) / 32768.0)
        f = ho.extract_features(sd, pcm.view(1, -1), 6)[0]
        d = ho.kmeans_distances(f.double(), cent.double())
        top2 = d.topk(2, dim=-1, largest=False).values
        sure = (top2[:, 1] - top2[:, 0]) / top2[:, 1].clamp(min=1e-12) > 1e-3
        assert torch.equal(torch.tensor(r["units"])[sure], d.argmin(-1)[sure])


_RANGE_CHECK_SCRIPT = r"""
import sys, torch, torchaudio
from oracle import hubert_oracle as ho
from dissc_b200.hubert import SpeechEncoder
torch.manual_seed(0)
sd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
cent = torch.randn(100, 768)
wave = 0.1 * torch.randn(2, 8000, device="cuda")
SpeechEncoder.from_state_dict(sd, cent).to("cuda").encode_batch(wave)          # in range: passes
hot = dict(sd)
k = [n for n in hot if n.endswith("post_extract_proj.weight")][0]
hot[k] = hot[k] * 1e6                                                           # projection output far beyond 65504
try:
    SpeechEncoder.from_state_dict(hot, cent).to("cuda").encode_batch(wave)
except RuntimeError as e:
    assert "fp16 limit" in str(e), e
    print("RANGE_CHECK_OK")
    sys.exit(0)
sys.exit("saturated activations were not reported")
"""


def test_range_check_reports_saturated_activations(cuda_device):
    """DISSC_HUB_RANGE_CHECK=1 (read once per process, hence the subprocess): a forward whose activations hit the fp16
    limit on their way into the split planes fails loudly instead of quantising silently."""
    import os
    import subprocess
    import sys
    pytest.importorskip("torchaudio")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DISSC_HUB_RANGE_CHECK="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", _RANGE_CHECK_SCRIPT], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "RANGE_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


_PAIR_SCRIPT = r"""
import sys, torch, torchaudio
from oracle import hubert_oracle as ho
from dissc_b200.hubert import SpeechEncoder
torch.manual_seed(0)
sd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
g = torch.Generator().manual_seed(5)
lens = [12000, 7000, 9001]
waves = [0.1 * torch.randn(n, generator=g) for n in lens]
cent = torch.randn(100, 768, generator=g)
wave = torch.zeros(len(lens), max(lens))
for b, w in enumerate(waves):
    wave[b, :len(w)] = w
units, nf, dense = SpeechEncoder.from_state_dict(sd, cent).to("cuda").encode_batch(wave.cuda(), torch.tensor(lens, dtype=torch.int32))
worst = 0.0
for b, w in enumerate(waves):
    f = ho.extract_features(sd, w.view(1, -1), 6)[0]
    T = f.shape[0]
    assert int(nf[b]) == T
    worst = max(worst, (dense[b, :T].cpu() - f).abs().max().item())
assert worst < 2e-4, worst
print("PAIR_OK %.2e" % worst)
"""


def test_cta_pair_gemms_opt_in(cuda_device):
    """DISSC_HUB_PAIR2=1: the encoder's GEMMs as CTA pairs issuing 256-row tcgen05.mma.cta_group::2 (conv_tc.cuh; opt-in,
    measured 4 % slower than the single-CTA form): same features within the encoder's 2e-4 bound."""
    import os
    import subprocess
    import sys
    pytest.importorskip("torchaudio")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DISSC_HUB_PAIR2="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", _PAIR_SCRIPT], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PAIR_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
