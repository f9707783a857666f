"""CPU: the HuBERT oracle (oracle/hubert_oracle.py) against TWO independent implementations of the same published
architecture (torchaudio's and Hugging Face transformers', each with its own random weights) and against the committed
rows of the transformers model (tests/golden/hubert_hf_small.npz).  The reference's own encoder (textlesslib / fairseq)
is third-party code that is not available offline, so it cannot generate vectors itself."""
import numpy as np
import pytest
import torch

from oracle import hubert_oracle as ho


class FakeKM:  # stands in for sklearn.cluster.MiniBatchKMeans in a km.bin look-alike
    pass


@pytest.fixture(scope="module")
def ta_model():
    torchaudio = pytest.importorskip("torchaudio")
    torch.manual_seed(0)
    return torchaudio.models.hubert_base().eval()


def test_frame_count():
    assert ho.num_frames(96000) == 299 and ho.num_frames(32000) == 99 and ho.num_frames(400) == 1 and ho.num_frames(399) == 0
    for n in (400, 719, 720, 16000, 12345):
        assert ho.num_frames(n) == (n - 400) // 320 + 1


def test_oracle_matches_torchaudio_layer6(ta_model):
    sd = ho.from_torchaudio(ta_model, 6)
    g = torch.Generator().manual_seed(1)
    wave = 0.1 * torch.randn(2, 6000, generator=g)
    with torch.no_grad():
        ref = ta_model.extract_features(wave, num_layers=6)[0][-1]
    got = ho.extract_features(sd, wave, 6)
    assert got.shape == ref.shape == (2, ho.num_frames(6000), 768)
    assert (got - ref).abs().max().item() < 2e-4


@pytest.fixture(scope="module")
def hf_model():
    pytest.importorskip("transformers")
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_hubert as mg
    return mg, mg.build_model()


@pytest.mark.parametrize("n", [6000, 96000, 160000])
def test_oracle_matches_transformers_layer6(hf_model, n):
    """Second independent graph: transformers.HubertModel hidden_states[6], incl. BASELINE configs[3]'s clip
    lengths (96 000 samples, varlen up to 160 000)."""
    mg, model = hf_model
    sd = ho.from_transformers(model, 6)
    g = torch.Generator().manual_seed(n)
    wave = 0.1 * torch.randn(1, n, generator=g)
    with torch.no_grad():
        ref = model(wave, output_hidden_states=True).hidden_states[6]
    got = ho.extract_features(sd, wave, 6)
    assert got.shape == ref.shape == (1, ho.num_frames(n), 768)
    assert (got - ref).abs().max().item() < 5e-5


def test_oracle_matches_committed_transformers_rows(hf_model):
    from _util import load_golden
    mg, model = hf_model
    gold = load_golden("hubert_hf_small.npz")
    if abs(mg.weights_checksum(model) - float(gold["checksum"])) > 1e-6 * float(gold["checksum"]):
        pytest.skip("this torch build initialises HubertModel differently from the one that wrote the fixture")
    sd = ho.from_transformers(model, 6)
    for n, w in mg.waves().items():
        assert np.array_equal(w[0, :8].numpy(), gold[f"wave_head_{n}"])
        got = ho.extract_features(sd, w, 6)[0][gold[f"rows_{n}"]]
        assert np.abs(got.numpy() - gold[f"feat_{n}"]).max() < 5e-5


def test_pos_conv_weight_norm_fold(ta_model):
    conv = ta_model.encoder.transformer.pos_conv_embed.conv
    params = dict(conv.named_parameters())
    g = params.get("weight_g", params.get("parametrizations.weight.original0"))
    v = params.get("weight_v", params.get("parametrizations.weight.original1"))
    assert g.shape == (1, 1, 128)
    assert torch.allclose(ho.fold_pos_conv_weight_norm(g, v), conv.weight, atol=1e-6)


def test_kmeans_first_index_on_ties():
    c = torch.tensor([[0.0, 0.0], [2.0, 0.0], [0.0, 0.0]])
    x = torch.tensor([[1.0, 0.0], [0.1, 0.0], [1.9, 0.0]])
    assert ho.kmeans_assign(x, c).tolist() == [0, 0, 1]


def test_checkpoint_loaders_on_lookalikes(tmp_path, ta_model):
    """fairseq-style .pt with a config object of an unknown class, and a pickled KMeans look-alike."""
    from dissc_b200 import checkpoints as ck
    import argparse
    import pickle
    sd = ho.from_torchaudio(ta_model, 1)
    torch.save({"args": argparse.Namespace(arch="hubert"), "model": sd}, tmp_path / "hubert.pt")
    got = ck.load_fairseq_hubert(str(tmp_path / "hubert.pt"))
    assert set(got) == set(sd) and torch.equal(got["layer_norm.weight"], sd["layer_norm.weight"])

    km = FakeKM()
    km.cluster_centers_ = np.arange(12, dtype=np.float64).reshape(3, 4)
    with open(tmp_path / "km.bin", "wb") as f:
        pickle.dump(km, f)
    c = ck.load_kmeans_centers(str(tmp_path / "km.bin"))
    assert c.dtype == torch.float32 and c.shape == (3, 4) and c[2, 3] == 11


def test_real_joblib_kmeans_dump_and_hostile_pickles(tmp_path):
    """km.bin as the GSLM release ships it: joblib.dump of a fitted sklearn MiniBatchKMeans -- loaded here WITHOUT
    instantiating any sklearn class; and pickles that try to call into os / builtins are defused, not executed."""
    import pickle
    joblib = pytest.importorskip("joblib")
    cluster = pytest.importorskip("sklearn.cluster")
    from dissc_b200 import checkpoints as ck
    rng = np.random.RandomState(0)
    km = cluster.MiniBatchKMeans(n_clusters=7, n_init=1, random_state=0, batch_size=64).fit(rng.randn(300, 12))
    for name, kw in (("km.bin", {}), ("km_z.bin", {"compress": 3})):
        joblib.dump(km, tmp_path / name, **kw)
        c = ck.load_kmeans_centers(str(tmp_path / name))
        assert c.dtype == torch.float32 and c.shape == (7, 12)
        assert np.allclose(c.numpy(), km.cluster_centers_.astype(np.float32))

    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, (f"touch {marker}",))

    for payload, found in ((Evil(), False), ({"cluster_centers_": np.ones((2, 3)), "cfg": Evil()}, True)):
        with open(tmp_path / "evil.bin", "wb") as f:
            pickle.dump(payload, f)
        try:
            c = ck.load_kmeans_centers(str(tmp_path / "evil.bin"))
            assert found and c.shape == (2, 3)
        except ValueError:
            assert not found                       # "no cluster_centers_ found"
        assert not marker.exists()
    torch.save({"model": {"w": torch.ones(2)}, "cfg": Evil()}, tmp_path / "evil.pt")
    got = ck.load_fairseq_hubert(str(tmp_path / "evil.pt"))
    assert not marker.exists() and torch.equal(got["w"], torch.ones(2))
    assert ("builtins", "eval") not in ck._ALLOWED_GLOBALS and ("os", "system") not in ck._ALLOWED_GLOBALS
