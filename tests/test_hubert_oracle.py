"""CPU: the HuBERT oracle (oracle/hubert_oracle.py) against the independent torchaudio implementation of the same
published architecture (random weights).  The real textless/fairseq code is not available offline -> parity unpinned."""
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
