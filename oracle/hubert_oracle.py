"""CPU restatement of the HuBERT-base unit encoder (TEST INFRASTRUCTURE ONLY).  **Parity unpinned.**

The reference calls ``textless.data.speech_encoder.SpeechEncoder`` (data/encode.py:7,21-22,32); its arithmetic lives in
textlesslib (unpinned HEAD, README.md:31-33) and fairseq@dd106d9534b22e7db859a6b87ffd7780c38341f8 (README.md:34), neither
of which is under /root/reference nor installable offline, and the reference holds no tests or vectors for it.  This
file restates the published fairseq graph (fairseq/models/hubert/hubert.py ``HubertModel.extract_features``,
fairseq/models/wav2vec/wav2vec2.py ``ConvFeatureExtractionModel`` mode "default", ``TransformerEncoder`` with
``layer_norm_first=False``, ``TransformerSentenceEncoderLayer`` post-LN; textless ``HubertFeatureReader.get_features``:
no input normalisation for the base model, ``output_layer=6``; ``KMeansQuantizer``: nearest centroid) over a plain
state dict with fairseq's parameter names.  tests/test_hubert_oracle.py cross-checks it against TWO independent
implementations of the same published architecture, each with its own random initialisation:
``torchaudio.models.hubert_base`` and Hugging Face ``transformers.HubertModel`` (``hidden_states[6]``), at 6 000,
96 000 and 160 000 samples; tests/golden/hubert_hf_small.npz pins a few output rows of the transformers model so the
GPU box re-checks the same numbers.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2


def num_frames(n: int) -> int:
    for _, k, s in CONV_LAYERS:
        n = (n - k) // s + 1 if n >= k else 0
    return n


def fold_pos_conv_weight_norm(weight_g: torch.Tensor, weight_v: torch.Tensor) -> torch.Tensor:
    """fairseq wav2vec2.py: ``nn.utils.weight_norm(self.pos_conv, name="weight", dim=2)``: g has shape (1,1,k) and the
    norm runs over dims (0,1) for every tap."""
    norm = weight_v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return weight_v * (weight_g / norm)


def from_torchaudio(model, n_layers: int = 6) -> dict:
    """State dict of ``torchaudio.models.hubert_base()`` under fairseq's names (the inverse of the mapping in
    torchaudio.models.wav2vec2.utils.import_fairseq_model), pos_conv weight-norm folded."""
    sd = {}
    fe = model.feature_extractor.conv_layers
    for i, layer in enumerate(fe):
        sd[f"feature_extractor.conv_layers.{i}.0.weight"] = layer.conv.weight.detach().clone()
        if layer.layer_norm is not None:
            sd[f"feature_extractor.conv_layers.{i}.2.weight"] = layer.layer_norm.weight.detach().clone()
            sd[f"feature_extractor.conv_layers.{i}.2.bias"] = layer.layer_norm.bias.detach().clone()
    fp = model.encoder.feature_projection
    sd["layer_norm.weight"], sd["layer_norm.bias"] = fp.layer_norm.weight.detach().clone(), fp.layer_norm.bias.detach().clone()
    sd["post_extract_proj.weight"] = fp.projection.weight.detach().clone()
    sd["post_extract_proj.bias"] = fp.projection.bias.detach().clone()
    tr = model.encoder.transformer
    sd["encoder.pos_conv.0.weight"] = tr.pos_conv_embed.conv.weight.detach().clone()   # property: g * v / ||v||
    sd["encoder.pos_conv.0.bias"] = tr.pos_conv_embed.conv.bias.detach().clone()
    sd["encoder.layer_norm.weight"] = tr.layer_norm.weight.detach().clone()
    sd["encoder.layer_norm.bias"] = tr.layer_norm.bias.detach().clone()
    for l, layer in enumerate(tr.layers[:n_layers]):
        p = f"encoder.layers.{l}."
        for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
            m = getattr(layer.attention, name)
            sd[p + f"self_attn.{name}.weight"], sd[p + f"self_attn.{name}.bias"] = m.weight.detach().clone(), m.bias.detach().clone()
        sd[p + "self_attn_layer_norm.weight"] = layer.layer_norm.weight.detach().clone()
        sd[p + "self_attn_layer_norm.bias"] = layer.layer_norm.bias.detach().clone()
        sd[p + "fc1.weight"] = layer.feed_forward.intermediate_dense.weight.detach().clone()
        sd[p + "fc1.bias"] = layer.feed_forward.intermediate_dense.bias.detach().clone()
        sd[p + "fc2.weight"] = layer.feed_forward.output_dense.weight.detach().clone()
        sd[p + "fc2.bias"] = layer.feed_forward.output_dense.bias.detach().clone()
        sd[p + "final_layer_norm.weight"] = layer.final_layer_norm.weight.detach().clone()
        sd[p + "final_layer_norm.bias"] = layer.final_layer_norm.bias.detach().clone()
    return sd


def from_transformers(model, n_layers: int = 6) -> dict:
    """State dict of a Hugging Face ``transformers.HubertModel`` (config = hubert-base: group-norm extractor,
    ``do_stable_layer_norm=False``) under fairseq's names -- the inverse of transformers'
    ``convert_hubert_original_pytorch_checkpoint_to_pytorch.py`` mapping -- pos_conv weight-norm (dim 2) folded.
    A SECOND independent implementation of the published graph to check this oracle against."""
    hf = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sd = {}
    for i in range(7):
        sd[f"feature_extractor.conv_layers.{i}.0.weight"] = hf[f"feature_extractor.conv_layers.{i}.conv.weight"]
    sd["feature_extractor.conv_layers.0.2.weight"] = hf["feature_extractor.conv_layers.0.layer_norm.weight"]
    sd["feature_extractor.conv_layers.0.2.bias"] = hf["feature_extractor.conv_layers.0.layer_norm.bias"]
    sd["layer_norm.weight"], sd["layer_norm.bias"] = hf["feature_projection.layer_norm.weight"], hf["feature_projection.layer_norm.bias"]
    sd["post_extract_proj.weight"] = hf["feature_projection.projection.weight"]
    sd["post_extract_proj.bias"] = hf["feature_projection.projection.bias"]
    pc = "encoder.pos_conv_embed.conv."
    if pc + "weight_g" in hf:
        g, v = hf[pc + "weight_g"], hf[pc + "weight_v"]
    else:
        g, v = hf[pc + "parametrizations.weight.original0"], hf[pc + "parametrizations.weight.original1"]
    sd["encoder.pos_conv.0.weight"] = fold_pos_conv_weight_norm(g, v)
    sd["encoder.pos_conv.0.bias"] = hf[pc + "bias"]
    sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"] = hf["encoder.layer_norm.weight"], hf["encoder.layer_norm.bias"]
    for l in range(n_layers):
        p, q = f"encoder.layers.{l}.", f"encoder.layers.{l}."
        for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{name}.weight"] = hf[q + f"attention.{name}.weight"]
            sd[p + f"self_attn.{name}.bias"] = hf[q + f"attention.{name}.bias"]
        sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"] = hf[q + "layer_norm.weight"], hf[q + "layer_norm.bias"]
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = hf[q + "feed_forward.intermediate_dense.weight"], hf[q + "feed_forward.intermediate_dense.bias"]
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = hf[q + "feed_forward.output_dense.weight"], hf[q + "feed_forward.output_dense.bias"]
        sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"] = hf[q + "final_layer_norm.weight"], hf[q + "final_layer_norm.bias"]
    return sd


@torch.no_grad()
def extract_features(sd: dict, wave: torch.Tensor, n_layers: int = 6, n_heads: int = 12, pos_groups: int = 16,
                     dtype=torch.float32) -> torch.Tensor:
    """wave (B,N) -> layer-``n_layers`` features (B,T,D).  No padding mask: every row is a full-length clip
    (textless runs B=1, ``padding_mask=None``)."""
    g = lambda k: sd[k].to(dtype)
    x = wave.to(dtype).unsqueeze(1)
    # ConvFeatureExtractionModel, mode "default": GroupNorm only after conv 0, no conv bias
    x = F.conv1d(x, g("feature_extractor.conv_layers.0.0.weight"), stride=5)
    x = F.group_norm(x, x.shape[1], g("feature_extractor.conv_layers.0.2.weight"), g("feature_extractor.conv_layers.0.2.bias"), eps=1e-5)
    x = F.gelu(x)
    for i in range(1, 7):
        x = F.gelu(F.conv1d(x, g(f"feature_extractor.conv_layers.{i}.0.weight"), stride=2))
    x = x.transpose(1, 2)                                                         # (B,T,512)
    x = F.layer_norm(x, (x.shape[-1],), g("layer_norm.weight"), g("layer_norm.bias"), eps=1e-5)
    x = F.linear(x, g("post_extract_proj.weight"), g("post_extract_proj.bias"))
    # TransformerEncoder.extract_features: x = x + gelu(SamePad(pos_conv(x))); then layer_norm (post-LN model)
    k = sd["encoder.pos_conv.0.weight"].shape[-1]
    xc = F.conv1d(x.transpose(1, 2), g("encoder.pos_conv.0.weight"), g("encoder.pos_conv.0.bias"), padding=k // 2, groups=pos_groups)
    if k % 2 == 0:
        xc = xc[:, :, :-1]                                                        # SamePad
    x = x + F.gelu(xc).transpose(1, 2)
    x = F.layer_norm(x, (x.shape[-1],), g("encoder.layer_norm.weight"), g("encoder.layer_norm.bias"), eps=1e-5)
    B, T, D = x.shape
    hd = D // n_heads
    for l in range(n_layers):
        p = f"encoder.layers.{l}."
        q = F.linear(x, g(p + "self_attn.q_proj.weight"), g(p + "self_attn.q_proj.bias")) * (hd ** -0.5)
        kk = F.linear(x, g(p + "self_attn.k_proj.weight"), g(p + "self_attn.k_proj.bias"))
        v = F.linear(x, g(p + "self_attn.v_proj.weight"), g(p + "self_attn.v_proj.bias"))
        sh = lambda t: t.view(B, T, n_heads, hd).transpose(1, 2)
        a = torch.softmax(sh(q) @ sh(kk).transpose(-1, -2), dim=-1) @ sh(v)
        a = a.transpose(1, 2).reshape(B, T, D)
        a = F.linear(a, g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"))
        x = F.layer_norm(x + a, (D,), g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias"), eps=1e-5)
        h = F.linear(F.gelu(F.linear(x, g(p + "fc1.weight"), g(p + "fc1.bias"))), g(p + "fc2.weight"), g(p + "fc2.bias"))
        x = F.layer_norm(x + h, (D,), g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"), eps=1e-5)
    return x


def kmeans_distances(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """Squared euclidean distances (M,K), direct (x-c)^2 sum (no |x|^2+|c|^2-2xc expansion)."""
    return (x.unsqueeze(1) - centroids.unsqueeze(0)).pow(2).sum(-1)


def kmeans_assign(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """KMeansQuantizer: nearest centroid, lowest index on ties (torch.argmin / sklearn predict semantics)."""
    return kmeans_distances(x, centroids).argmin(-1)
