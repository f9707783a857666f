"""CPU restatement of the reference vocoder forward (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/sr/models.py line by line, using the same ATen calls
the reference makes (``F.conv1d`` / ``F.conv_transpose1d`` / ``F.leaky_relu``),
so on a CPU it reproduces the reference's numerics (oneDNN convolutions,
threaded over all host cores).  Works on a plain ``{name: tensor}`` state dict
in the reference's checkpoint format, so it has no dependency on the
reference's classes and travels to the GPU box.

Pinned by tests/test_oracle_golden.py against vectors produced by the real
``sr/models.py::CodeGenerator`` (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # sr/models.py:13


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """sr/utils.py:44-45."""
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(weight_g: torch.Tensor, weight_v: torch.Tensor) -> torch.Tensor:
    """``torch.nn.utils.weight_norm`` (dim=0) folded the way
    ``remove_weight_norm`` does it (sr/models.py:116-122, :43-47):
    ``w = g * v / ||v||`` with the norm over every dim but 0."""
    norm = weight_v.reshape(weight_v.shape[0], -1).norm(dim=1).reshape(weight_g.shape)
    return weight_v * (weight_g / norm)


def _w(sd: dict, prefix: str) -> torch.Tensor:
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])


def folded_state_dict(sd: dict) -> dict:
    """Checkpoint dict (weight_g/weight_v) -> plain weight/bias dict."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            out[p + ".weight"] = fold_weight_norm(v, sd[p + ".weight_v"])
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def _upsample(signal: torch.Tensor, max_frames: int) -> torch.Tensor:
    """sr/models.py:158-177 (nearest-repeat; raises on misalignment)."""
    if signal.dim() == 3:
        bsz, channels, cond_length = signal.size()
    elif signal.dim() == 2:
        signal = signal.unsqueeze(2)
        bsz, channels, cond_length = signal.size()
    else:
        signal = signal.view(-1, 1, 1)
        bsz, channels, cond_length = signal.size()
    signal = signal.unsqueeze(3).repeat(1, 1, 1, max_frames // cond_length)
    reminder = (max_frames - signal.shape[2] * signal.shape[3]) // signal.shape[3]
    if reminder > 0:
        raise NotImplementedError("Padding condition signal - misalignment between condition features.")
    return signal.view(bsz, channels, max_frames)


def resblock1(sd: dict, prefix: str, x: torch.Tensor, kernel_size: int, dilation) -> torch.Tensor:
    """sr/models.py:34-41."""
    for m, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs1.{m}"), sd[f"{prefix}.convs1.{m}.bias"],
                      stride=1, padding=get_padding(kernel_size, d), dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs2.{m}"), sd[f"{prefix}.convs2.{m}.bias"],
                      stride=1, padding=get_padding(kernel_size, 1), dilation=1)
        x = xt + x
    return x


def resblock2(sd: dict, prefix: str, x: torch.Tensor, kernel_size: int, dilation) -> torch.Tensor:
    """sr/models.py:62-67."""
    for m, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs.{m}"), sd[f"{prefix}.convs.{m}.bias"],
                      stride=1, padding=get_padding(kernel_size, d), dilation=d)
        x = xt + x
    return x


def generator_forward(sd: dict, h: dict, x: torch.Tensor, return_intermediates: bool = False):
    """sr/models.py:98-114 (Generator.forward) on an already-built input x (B,model_in_dim,T)."""
    inter = {}
    x = F.conv1d(x, _w(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)           # :99
    inter["conv_pre"] = x
    num_kernels = len(h["resblock_kernel_sizes"])
    rb = resblock1 if h["resblock"] == "1" else resblock2
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)                                           # :101
        x = F.conv_transpose1d(x, _w(sd, f"ups.{i}"), sd[f"ups.{i}.bias"],
                               stride=u, padding=(k - u) // 2)                     # :102, :84-86
        inter[f"ups.{i}"] = x
        xs = None
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            r = rb(sd, f"resblocks.{i * num_kernels + j}", x, rk, rd)
            if xs is None:                                                         # :104-108
                xs = r
            else:
                xs += r
        x = xs / num_kernels                                                       # :109
        inter[f"mrf.{i}"] = x
    x = F.leaky_relu(x)                                                            # :110 (slope 0.01!)
    x = F.conv1d(x, _w(sd, "conv_post"), sd["conv_post.bias"], padding=3)          # :111
    inter["conv_post"] = x
    x = torch.tanh(x)                                                              # :112
    if return_intermediates:
        return x, inter
    return x


def build_input(sd: dict, h: dict, code: torch.Tensor, f0: torch.Tensor | None,
                spkr: torch.Tensor | None, **feats) -> torch.Tensor:
    """sr/models.py:189, :206-221 -- the live branch for the shipped configs
    (no code_vq / f0 vq / f0 quantizer); ``feats`` = any further keyword arguments of
    ``CodeGenerator.forward`` (e.g. ``f0_stats`` for ``f0_feats`` configs)."""
    x = F.embedding(code, sd["dict.weight"]).transpose(1, 2)                       # :189
    if h.get("f0", None):
        if x.shape[-1] < f0.shape[-1]:                                             # :207-210
            x = _upsample(x, f0.shape[-1])
        else:
            f0 = _upsample(f0, x.shape[-1])
        x = torch.cat([x, f0.to(x.dtype)], dim=1)                                  # :211
    if h.get("multispkr", None):
        s = F.embedding(spkr, sd["spkr.weight"]).transpose(1, 2)                   # :213
        s = _upsample(s, x.shape[-1])                                              # :214
        x = torch.cat([x, s], dim=1)                                               # :215
    for k, feat in feats.items():                                                  # :216-221
        feat = _upsample(feat.to(x.dtype), x.shape[-1])
        x = torch.cat([x, feat], dim=1)
    return x


@torch.no_grad()
def code_generator_forward(sd: dict, h: dict, code, f0=None, spkr=None, dtype=torch.float32,
                           return_intermediates: bool = False, **feats):
    """CodeGenerator.forward (sr/models.py:179-225) for the shipped configs.
    ``sd`` is a checkpoint-format (weight_g/weight_v) or folded state dict."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    if any(k.endswith(".weight_g") for k in sd):
        sd = folded_state_dict(sd)
    f0 = None if f0 is None else f0.to(dtype)
    x = build_input(sd, h, code, f0, spkr, **feats)
    return generator_forward(sd, h, x, return_intermediates)


def generate_int16(y: torch.Tensor):
    """sr/inference.py:73-75: squeeze, *32768, numpy astype int16 (wraps, no clip)."""
    import numpy as np
    audio = y.squeeze() * 32768.0
    return audio.cpu().numpy().astype(np.int16)
