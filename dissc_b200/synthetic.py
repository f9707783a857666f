"""Seeded synthetic checkpoints and inputs (there are no pretrained weights offline).

The reference's default init (``init_weights`` std 0.01, sr/utils.py:32-35)
gives an output std of ~0.003, which makes a 1e-4 parity tolerance vacuous,
and a naive He init saturates tanh.  The recipe below
keeps activations O(1) through all 97 convolutions and lands the pre-tanh
signal at std ~0.5, so parity checks exercise the full dynamic range.

Everything here is host-side data generation; no model arithmetic.
"""
from __future__ import annotations

import math

import torch

VCTK_CONFIG = {
    # sr/configs/VCTK/hubert100_lut.json (geometry + inference-relevant keys)
    "resblock": "1",
    "upsample_rates": [5, 4, 4, 2, 2],
    "upsample_kernel_sizes": [11, 8, 8, 4, 4],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "num_embeddings": 100,
    "embedding_dim": 128,
    "model_in_dim": 257,
    "code_hop_size": 320,
    "f0": True,
    "multispkr": "_",
    "sampling_rate": 16000,
}

# A small geometry with the same structure (5 stages, x320, MRF of 3 ResBlock1)
# whose full state dict fits in a committed fixture.
SMALL_CONFIG = dict(VCTK_CONFIG, upsample_initial_channel=64, embedding_dim=12, model_in_dim=25)

_GAIN = math.sqrt(2.0 / (1.0 + 0.1 ** 2))


def _wn_pair(gen, shape, fan_in, scale=1.0):
    """weight_v ~ N(0, gain^2/fan_in); weight_g = ||v|| * scale * U(0.8,1.2) per dim-0 slice."""
    v = torch.randn(shape, generator=gen) * (_GAIN / math.sqrt(fan_in))
    norm = v.reshape(shape[0], -1).norm(dim=1)
    jitter = 0.8 + 0.4 * torch.rand(shape[0], generator=gen)
    g = (norm * jitter * scale).reshape(shape[0], *([1] * (len(shape) - 1)))
    return g, v


RECIPES = ("calibrated", "hot", "init_weights", "torch_default")


def synthetic_generator_state_dict(h: dict, seed: int = 0, post_std: float = 0.5, recipe: str = "calibrated") -> dict:
    """Checkpoint-format (``weight_g``/``weight_v``/``bias``) state dict for
    ``CodeGenerator(h)`` -- key set identical to what ``sr/train.py:206-214`` saves.

    ``recipe``:
      * ``calibrated``    the O(1)-activation recipe described above (default; the benchmark weights);
      * ``hot``           the same with 2x the ``convs2`` gain, 3x bias and a pre-tanh std of 1.2 (tanh saturates often);
      * ``init_weights``  every conv weight ~ N(0, 0.01) as ``init_weights`` asks for (sr/utils.py:32-35), ``g = ||v||``,
                          biases as ``nn.Conv1d`` initialises them, U(+-1/sqrt(fan_in)): activations shrink by orders of
                          magnitude through the upsamplers (output std ~3e-3), which exercises the SMALL end of the
                          split-fp16 operand range;
      * ``torch_default`` what a freshly constructed reference ``Generator`` really holds (``init_weights`` writes the
                          derived ``weight`` attribute that old-style weight-norm recomputes from g / v at the next forward,
                          so the effective weights are ``nn.Conv1d``'s own kaiming-uniform init)."""
    if recipe not in RECIPES:
        raise ValueError(f"unknown recipe {recipe!r}; expected one of {RECIPES}")
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    c0 = h["upsample_initial_channel"]
    cin = h.get("model_in_dim", 128)
    hot = recipe == "hot"
    if hot:
        post_std = 1.2

    def put(prefix, shape, fan_in, scale=1.0):
        nb = shape[1] if prefix.startswith("ups.") else shape[0]
        if recipe in ("init_weights", "torch_default"):
            torch_fan_in = shape[1] * shape[2]   # nn.init._calculate_fan_in_and_fan_out: size(1) * receptive field
            bound = 1.0 / math.sqrt(torch_fan_in)
            if recipe == "init_weights":
                v = 0.01 * torch.randn(shape, generator=gen)
            else:
                v = (2 * torch.rand(shape, generator=gen) - 1) * bound   # kaiming_uniform_(a=sqrt(5))
            g = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], *([1] * (len(shape) - 1)))
            sd[prefix + ".weight_g"], sd[prefix + ".weight_v"] = g, v
            sd[prefix + ".bias"] = (2 * torch.rand(nb, generator=gen) - 1) * bound
            return
        if hot and ".convs2." in prefix:
            scale *= 2.0
        g, v = _wn_pair(gen, shape, fan_in, scale)
        sd[prefix + ".weight_g"], sd[prefix + ".weight_v"] = g, v
        sd[prefix + ".bias"] = (0.03 if hot else 0.01) * torch.randn(nb, generator=gen)

    put("conv_pre", (c0, cin, 7), cin * 7)
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        ci, co = c0 // (2 ** i), c0 // (2 ** (i + 1))
        put(f"ups.{i}", (ci, co, k), ci * k / u)
        ch = co
        for j, (rk, rd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * len(h['resblock_kernel_sizes']) + j}"
            if h["resblock"] == "1":
                for m in range(len(rd)):
                    put(f"{p}.convs1.{m}", (ch, ch, rk), ch * rk)
                    put(f"{p}.convs2.{m}", (ch, ch, rk), ch * rk, scale=0.25)
            else:
                for m in range(len(rd)):
                    put(f"{p}.convs.{m}", (ch, ch, rk), ch * rk, scale=0.25)
    # conv_post: MRF output has std ~1.3-1.5 with this recipe; gain picked so the
    # pre-tanh signal has std ~post_std (measured with the reference).
    put("conv_post", (1, ch, 7), ch * 7, scale=post_std / 1.4)
    sd["dict.weight"] = torch.randn(h["num_embeddings"], h["embedding_dim"], generator=gen)
    if h.get("multispkr", None):
        sd["spkr.weight"] = torch.randn(200, h["embedding_dim"], generator=gen)
    return sd


def synthetic_inputs(batch: int, frames: int, seed: int = 1234, n_units: int = 100, n_spkr: int = 108):
    """code ~ U{0..n_units-1}; f0 ~ N(0,1) with 30% of
    frames exactly 0 (unvoiced); spkr ~ U{0..n_spkr-1}."""
    gen = torch.Generator().manual_seed(seed)
    code = torch.randint(0, n_units, (batch, frames), generator=gen, dtype=torch.int64)
    f0 = torch.randn(batch, 1, frames, generator=gen)
    f0[torch.rand(batch, 1, frames, generator=gen) < 0.3] = 0.0
    spkr = torch.randint(0, n_spkr, (batch, 1), generator=gen, dtype=torch.int64)
    return code, f0, spkr


def state_dict_checksum(sd: dict) -> float:
    """Order-independent fp64 checksum used by fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64).reshape(v.shape).remainder(7).add(1)).sum())
    return tot


# --------------------------------------------------------------------------
# prosody predictors (model/len_predictor.py, model/pitch_predictor.py)
# --------------------------------------------------------------------------
def _conv_bn_entries(gen, sd, conv, cin, cout, k, bn=None):
    sd[f"{conv}.weight"] = torch.randn(cout, cin, k, generator=gen) * math.sqrt(2.0 / (cin * k))
    sd[f"{conv}.bias"] = 0.1 * torch.randn(cout, generator=gen)
    if bn:
        sd[f"{bn}.weight"] = 1.0 + 0.1 * torch.randn(cout, generator=gen)
        sd[f"{bn}.bias"] = 0.1 * torch.randn(cout, generator=gen)
        sd[f"{bn}.running_mean"] = 0.1 * torch.randn(cout, generator=gen)
        sd[f"{bn}.running_var"] = 0.5 + torch.rand(cout, generator=gen)
        sd[f"{bn}.num_batches_tracked"] = torch.tensor(100, dtype=torch.int64)


def synthetic_len_predictor_state_dict(n_tokens=100, n_speakers=108, emb=32, seed=0):
    """Keys/shapes of ``LenPredictor.state_dict()`` (model/len_predictor.py:15-33), O(1) activations,
    non-trivial BatchNorm running statistics."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    sd["token_emb.weight"] = torch.randn(n_tokens + 1, emb, generator=gen)
    sd["token_emb.weight"][n_tokens] = 0.0  # padding_idx row
    sd["spk_emb.weight"] = torch.randn(n_speakers, emb, generator=gen)
    _conv_bn_entries(gen, sd, "cnn1", 2 * emb, 128, 3, "bn1")
    for i in range(1, 7):
        _conv_bn_entries(gen, sd, f"cnn1{i}", 128, 128, 3, f"bn1{i}")
    _conv_bn_entries(gen, sd, "cnn2", 128, 1, 3)
    return sd


def synthetic_pitch_predictor_state_dict(kind="new", n_tokens=100, n_speakers=108, emb=32, seed=0):
    """Keys/shapes of ``PitchPredictor`` ("new", model/pitch_predictor.py:41-70) or
    ``PitchPredictorBase`` ("base", :106-143)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    sd["token_emb.weight"] = torch.randn(n_tokens + 1, emb, generator=gen)
    sd["token_emb.weight"][n_tokens] = 0.0
    sd["spk_emb.weight"] = torch.randn(n_speakers + 1, emb, generator=gen)
    sd["spk_emb.weight"][n_speakers] = 0.0
    base = kind == "base"
    if not base:
        max_len = 850  # PositionalEncoding buffer, model/pitch_predictor.py:7-17
        pe_start = torch.repeat_interleave(torch.linspace(0, 1, max_len).unsqueeze(-1), emb // 2, dim=-1)
        pe_end = torch.repeat_interleave(torch.linspace(1, 0, max_len).unsqueeze(-1), emb // 2, dim=-1)
        sd["pe.pe"] = torch.cat([pe_start, pe_end], dim=-1).unsqueeze(0)
    _conv_bn_entries(gen, sd, "cnn1", 2 * emb, 128, 3, "bn1" if base else None)
    for i in range(1, 8):
        _conv_bn_entries(gen, sd, f"cnn1{i}", 128, 128, 3, f"bn1{i}" if base else None)
    _conv_bn_entries(gen, sd, "cnn2", 128, 128, 3, None if base else "bn2")
    _conv_bn_entries(gen, sd, "cnn_class1", 128, 128, 3, "bn_c1" if base else None)
    _conv_bn_entries(gen, sd, "cnn_class2", 128, 1, 1)
    _conv_bn_entries(gen, sd, "cnn_reg1", 128, 128, 3, "bn_r1" if base else None)
    _conv_bn_entries(gen, sd, "cnn_reg2", 128, 1, 1)
    return sd


def synthetic_pitch_stats(n_speakers=108, seed=0):
    gen = torch.Generator().manual_seed(seed)
    mean = 180 + 40 * torch.randn(n_speakers, generator=gen)
    std = 30 + 8 * torch.rand(n_speakers, generator=gen)
    return mean, std
