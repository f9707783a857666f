"""CPU restatement of the reference prosody predictors and infer.py glue (TEST INFRASTRUCTURE ONLY).

Functional torch code over plain ``{name: tensor}`` state dicts, following
/root/reference/model/len_predictor.py, model/pitch_predictor.py and infer.py line by
line (eval mode: no masking, BatchNorm uses running statistics).  Pinned by
tests/test_oracle_golden.py against tests/golden/predictors.npz and len_carryover.npz,
which were produced by the real reference classes (tests/golden/make_golden.py).
"""
from __future__ import annotations

from itertools import groupby

import numpy as np
import torch
import torch.nn.functional as F


def _conv(sd, name, x, pad):
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def _bn(sd, name, x):
    # nn.BatchNorm1d in eval(): running statistics, eps 1e-5
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=1e-5)


def _embed(sd, seq, spk_id, pe=None):
    """token_emb(seq) ++ repeat(spk_emb(spk_id)) [+ pe]   (len_predictor.py:36-42, pitch_predictor.py:73-80)."""
    emb_seq = F.embedding(seq.long(), sd["token_emb.weight"])
    emb_spk = torch.repeat_interleave(F.embedding(spk_id.long(), sd["spk_emb.weight"]), seq.shape[-1], dim=1)
    if pe is not None:
        emb_spk = emb_spk + pe[:, :emb_spk.size(1)]                      # PositionalEncoding.forward :31-38
    return torch.cat([emb_seq, emb_spk], dim=-1).transpose(1, 2)


@torch.no_grad()
def len_predictor_forward(sd, seq, spk_id, norm_mean, norm_std):
    """LenPredictor.forward, model/len_predictor.py:35-52."""
    x = _embed(sd, seq, spk_id)
    for sfx in ["1"] + [f"1{i}" for i in range(1, 7)]:
        x = F.leaky_relu(_bn(sd, "bn" + sfx, _conv(sd, "cnn" + sfx, x, 1)))   # :44-50
    return _conv(sd, "cnn2", x, 1).squeeze(1) * norm_std + norm_mean           # :52


@torch.no_grad()
def pitch_predictor_forward(sd, seq, spk_id, kind="new"):
    """PitchPredictor.forward (:72-94) / PitchPredictorBase.forward (:145-166) -> (class_logits, reg)."""
    base = kind == "base"
    x = _embed(sd, seq, spk_id, None if base else sd["pe.pe"])
    for sfx in ["1"] + [f"1{i}" for i in range(1, 8)]:
        x = _conv(sd, "cnn" + sfx, x, 1)
        if base:
            x = _bn(sd, "bn" + sfx, x)
        x = F.leaky_relu(x)
    x = _conv(sd, "cnn2", x, 1)
    if not base:
        x = _bn(sd, "bn2", x)
    x = F.leaky_relu(x)
    c = _conv(sd, "cnn_class1", x, 1)
    r = _conv(sd, "cnn_reg1", x, 1)
    if base:
        c, r = _bn(sd, "bn_c1", c), _bn(sd, "bn_r1", r)
    c, r = F.leaky_relu(c), F.leaky_relu(r)
    return _conv(sd, "cnn_class2", c, 0).squeeze(1), _conv(sd, "cnn_reg2", r, 0).squeeze(1)


def calc_freq(class_preds, reg_preds, spk_id, id2pitch_mean=None, id2pitch_std=None, norm=False):
    """model/pitch_predictor.py:100-104."""
    mask = class_preds > 0
    if not norm:
        reg_preds = id2pitch_mean[spk_id.long()] + reg_preds * id2pitch_std[spk_id.long()]
    return mask * reg_preds


def dedup_seq(seq):
    """dataset/utils.py:14-16."""
    vals, counts = zip(*[(k, sum(1 for _ in g)) for k, g in groupby(seq)])
    return list(vals), list(counts)


def len_carryover_correction(lens: np.ndarray) -> np.ndarray:
    """infer.py:158-172 with explicit fp32 arithmetic: lens (1,L) float32 -> (L,) int64."""
    lens = np.asarray(lens, dtype=np.float32)
    r = np.rint(np.maximum(lens[0], np.float32(1.0))).astype(np.float32)   # torch.round = half to even
    a = (lens - r)[0].astype(np.float32)
    total = np.float32(0.0)
    vals = []
    for n in a:
        total = np.float32(total + n)
        if total >= 1:
            vals.append(1)
            total = np.float32(total - np.float32(1.0))
        elif total <= -1:
            vals.append(-1)
            total = np.float32(total + np.float32(1.0))
        else:
            vals.append(0)
    return r.astype(np.int64) + np.asarray(vals, dtype=np.int64)


@torch.no_grad()
def infer_sample(units, spk_id, n_tokens, len_sd=None, len_stats=None, pitch_sd=None, pitch_kind="new",
                 id2pitch_mean=None, id2pitch_std=None, norm_pitch=True):
    """_infer_sample (infer.py:24-45) for the pred_len + pred_pitch mode: units (list/1-D) -> (out units, f0)."""
    seq = torch.as_tensor(units, dtype=torch.int64)
    seq = seq[seq != n_tokens].view(1, -1)
    spk = torch.as_tensor([[int(spk_id)]])
    if len_sd is not None:
        dd, _ = dedup_seq(seq[0].tolist())
        dd = torch.tensor(dd).unsqueeze(0)
        lens = len_predictor_forward(len_sd, dd, spk, len_stats[0], len_stats[1])
        lens_i = torch.from_numpy(len_carryover_correction(lens.numpy()))
        out_seq = torch.repeat_interleave(dd, lens_i.clamp(min=0)).view(1, -1)
    else:
        out_seq = seq
    c, r = pitch_predictor_forward(pitch_sd, out_seq, spk, pitch_kind)
    f0 = calc_freq(c, r, spk, id2pitch_mean, id2pitch_std, norm_pitch)
    return out_seq[0], f0[0]
