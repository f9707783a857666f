"""Prosody conversion (rhythm + pitch) on the GPU -- the host side of the reference's ``infer.py``.

``_infer_sample`` (infer.py:24-45) runs, per utterance and target speaker: strip pad tokens -> ``dedup_seq``
(itertools.groupby on the host) -> ``LenPredictor`` -> ``len_carryover_correction`` (a Python loop over a CUDA tensor:
one device sync per token) -> ``repeat_interleave`` -> ``PitchPredictor.infer_freq`` -> one JSON line.  Here the same
steps run for a whole padded batch on the device (csrc/predictors.cu): one launch each for dedup, carry-over and
repeat_interleave, batched predictor forwards with per-row ``lengths``, and a single host read-back (the total output
length) per batch.  Wire format and CLI flags are the reference's.
"""
from __future__ import annotations

import argparse
import ast
import ctypes
import json
import os
import pickle
from typing import List, Optional, Sequence

import torch

from . import _lib
from .predictors import LenPredictor, PitchPredictor, PitchPredictorBase, _p, _stream


# ---------------------------------------------------------------------------------------------------------
# device glue
# ---------------------------------------------------------------------------------------------------------
def dedup_units(seq: torch.Tensor, pad_token: int, lengths: Optional[torch.Tensor] = None):
    """Batched ``seqs[seqs != n_tokens]`` + ``dedup_seq`` (infer.py:25-27, dataset/utils.py:14-16).
    seq int64 (B,L) on CUDA -> (dd int64 (B,L) padded with pad_token, counts int32 (B,L), dd_len int32 (B))."""
    if not seq.is_cuda:
        raise _lib.DisscError("dedup_units runs only on CUDA")
    seq = seq.to(torch.int64).contiguous()
    B, L = seq.shape
    dev = seq.device
    dd = torch.empty_like(seq)
    counts = torch.empty((B, L), dtype=torch.int32, device=dev)
    dd_len = torch.empty((B,), dtype=torch.int32, device=dev)
    if lengths is not None:
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_dedup_units(_p(seq), _p(lengths), int(pad_token), B, L, _p(dd), _p(counts),
                                                _p(dd_len), _stream(dev)), "dissc_dedup_units")
    return dd, counts, dd_len


def len_carryover_correction(lens: torch.Tensor, lengths: Optional[torch.Tensor] = None, return_totals=False):
    """``len_carryover_correction`` (infer.py:158-172).  Reference call shape: lens (1,L) fp32 -> (L,) int64.
    A (B,L) batch with ``lengths`` returns (B,L) int64 (0 past each row's valid length)."""
    if not lens.is_cuda:
        raise _lib.DisscError("len_carryover_correction runs only on CUDA")
    lens = lens.to(torch.float32).contiguous()
    B, L = lens.shape
    dev = lens.device
    out = torch.empty((B, L), dtype=torch.int32, device=dev)
    totals = torch.empty((B,), dtype=torch.int32, device=dev)
    if lengths is not None:
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_len_carryover(_p(lens), _p(lengths), B, L, _p(out), _p(totals), _stream(dev)),
                   "dissc_len_carryover")
    res = out.to(torch.int64)
    if B == 1 and lengths is None:
        res = res[0]
    return (res, totals) if return_totals else res


def repeat_interleave(dd: torch.Tensor, counts: torch.Tensor, dd_len: torch.Tensor, pad_token: int, L_out: int):
    """Batched ``torch.repeat_interleave(dd_seq, lens)`` (infer.py:32) -> (out int64 (B,L_out), out_len int32 (B))."""
    B, L = dd.shape
    dev = dd.device
    out = torch.empty((B, L_out), dtype=torch.int64, device=dev)
    out_len = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_repeat_interleave(_p(dd.contiguous()), _p(counts.to(torch.int32).contiguous()),
                                                      _p(dd_len.to(torch.int32).contiguous()), int(pad_token), B, L,
                                                      int(L_out), _p(out), _p(out_len), _stream(dev)),
                   "dissc_repeat_interleave")
    return out, out_len


@torch.no_grad()
def convert_batch(seqs: torch.Tensor, spk_id: torch.Tensor, n_tokens: int, len_model: Optional[LenPredictor] = None,
                  pitch_model=None, norm_pitch: bool = True, lengths: Optional[torch.Tensor] = None,
                  return_lens: bool = False):
    """Batched ``_infer_sample`` (infer.py:24-45); ``f0`` is None without a pitch model (the caller then interpolates the
    original contour with ``morph_seq_len``).  ``return_lens`` adds the per-unit predicted lengths (B,Ld) and dd_len.

    seqs int64 (B,L) padded with ``n_tokens``; spk_id (B,1) target speakers.
    -> (out_seq int64 (B,L_out) padded with n_tokens, f0 fp32 (B,L_out) or None, out_len int32 (B))."""
    dev = seqs.device
    B = seqs.shape[0]
    dd, counts, dd_len = dedup_units(seqs, n_tokens, lengths)
    new_counts = counts
    if len_model is not None:
        Ld = max(int(dd_len.max().item()), 1)
        dd_c = dd[:, :Ld].contiguous()
        lens = len_model(dd_c, spk_id, lengths=dd_len)
        new_counts, totals = len_carryover_correction(lens, dd_len, return_totals=True)
        if new_counts.dim() == 1:
            new_counts = new_counts.view(1, -1)
        L_out = max(int(totals.max().item()), 1)
        out_seq, out_len = repeat_interleave(dd_c, new_counts, dd_len, n_tokens, L_out)
    else:
        # no rhythm prediction: keep the (pad-stripped) input sequence: repeat every deduped unit by its own run length
        L_out = max(int(counts.sum(dim=1).max().item()), 1)
        out_seq, out_len = repeat_interleave(dd, counts, dd_len, n_tokens, L_out)
    f0 = None
    if pitch_model is not None:
        f0 = pitch_model.infer_freq(out_seq, spk_id, norm_pitch, lengths=out_len)
    if return_lens:
        return out_seq, f0, out_len, new_counts, dd_len
    return out_seq, f0, out_len


def interp(vals, target_len: int):
    """utils.py:39-45: nearest-neighbour resampling of one run's pitch values to ``target_len`` points."""
    import numpy as np
    vals = list(vals)
    cur_len = len(vals)
    if cur_len == 1:
        return np.array(int(target_len) * vals)
    if target_len == cur_len:
        return np.array(vals)
    from scipy.interpolate import interp1d
    return interp1d(np.linspace(0., 1., cur_len), vals, bounds_error=False, kind="nearest", fill_value=0)(
        np.linspace(0., 1., int(target_len)))


def morph_seq_len(units, pitch, t_lens):
    """utils.py:47-52 -- the heuristic used when the pitch is NOT predicted (``--pred_len`` without ``--pred_pitch``):
    every run of equal units keeps its own pitch values, stretched / squeezed to the run's predicted length.  Host numpy,
    exactly as in the reference (a few hundred values per utterance)."""
    import numpy as np
    from itertools import groupby
    out = []
    for i, (_, g) in enumerate(groupby(zip(units, pitch), key=lambda x: x[0])):
        out.append(interp([f for _, f in g], int(t_lens[i])))
    return np.concatenate(out) if out else np.zeros((0,))


def infer_sample(seqs, pitch, spk_id, name, out_path, len_model=None, pitch_model=None, norm_pitch=False, n_tokens=100):
    """Signature-compatible ``_infer_sample`` (infer.py:24-45), B=1: appends one JSON line to ``out_path``.
    Without a pitch model the original ``pitch`` contour is morphed to the predicted run lengths (utils.morph_seq_len)."""
    out_seq, f0, out_len, new_counts, dd_len = convert_batch(seqs.view(1, -1), spk_id.view(1, 1), n_tokens, len_model,
                                                             pitch_model, norm_pitch, return_lens=True)
    n = int(out_len[0].item())
    units = out_seq[0, :n].cpu().numpy().tolist()
    for m in (len_model, pitch_model):
        if m is not None:
            m.check_indices(synchronize=False)   # the .cpu() above synchronised
    if pitch_model is not None:
        pitches = f0[0, :n].cpu().numpy().tolist()
    else:
        in_seq = seqs.view(-1)
        in_seq = in_seq[in_seq != n_tokens].cpu().numpy()
        pitches = morph_seq_len(in_seq, torch.as_tensor(pitch).cpu().numpy(),
                                new_counts[0, :int(dd_len[0])].cpu().numpy()).tolist()
    out = {"units": units, "f0": pitches, "audio": name}
    with open(out_path, "a+") as f:
        f.write(f"{json.dumps(out)}\n")
    return out


# ---------------------------------------------------------------------------------------------------------
# manifest / stats plumbing (dataset/pitch_dataset.py:21-42, dataset/utils.py:18-26)
# ---------------------------------------------------------------------------------------------------------
def parse_line(line: str) -> dict:
    """One manifest line: a JSON object or a Python dict literal (the reference ``eval()``s it)."""
    line = line.strip()
    try:
        return json.loads(line)
    except json.JSONDecodeError:
        return ast.literal_eval(line)


def prep_stats_tensors(spk_id_dict: dict, f0_param_dict: dict):
    """dataset/utils.py:18-26."""
    mean = torch.empty(len(spk_id_dict))
    std = torch.empty(len(spk_id_dict))
    for n, v in spk_id_dict.items():
        mean[v] = f0_param_dict[n]["mean"]
        std[v] = f0_param_dict[n]["std"]
    return mean, std


def load_models(args, n_speakers, id2pitch_mean, id2pitch_std, device):
    """infer.py:66-84."""
    len_model = pitch_model = None
    if args.pred_len:
        len_model = LenPredictor(n_tokens=args.n_tokens, n_speakers=n_speakers).to(device)
        len_model.load_state_dict(torch.load(args.len_model + "best_model.pth", map_location="cpu"))
        len_model.norm_mean, len_model.norm_std = torch.load(args.len_model + "len_norm_stats.pth", map_location="cpu")
    if args.pred_pitch:
        cls = PitchPredictorBase if args.f0_model_type == "base" else PitchPredictor
        pitch_model = cls(args.n_tokens, n_speakers, id2pitch_mean=id2pitch_mean.to(device),
                          id2pitch_std=id2pitch_std.to(device)).to(device)
        pitch_model.load_state_dict(torch.load(args.f0_model + "best_model.pth", map_location="cpu"))
    return len_model, pitch_model


def run(args, batch_size: int = 256):
    """``infer`` / ``infer_wild`` (infer.py:47-155) with utterances batched; output files and line format unchanged:
    ``<out_path>/<basename(input)>`` (reconstruction) and ``<out_path>/<target>_<basename(input)>`` (voice conversion)."""
    device = torch.device(args.device)
    spk_pkl = args.id_to_spkr if args.wild_sample else f"{os.path.dirname(args.input_path)}/id_to_spkr.pkl"
    with open(spk_pkl, "rb") as f:
        spk_id_dict = {v: k for k, v in enumerate(pickle.load(f))}
    with open(args.f0_path, "rb") as f:
        f0_param_dict = pickle.load(f)
    mean, std = prep_stats_tensors(spk_id_dict, f0_param_dict)
    len_model, pitch_model = load_models(args, len(spk_id_dict), mean, std, device)
    base = os.path.basename(args.input_path)
    os.makedirs(args.out_path, exist_ok=True)
    rows = [parse_line(l) for l in open(args.input_path) if l.strip()]
    if not args.wild_sample:
        rows = rows[:args.n]
    df = None
    if args.sample_df and not args.wild_sample:                # infer.py:50-51
        import pandas as pd
        df = pd.read_csv(args.sample_df, index_col=0)
    vc_targets: List[str] = []
    if args.wild_sample:
        vc_targets = list(args.target_speakers or [])
    elif args.vc:                                              # infer.py:87-92
        if args.target_speakers:
            vc_targets = list(args.target_speakers)
        else:
            import random
            vc_targets = random.sample(list(spk_id_dict.keys()), k=min(1, len(spk_id_dict)))
    targets: List[Optional[str]] = []
    if not args.wild_sample and df is None:
        targets.append(None)                       # reconstruction with the source speaker (infer.py:113)
    targets += vc_targets
    if df is not None and vc_targets:              # only the conversions the table lists per sample (infer.py:118-119)
        per_row = [set(df[df.syn_sample == os.path.splitext(r["audio"])[0].split("_mic2")[0]].syn_trgt.unique())
                   for r in rows]
        targets = sorted({t for ts in per_row for t in ts}, key=str)
    else:
        per_row = None
    for t in set(targets) | set(vc_targets):
        p = f"{args.out_path}/{base}" if t is None else f"{args.out_path}/{t}_{base}"
        if os.path.exists(p):
            os.remove(p)
    morph = pitch_model is None                    # infer.py:39-40: interpolate the original contour heuristically
    for t in targets:
        sel = list(range(len(rows))) if per_row is None else [i for i in range(len(rows)) if t in per_row[i]]
        p = f"{args.out_path}/{base}" if t is None else f"{args.out_path}/{t}_{base}"
        for i0 in range(0, len(sel), batch_size):
            chunk = [rows[i] for i in sel[i0:i0 + batch_size]]
            L = max(len(r["units"]) for r in chunk)
            seqs = torch.full((len(chunk), L), args.n_tokens, dtype=torch.int64)
            for b, r in enumerate(chunk):
                seqs[b, :len(r["units"])] = torch.as_tensor(r["units"], dtype=torch.int64)
            seqs = seqs.to(device)
            src_ids = None
            if not args.wild_sample:
                src_ids = [spk_id_dict[r["audio"].split("_")[0]] for r in chunk]
            if t is None:
                spk = torch.tensor([[i] for i in src_ids], device=device)
            else:
                spk = torch.full((len(chunk), 1), spk_id_dict[t], device=device)
            out_seq, f0, out_len, new_counts, dd_len = convert_batch(seqs, spk, args.n_tokens, len_model, pitch_model,
                                                                     args.norm_pitch, return_lens=True)
            out_seq, out_len = out_seq.cpu(), out_len.cpu()
            for m in (len_model, pitch_model):
                if m is not None:
                    m.check_indices(synchronize=False)   # the .cpu() above synchronised
            f0 = None if f0 is None else f0.cpu()
            if morph:
                new_counts, dd_len = new_counts.cpu().numpy(), dd_len.cpu().numpy()
            with open(p, "a+") as f:
                for b, r in enumerate(chunk):
                    n = int(out_len[b])
                    if not morph:
                        pitches = f0[b, :n].tolist()
                    else:
                        # the ORIGINAL contour, normalised with the SOURCE speaker's statistics (infer.py:104-108: done
                        # before spk_id is overwritten with the target), stretched per run (utils.morph_seq_len)
                        pitch = torch.tensor(r["f0"], dtype=torch.float32)
                        if args.norm_pitch:
                            ii = pitch != 0
                            pitch[ii] -= mean[src_ids[b]]
                            pitch[ii] /= std[src_ids[b]]
                        pitches = morph_seq_len(r["units"], pitch.numpy(), new_counts[b, :int(dd_len[b])]).tolist()
                    f.write(json.dumps({"units": out_seq[b, :n].tolist(), "f0": pitches, "audio": r["audio"]}) + "\n")


def build_parser():
    """infer.py:174-193 (same flags, defaults and store_false quirk of --norm_pitch)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--input_path", default="data/VCTK/hubert100/val.txt")
    ap.add_argument("-n", default=10, type=int)
    ap.add_argument("--out_path", default="data/VCTK/pred_hubert")
    ap.add_argument("--pred_len", action="store_true")
    ap.add_argument("--pred_pitch", action="store_true")
    ap.add_argument("--len_model", default="checkpoints/vctk/len/")
    ap.add_argument("--f0_model", default="checkpoints/vctk/pitch/")
    ap.add_argument("--f0_model_type", default="new")
    ap.add_argument("--n_tokens", default=100, type=int)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--seed", default=42, type=int)
    ap.add_argument("--f0_path", default="data/VCTK/hubert100/f0_stats.pkl")
    ap.add_argument("--vc", action="store_true")
    ap.add_argument("--norm_pitch", action="store_false")
    ap.add_argument("--target_speakers", nargs="+", default=None)
    ap.add_argument("--sample_df", default=None)
    ap.add_argument("--wild_sample", action="store_true")
    ap.add_argument("--id_to_spkr", default=None)
    return ap


def main(argv: Optional[Sequence[str]] = None):
    args = build_parser().parse_args(argv)
    assert args.pred_len | args.pred_pitch, "Inference must at least convert pitch or rhythm (or both)"
    assert (args.wild_sample & args.pred_len & args.pred_pitch) | (not args.wild_sample), \
        "If we use an unknown speaker we must convert both pitch and rhythm"
    torch.manual_seed(args.seed if args.seed >= 0 else 0)
    run(args)


if __name__ == "__main__":
    main()
