"""Vocoder CLI -- the host side of the reference's ``sr/inference.py`` on B200.

Same flags, ``config.json`` + newest ``g_*`` checkpoint discovery, manifest format and output names
(``<stem>_gen.wav`` resynthesis, ``<stem>_<spkr_id>_gen.wav`` voice conversion; float32 wav, peak-normalised,
``sampling_rate`` Hz).  What changes is the execution model: the reference spawns ``Pool(8)`` workers that each run
B=1 forwards (sr/inference.py:288-292,351-359); here there is one process per GPU (``torchrun --nproc-per-node N -m
dissc_b200.inference ...`` or a single process), utterances are length-sorted, dealt to ranks, and vocoded in padded
batches with per-utterance ``lengths`` -- each row equals the reference's B=1 output for that utterance.

The input side keeps only what feeds the Generator (sr/dataset.py:107-122,291-312): units, F0 (speaker-normalised on
voiced frames iff ``f0_normalize``) and the speaker id.  The reference additionally computes a mel-spectrogram per
item that inference never uses (sr/dataset.py:269-271) -- not done here.  The ground-truth wav is read only to write
the ``<stem>_gt.wav`` copy (and for the ``--eval_mode`` length clipping); a missing / unreadable file is skipped
instead of aborting the run.  ``--f0-stats``, ``--unseen-f0``, ``--sample_df`` and ``--code_file`` follow
sr/inference.py:97-100,122-129,146-147,214-235.
"""
from __future__ import annotations

import argparse
import ast
import glob
import json
import os
import pickle
import random
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .models import AttrDict, CodeGenerator


def scan_checkpoint(cp_dir: str, prefix: str) -> str:
    """sr/inference.py:59-64: newest ``<prefix>*`` by sorted name ('' if none)."""
    cp_list = glob.glob(os.path.join(cp_dir, prefix + "*"))
    return sorted(cp_list)[-1] if cp_list else ""


def load_config(checkpoint_file: str) -> AttrDict:
    """sr/inference.py:105-112: config.json next to the checkpoint (file) or inside it (dir)."""
    d = checkpoint_file if os.path.isdir(checkpoint_file) else os.path.split(checkpoint_file)[0]
    with open(os.path.join(d, "config.json")) as f:
        return AttrDict(json.loads(f.read()))


def parse_manifest(manifest: str, base_path: str):
    """sr/dataset.py:107-122 (dict lines are parsed with json / literal_eval instead of eval)."""
    audio_files, codes, pitch = [], [], []
    with open(manifest) as info:
        for line in info.readlines():
            if line[:1] == "{":
                try:
                    sample = json.loads(line.strip())
                except json.JSONDecodeError:
                    sample = ast.literal_eval(line.strip())
                codes.append(np.asarray(sample["units"], dtype=np.int64))
                audio_files.append(Path(str(base_path) + "/" + sample["audio"].split("/")[-1]))
                if "f0" in sample:
                    pitch.append(np.array(sample["f0"]))
            elif line.strip():
                audio_files.append(Path(line.strip()))
    return audio_files, codes, pitch


def parse_speaker(path, method) -> str:
    """sr/dataset.py:132-146."""
    path = Path(path)
    if method == "parent_name":
        return path.parent.name
    if method == "parent_parent_name":
        return path.parent.parent.name
    if method == "_":
        return path.name.split("_")[0]
    if method == "single":
        return "A"
    if callable(method):
        return method(path)
    raise NotImplementedError(method)


def normalize_f0(f0: np.ndarray, mean: float, std: float, f0_median: bool = False) -> np.ndarray:
    """sr/dataset.py:297-312: voiced frames (f0 != 0) -> (f0 - mean) / std; unvoiced frames stay 0, or with
    ``f0_median`` are filled with the utterance's voiced median first and normalised like the rest (:306-309)."""
    f0 = f0.astype(np.float32).copy()
    ii = f0 != 0
    if f0_median:
        med = np.median(f0[ii])          # NaN (with numpy's warning) for an all-unvoiced utterance, as in the reference
        f0[~ii] = med
        f0[~ii] = (f0[~ii] - mean) / std
    f0[ii] = (f0[ii] - mean) / std
    return f0


def prepare_items(h, audio_files, codes, pitch, id_to_spkr: Sequence[str], f0_stats: Optional[dict],
                  unseen_speaker: bool = False) -> List[dict]:
    """The slice of CodeDataset.__getitem__ (sr/dataset.py:221-317) that feeds the Generator at inference."""
    spkr_to_id = {k: v for v, k in enumerate(id_to_spkr)}
    items = []
    for i, (path, code) in enumerate(zip(audio_files, codes)):
        it = {"name": str(path), "code": np.asarray(code, dtype=np.int64)}
        if h.get("f0", None):
            if i >= len(pitch) or len(pitch[i]) == 0:
                raise NotImplementedError(f"{path}: manifest line has no 'f0' (YAAPT extraction is CPU-only, not ported)")
            f0 = np.asarray(pitch[i], dtype=np.float32)
            if h.get("f0_normalize", False):
                name = parse_speaker(path, h.get("multispkr", None))
                st = f0_stats if name not in f0_stats else f0_stats[name]
                mean, std = (st["f0_mean"], st["f0_std"]) if name not in f0_stats else (st["mean"], st["std"])
                f0 = normalize_f0(f0, mean, std, bool(h.get("f0_median", False)))
                if h.get("f0_feats", False):           # sr/dataset.py:314-315: the source speaker's [mean, std]
                    it["f0_stats"] = np.asarray([mean, std], dtype=np.float32)
            it["f0"] = f0
            n = min(len(it["code"]), len(f0))
            if len(it["code"]) != len(f0):
                if max(len(it["code"]), len(f0)) % n:
                    raise NotImplementedError("Padding condition signal - misalignment between condition features.")
        if h.get("multispkr", None):
            it["spkr"] = 0 if unseen_speaker else spkr_to_id[parse_speaker(path, h.get("multispkr"))]
        items.append(it)
    return items


def peak_normalize(audio_i16: np.ndarray) -> np.ndarray:
    """``librosa.util.normalize(audio.astype(np.float32))`` (sr/inference.py:250): divide by max |x| (inf-norm);
    all-zero (or tiny, < float32 tiny) signals are returned unchanged."""
    x = audio_i16.astype(np.float32)
    peak = np.max(np.abs(x)) if x.size else 0.0
    if peak < np.finfo(np.float32).tiny:
        return x
    return x / peak


def peak_normalize_f32(x: np.ndarray) -> np.ndarray:
    """``librosa.util.normalize`` on a float signal (the ground-truth copy, sr/inference.py:255)."""
    x = np.asarray(x, dtype=np.float32)
    peak = np.max(np.abs(x)) if x.size else 0.0
    return x if peak < np.finfo(np.float32).tiny else x / peak


def write_wav(path: str, rate: int, audio: np.ndarray) -> None:
    from scipy.io.wavfile import write
    write(path, rate, audio)


class WavWriter:
    """Peak-normalise + write wavs on a few background threads so the disk never stalls the next GPU batch (the reference
    does both inline, once per utterance and target speaker: sr/inference.py:250-251).  ``close()`` waits for every file
    and re-raises the first error."""

    def __init__(self, workers: int = 4):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._futs = []

    @staticmethod
    def _job(path, rate, audio, is_int16):
        write_wav(path, rate, peak_normalize(audio) if is_int16 else peak_normalize_f32(audio))

    def submit(self, path: str, rate: int, audio: np.ndarray) -> None:
        self._futs.append(self._pool.submit(self._job, path, rate, audio, audio.dtype == np.int16))

    def close(self) -> None:
        try:
            for f in self._futs:
                f.result()
        finally:
            self._pool.shutdown(wait=True)
            self._futs = []


def batches_by_length(lengths: Sequence[int], max_batch: int, max_frames: int) -> List[List[int]]:
    """Greedy length-sorted batching: indices sorted by length (desc), cut when the padded batch would exceed
    ``max_batch`` rows or ``max_frames`` padded frames."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    out, cur = [], []
    for i in order:
        L = int(lengths[cur[0]]) if cur else int(lengths[i])
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * L > max_frames):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


@torch.no_grad()
def vocode_items(generator: CodeGenerator, items: List[dict], device, spkr_override: Optional[int] = None,
                 max_batch: int = 64, max_frames: int = 64 * 400,
                 f0_stats_override: Optional[Sequence[float]] = None) -> Dict[int, np.ndarray]:
    """-> {item index: int16 waveform (hop * n_frames,)}, identical per item to ``generate()`` (sr/inference.py:67-76).
    Items carrying ``f0_stats`` (``f0_feats`` configs) pass it on as the extra conditioning feature; ``f0_stats_override``
    replaces it with the TARGET speaker's [mean, std] for voice conversion (sr/inference.py:237-245)."""
    hop = generator.hop
    out: Dict[int, np.ndarray] = {}
    n_frames = []
    for it in items:
        T = len(it["code"])
        if "f0" in it and len(it["f0"]) > T:
            T = len(it["f0"])
        n_frames.append(T)
    for idx in batches_by_length(n_frames, max_batch, max_frames):
        T = n_frames[idx[0]]
        B = len(idx)
        code = torch.zeros((B, T), dtype=torch.int64)
        f0 = torch.zeros((B, T), dtype=torch.float32)
        spkr = torch.zeros((B, 1), dtype=torch.int64)
        lengths = torch.zeros((B,), dtype=torch.int32)
        feats = {}
        if any("f0_stats" in items[i] for i in idx) or f0_stats_override is not None:
            st = [np.asarray(f0_stats_override if f0_stats_override is not None else items[i]["f0_stats"], dtype=np.float32)
                  for i in idx]
            feats["f0_stats"] = torch.from_numpy(np.stack(st)).to(device)
        for b, i in enumerate(idx):
            it = items[i]
            n = n_frames[i]
            c = torch.from_numpy(it["code"])
            if len(c) != n:                      # _upsample: nearest-repeat the shorter signal (sr/models.py:158-177)
                c = c.repeat_interleave(n // len(c))
            code[b, :n] = c
            if "f0" in it:
                f = torch.from_numpy(np.asarray(it["f0"], dtype=np.float32))
                if len(f) != n:
                    f = f.repeat_interleave(n // len(f))
                f0[b, :n] = f
            spkr[b, 0] = it.get("spkr", 0) if spkr_override is None else spkr_override
            lengths[b] = n
        y = generator.generate_int16(code.to(device), f0.to(device) if generator.f0 else None,
                                     spkr.to(device) if generator.multispkr else None, lengths=lengths.to(device), **feats)
        y = y.cpu().numpy()
        generator.check_indices(synchronize=False)   # the .cpu() above synchronised: bad unit / speaker ids raise here
        for b, i in enumerate(idx):
            out[i] = y[b, :hop * n_frames[i]].copy()
    return out


def rescale_f0(f0: np.ndarray, new_mean: float, new_std: float) -> np.ndarray:
    """``--f0-stats`` re-scaling of sr/inference.py:220-235: voiced frames are standardised with their OWN mean / std
    (``torch.std``: unbiased) and moved to the target speaker's statistics; unvoiced frames stay 0."""
    f0 = torch.from_numpy(np.asarray(f0, dtype=np.float32)).clone()
    ii = f0 != 0
    mean_, std_ = f0[ii].mean(), f0[ii].std()
    f0[ii] -= mean_
    f0[ii] /= std_
    f0[ii] *= new_std
    f0[ii] += new_mean
    return f0.numpy()


def target_f0_stats(f0_stats: dict, spkr: int):
    """sr/inference.py:227-230: per-speaker entry of the ``--f0-stats`` dict, else its global ``f0_mean`` / ``f0_std``."""
    if spkr not in f0_stats:
        return float(f0_stats["f0_mean"]), float(f0_stats["f0_std"])
    return float(f0_stats[spkr]["f0_mean"]), float(f0_stats[spkr]["f0_std"])


def parse_code_file(path: str) -> List[dict]:
    """``--code_file`` (sr/inference.py:122-129): one ``<name>|<c0 c1 c2 ...>`` line per utterance, units only."""
    items = []
    with open(path) as f:
        for line in f.readlines():
            x = line.strip().split("|")
            if len(x) < 2:
                continue
            items.append({"name": x[0], "code": np.asarray([int(v) for v in x[1].split(" ") if v], dtype=np.int64)})
    return items


def load_gt_audio(path, sampling_rate: int, pad: Optional[int] = None) -> Optional[np.ndarray]:
    """Ground-truth audio the way CodeDataset.__getitem__ prepares it (sr/dataset.py:223-234): int16 wav -> optional
    zero padding to a multiple of ``pad`` -> / 32768 -> peak-normalised * 0.95.  None when the file is missing, is not
    a PCM wav scipy can read, or has another sampling rate (the reference resamples with resampy, absent here)."""
    from scipy.io import wavfile
    try:
        rate, audio = wavfile.read(str(path))
    except (OSError, ValueError):
        return None
    if rate != sampling_rate:
        return None
    if audio.ndim > 1:
        audio = audio[:, 0]
    audio = audio.astype(np.float64)
    if pad:
        audio = np.pad(audio, (0, pad - (audio.shape[-1] % pad)), "constant", constant_values=0)
    audio = audio / 32768.0
    peak = np.max(np.abs(audio)) if audio.size else 0.0
    if peak >= np.finfo(np.float32).tiny:
        audio = audio / peak
    return (audio * 0.95).astype(np.float32)


def sample_df_targets(df, out_name: str, spkr_to_id: dict) -> List[int]:
    """sr/inference.py:214-216: the conversions listed for this sample in the ``--sample_df`` table."""
    cur_name = out_name.split("_mic2")[0]
    return [spkr_to_id[i] for i in df[df.syn_sample == cur_name].syn_trgt.unique()]


def select_items(n_items: int, n: int, debug: bool) -> List[int]:
    """Which manifest entries get processed, and in which order (sr/inference.py:340-359).

    * ``--debug``: manifest order; the loop breaks once ``i > n``, i.e. entries ``0 .. n+1`` (n + 2 of them).
    * otherwise: ``random.shuffle`` of all indices, then ``n + 1`` results are awaited from ``pool.imap`` (its counter
      starts at 1).  ``main`` seeds the global RNG with 52 (:283-286) but ``CodeDataset.__init__`` re-seeds it with 1234
      (sr/dataset.py:157) before the shuffle (:351), so the permutation is that of ``random.Random(1234)``.
      (The reference's pool keeps working on a few more entries in the background until it is torn down; which extra
      files appear is a race there and is not reproduced.)"""
    idx = list(range(n_items))
    if debug:
        return idx if n == -1 else idx[: n + 2]
    random.Random(1234).shuffle(idx)
    return idx if n == -1 else idx[: n + 1]


def vc_targets_for_item(item_index: int, n_speakers: int) -> List[int]:
    """``--vc`` without ``--target-speakers``: five random target speakers PER UTTERANCE (sr/inference.py:210-211).
    The reference draws them from its worker's RNG stream (seed 52 + GPU ordinal, :165-169), so which speakers an
    utterance gets depends on how the pool happened to schedule it; here the draw is seeded by the utterance's
    manifest index alone, hence reproducible and independent of rank / world size."""
    return random.Random(52 * 1000003 + item_index).sample(range(n_speakers), k=min(5, n_speakers))


def build_parser():
    """sr/inference.py:263-281."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--code_file", default=None)
    ap.add_argument("--input_code_file", default="./datasets/LJSpeech/cpc100/test.txt")
    ap.add_argument("--data_path", default=None)
    ap.add_argument("--output_dir", default="generated_files")
    ap.add_argument("--checkpoint_file", required=True)
    ap.add_argument("--f0-stats", type=Path)
    ap.add_argument("--vc", action="store_true")
    ap.add_argument("--target-speakers", type=str, nargs="+", default=None)
    ap.add_argument("--pad", default=None, type=int)
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--eval_mode", action="store_false")
    ap.add_argument("--parts", action="store_true")
    ap.add_argument("--unseen-f0", type=Path)
    ap.add_argument("--unseen_speaker", action="store_true")
    ap.add_argument("--id_to_spkr", default=None)
    ap.add_argument("--sample_df", default=None)
    ap.add_argument("-n", type=int, default=2508)
    ap.add_argument("--batch", type=int, default=64, help="utterances per forward (extension)")
    return ap


def _torch_load(path):
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except TypeError:  # torch without the weights_only keyword
        return torch.load(path, map_location="cpu")


def main(argv: Optional[Sequence[str]] = None):
    a = build_parser().parse_args(argv)
    from . import dist as ddist
    rank, world, local = ddist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("dissc_b200.inference needs a CUDA device (sm_100a); there is no CPU path")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    h = load_config(a.checkpoint_file)
    cp_g = scan_checkpoint(a.checkpoint_file, "g_") if os.path.isdir(a.checkpoint_file) else a.checkpoint_file
    if not cp_g or not os.path.isfile(cp_g):
        print(f"Didn't find checkpoints for {cp_g}")   # sr/inference.py:307-309
        return
    generator = CodeGenerator(h).to(device)
    generator.load_state_dict(_torch_load(cp_g)["generator"])
    generator.eval()
    generator.remove_weight_norm()
    os.makedirs(a.output_dir, exist_ok=True)

    df = None
    if a.sample_df:                                            # sr/inference.py:97-100
        import pandas as pd
        df = pd.read_csv(a.sample_df, index_col=0)
        if a.target_speakers:
            df = df[df.syn_trgt.isin(a.target_speakers)]

    id_to_spkr: Sequence[str] = []
    gt_paths: List[Optional[Path]] = []
    if a.code_file is not None:                                # units only: no F0 / speaker conditioning available
        if h.get("f0", None) or h.get("multispkr", None):
            raise NotImplementedError("--code_file carries units only; this checkpoint's config also conditions on "
                                      "f0 / speaker (the reference crashes on this combination too: parse_code "
                                      "returns a list, sr/inference.py:125-129,178)")
        items = parse_code_file(a.code_file)
        gt_paths = [None] * len(items)
    else:
        base_path = h.test_base_path if a.data_path is None else a.data_path
        audio_files, codes, pitch = parse_manifest(a.input_code_file, base_path)
        spk_pkl = a.id_to_spkr if a.unseen_speaker else f"{os.path.dirname(h.input_training_file)}/id_to_spkr.pkl"
        with open(spk_pkl, "rb") as f:
            id_to_spkr = pickle.load(f)
        f0_stats = None
        if a.unseen_f0:                                        # sr/inference.py:146-147
            f0_stats = _torch_load(a.unseen_f0)
        elif h.get("f0_stats", None):
            with open(h.f0_stats, "rb") as f:
                f0_stats = pickle.load(f)
        items = prepare_items(h, audio_files, codes, pitch, id_to_spkr, f0_stats, a.unseen_speaker)
        gt_paths = list(audio_files)
    chosen = select_items(len(items), a.n, a.debug)            # sr/inference.py:340-359 (seeded shuffle, -n)
    for pos, it in enumerate(items):
        it["manifest_index"] = pos
    items = [items[i] for i in chosen]
    gt_paths = [gt_paths[i] for i in chosen]
    mine = ddist.shard_by_length([len(it["code"]) for it in items], world)[rank]
    my_items = [items[i] for i in mine]
    my_gt = [gt_paths[i] for i in mine]
    spkr_to_id = {k: v for v, k in enumerate(id_to_spkr)}
    hop = int(h.get("code_hop_size", generator.hop))

    def out_name(it):
        p = Path(it["name"])
        return "_".join(p.parts[-3:])[:-4] if a.parts else p.stem

    # ground truth copies (sr/inference.py:253-256); with --eval_mode (store_false -> not eval) codes / f0 are clipped to
    # the audio length first (sr/dataset.py:244-252)
    write_gt = df is None and a.code_file is None
    gts: List[Optional[np.ndarray]] = [None] * len(my_items)
    if write_gt or not a.eval_mode:
        for j, pth in enumerate(my_gt):
            gts[j] = load_gt_audio(pth, h.sampling_rate, a.pad) if pth is not None else None
    if not a.eval_mode:
        for it, gt in zip(my_items, gts):
            if gt is None:
                continue
            n = min(gt.shape[0] // hop, len(it["code"]))
            it["code"] = it["code"][:n]
            if "f0" in it:
                it["f0"] = it["f0"][:n]

    writer = WavWriter()
    if df is None and not a.unseen_speaker:                    # resynthesis (sr/inference.py:203-207)
        for i, audio in vocode_items(generator, my_items, device, None, a.batch).items():
            writer.submit(os.path.join(a.output_dir, out_name(my_items[i]) + "_gen.wav"), h.sampling_rate, audio)
    if h.get("multispkr", None) and a.vc:                      # voice conversion (sr/inference.py:209-251)
        if a.target_speakers is not None:
            spkrs = [spkr_to_id[s] for s in a.target_speakers]
            per_item = [list(spkrs)] * len(my_items)
        else:
            per_item = [vc_targets_for_item(it["manifest_index"], len(id_to_spkr)) for it in my_items]
        if df is not None:
            per_item = [sample_df_targets(df, out_name(it), spkr_to_id) for it in my_items]
        f0_tgt = None
        if a.f0_stats and h.get("f0", None) is not None:
            f0_tgt = _torch_load(a.f0_stats)                   # sr/inference.py:157-158
        rescale = f0_tgt is not None and not h.get("f0_normalize", False)
        feats_stats = None
        order: List[int] = []
        for ks in per_item:
            for k in ks:
                if k not in order:
                    order.append(k)
        for k in order:
            sel = [j for j, ks in enumerate(per_item) if k in ks]
            if rescale:
                # like the reference, the re-scaled contour REPLACES the item's f0 (code['f0'] = f0, :235), so the next
                # target speaker standardises the already re-scaled contour
                for j in sel:
                    if (my_items[j]["f0"] != 0).any():
                        my_items[j]["f0"] = rescale_f0(my_items[j]["f0"], *target_f0_stats(f0_tgt, k))
            sub = [my_items[j] for j in sel]
            tgt_stats = None
            if h.get("f0_feats", False):                       # sr/inference.py:237-245: the TARGET speaker's [mean, std]
                if feats_stats is None:
                    feats_stats = _torch_load(h["f0_stats"])
                tgt_stats = list(target_f0_stats(feats_stats, k))
            for i, audio in vocode_items(generator, sub, device, k, a.batch, f0_stats_override=tgt_stats).items():
                writer.submit(os.path.join(a.output_dir, out_name(sub[i]) + f"_{k}_gen.wav"), h.sampling_rate, audio)
    if write_gt:
        for it, gt in zip(my_items, gts):
            if gt is not None:
                writer.submit(os.path.join(a.output_dir, out_name(it) + "_gt.wav"), h.sampling_rate, gt)
    writer.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
