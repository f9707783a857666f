"""Unit-extraction CLI -- the host side of the reference's ``data/encode.py`` on B200.

Same flags and output (one JSON line per clip: ``{"units": [...], "durations": [...], "audio": <file>}`` appended to
``--out_file``), plus ``--hubert_checkpoint`` / ``--kmeans_path`` because nothing can be downloaded by name here.
The reference encodes one file at a time (B=1, data/encode.py:27-32); here clips are read on the host, length-sorted,
zero-padded into batches with per-clip sample counts and encoded by one ``dissc_hubert_forward`` call per batch.
``f0`` (YAAPT, CPU-only amfm_decompy in textless) is not produced.
"""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np
import torch

from .hubert import SpeechEncoder


def read_wav(path: str):
    """16-bit / float PCM wav -> (fp32 mono waveform in [-1,1], sample rate), like ``torchaudio.load`` + flatten."""
    from scipy.io import wavfile
    rate, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    # textless' reader flattens whatever torchaudio returned ((channels, N) -> (1, channels*N)); mono files are unaffected
    return torch.from_numpy(np.ascontiguousarray(x.T if x.ndim == 2 else x).reshape(-1)), rate


def encode_files(encoder: SpeechEncoder, base_dir: str, files: Sequence[str], out_file: str, batch: int = 32,
                 max_samples: int = 32 * 160000) -> int:
    waves = [(f, read_wav(os.path.join(base_dir, f))[0]) for f in files]
    order = sorted(range(len(waves)), key=lambda i: -len(waves[i][1]))
    results = {}
    i = 0
    while i < len(order):
        N = len(waves[order[i]][1])
        nb = max(1, min(batch, max_samples // max(N, 1)))
        idx = order[i:i + nb]
        i += nb
        idx = [j for j in idx if len(waves[j][1]) >= 400]       # shorter than the receptive field: no frames
        if not idx:
            continue
        wave = torch.zeros(len(idx), N)
        for b, j in enumerate(idx):
            wave[b, :len(waves[j][1])] = waves[j][1]
        units, n_frames, _ = encoder.encode_batch(wave.to(encoder.device),
                                                  torch.tensor([len(waves[j][1]) for j in idx], dtype=torch.int32),
                                                  return_dense=False)
        units, n_frames = units.cpu(), n_frames.cpu()
        for b, j in enumerate(idx):
            u = units[b, :int(n_frames[b])]
            if encoder.deduplicate:
                u, d = torch.unique_consecutive(u, return_counts=True)
            else:
                d = torch.ones_like(u)
            results[j] = {"units": u.tolist(), "durations": d.tolist(), "audio": waves[j][0]}
    with open(out_file, "a+") as f:
        for j in range(len(waves)):                              # original listing order, like the reference loop
            if j in results:
                f.write(f"{json.dumps(results[j])}\n")
    return len(results)


def build_parser():
    ap = argparse.ArgumentParser(
        description="HuBERT-base layer-6 + k-means units for every wav under --base_dir (data/encode.py).  NOTE: the "
                    "reference also writes a YAAPT 'f0' per clip (CPU DSP inside textless, data/encode.py:32-38); that is "
                    "not ported, so these manifests carry 'units' / 'durations' / 'audio' only.  Vocoder configs with "
                    "\"f0\": true need an F0 contour: predict it with `python -m dissc_b200.infer --pred_pitch ...`, or "
                    "merge an externally computed 50 Hz F0 track into the manifest (INTEGRATION.md section 3).")
    ap.add_argument("--model_name", default="hubert-base-ls960")
    ap.add_argument("--quantizer_name", default="kmeans")
    ap.add_argument("--vocab_size", default=100, type=int)
    ap.add_argument("--base_dir", required=True)
    ap.add_argument("--out_file", default="ESD/hubert100/train.txt")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--hubert_checkpoint", default=None, help="fairseq hubert_base_ls960.pt (textless downloads it by name)")
    ap.add_argument("--kmeans_path", default=None, help="km.bin of the matching k-means model")
    ap.add_argument("--batch", type=int, default=32)
    return ap


def main(argv: Optional[Sequence[str]] = None):
    args = build_parser().parse_args(argv)
    encoder = SpeechEncoder.by_name(dense_model_name=args.model_name, quantizer_model_name=args.quantizer_name,
                                    vocab_size=args.vocab_size, deduplicate=False,
                                    hubert_checkpoint=args.hubert_checkpoint, kmeans_path=args.kmeans_path).to(args.device)
    os.makedirs(Path(args.out_file).parent.parent.absolute(), exist_ok=True)
    os.makedirs(Path(args.out_file).parent.absolute(), exist_ok=True)
    files: List[str] = sorted(os.listdir(args.base_dir))
    n = encode_files(encoder, args.base_dir, files, args.out_file, args.batch)
    print(f"encoded {n}/{len(files)} clips -> {args.out_file}")


if __name__ == "__main__":
    main()
