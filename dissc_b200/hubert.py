"""Unit encoder: HuBERT-base layer-6 features -> k-means units, the host side of what ``data/encode.py`` calls.

``data/encode.py:21-22,32`` does ``SpeechEncoder.by_name(dense_model_name, quantizer_model_name, vocab_size,
deduplicate).to(device)`` and then ``encoder(waveform)`` per file (B=1), keeping ``units`` / ``durations`` (and the
CPU YAAPT ``f0``, which is not part of this GPU path).  ``SpeechEncoder`` below keeps that call surface; the arithmetic is
``dissc_hubert_forward`` in ``libdissc_b200.so`` (csrc/hubert.cu).  There is no CPU fallback.

textless downloads ``hubert_base_ls960.pt`` and ``km.bin`` by name; there is no network here, so ``by_name`` takes the
two files explicitly (``hubert_checkpoint=``, ``kmeans_path=``) or the weights directly (``from_state_dict``).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib

BASE_CFG = dict(n_layers=6, embed_dim=768, ffn_dim=3072, n_heads=12, conv_dim=512, pos_kernel=128, pos_groups=16)


def _fold_pos_conv(sd: dict) -> dict:
    """fairseq checkpoints store encoder.pos_conv.0.weight_g (1,1,k) / weight_v; fold w = g * v / ||v||_{dims 0,1}."""
    if "encoder.pos_conv.0.weight" in sd:
        return sd
    sd = dict(sd)
    v, g = sd.pop("encoder.pos_conv.0.weight_v").float(), sd.pop("encoder.pos_conv.0.weight_g").float()
    sd["encoder.pos_conv.0.weight"] = v * (g / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())
    return sd


class SpeechEncoder:
    """Drop-in for the way data/encode.py uses ``textless.data.speech_encoder.SpeechEncoder``."""

    def __init__(self, state_dict: dict, cluster_centers: torch.Tensor, deduplicate: bool = False, n_layers: int = 6,
                 **cfg):
        self.cfg = dict(BASE_CFG, n_layers=n_layers, **cfg)
        sd = _fold_pos_conv({k: v for k, v in state_dict.items() if torch.is_tensor(v) and v.is_floating_point()})
        keep_layers = {f"encoder.layers.{l}." for l in range(self.cfg["n_layers"])}
        self._sd = {k: v.detach().float().cpu().contiguous() for k, v in sd.items()
                    if not k.startswith("encoder.layers.") or any(k.startswith(p) for p in keep_layers)}
        self._sd["kmeans.cluster_centers"] = torch.as_tensor(cluster_centers).detach().float().cpu().contiguous()
        self.cfg["n_clusters"] = int(self._sd["kmeans.cluster_centers"].shape[0])
        self.deduplicate = deduplicate
        self.device: Optional[torch.device] = None
        self._handle = None
        self._ws = None

    # ---- construction -----------------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, state_dict, cluster_centers, deduplicate=False, **kw):
        return cls(state_dict, cluster_centers, deduplicate, **kw)

    @classmethod
    def by_name(cls, dense_model_name: str, quantizer_model_name: str, vocab_size: int, deduplicate: bool,
                hubert_checkpoint: Optional[str] = None, kmeans_path: Optional[str] = None, **kw):
        """data/encode.py:21-22.  Only 'hubert-base-ls960' + 'kmeans' are implemented."""
        if dense_model_name != "hubert-base-ls960" or quantizer_model_name != "kmeans":
            raise NotImplementedError(f"only hubert-base-ls960 + kmeans are implemented, not {dense_model_name}/{quantizer_model_name}")
        if hubert_checkpoint is None or kmeans_path is None:
            raise FileNotFoundError("no network: pass hubert_checkpoint= (fairseq hubert_base_ls960.pt) and kmeans_path= "
                                    "(km.bin) explicitly; textless would download them")
        from .checkpoints import load_fairseq_hubert, load_kmeans_centers
        centers = load_kmeans_centers(kmeans_path)
        if centers.shape[0] != vocab_size:
            raise ValueError(f"{kmeans_path} has {centers.shape[0]} clusters, vocab_size={vocab_size}")
        return cls(load_fairseq_hubert(hubert_checkpoint), centers, deduplicate, **kw)

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DisscError("dissc_b200.SpeechEncoder runs only on CUDA (sm_100a); there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self.device != device:
            self._drop()
        self.device = device
        return self

    def cuda(self):
        return self.to("cuda")

    def _drop(self):
        if self._handle is not None:
            _lib.lib().dissc_hubert_destroy(self._handle)
        self._handle, self._ws = None, None

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass

    def _ensure(self):
        if self.device is None:
            raise _lib.DisscError("call .to(device) first (data/encode.py:22)")
        if self._handle is not None:
            return self._handle
        L = _lib.lib()
        arr = (_lib.Tensor * len(self._sd))()
        keep = []
        for i, (k, v) in enumerate(self._sd.items()):
            keep.append((k.encode(), v))
            arr[i].name = keep[-1][0]
            arr[i].data = ctypes.cast(v.data_ptr(), ctypes.POINTER(ctypes.c_float))
            arr[i].numel = v.numel()
        cfg = _lib.HubertCfg(**{k: int(self.cfg[k]) for k, _ in _lib.HubertCfg._fields_})
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.dissc_hubert_create(ctypes.byref(h), ctypes.byref(cfg), arr, len(self._sd), self.device.index),
                       "dissc_hubert_create")
        self._handle = h
        return h

    # ---- inference ----------------------------------------------------------------------------------------
    @staticmethod
    def num_frames(n_samples: int) -> int:
        return int(_lib.lib().dissc_hubert_num_frames(int(n_samples)))

    @torch.no_grad()
    def encode_batch(self, wave: torch.Tensor, n_samples: Optional[torch.Tensor] = None, return_dense: bool = True):
        """wave fp32 (B,N) on the device (rows zero-padded), n_samples int32 (B) -> (units int64 (B,T) with -1 padding,
        n_frames int32 (B), dense fp32 (B,T,D) or None)."""
        h = self._ensure()
        dev = self.device
        wave = wave.to(device=dev, dtype=torch.float32).contiguous()
        B, N = wave.shape
        T = self.num_frames(N)
        if T <= 0:
            raise ValueError(f"clips of {N} samples are shorter than HuBERT's 400-sample receptive field")
        L = _lib.lib()
        need = ctypes.c_size_t()
        _lib.check(L.dissc_hubert_workspace_bytes(h, B, N, ctypes.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        units = torch.empty((B, T), dtype=torch.int64, device=dev)
        n_frames = torch.empty((B,), dtype=torch.int32, device=dev)
        dense = torch.empty((B, T, self.cfg["embed_dim"]), dtype=torch.float32, device=dev) if return_dense else None
        if n_samples is not None:
            n_samples = n_samples.to(device=dev, dtype=torch.int32).contiguous()
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(L.dissc_hubert_forward(h, p(wave), p(n_samples), B, N, p(units), p(n_frames), p(dense),
                                              p(self._ws), self._ws.numel(),
                                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "dissc_hubert_forward")
        return units, n_frames, dense

    def __call__(self, waveform: torch.Tensor) -> dict:
        """``encoder(waveform)`` (data/encode.py:32): waveform (channels, N) or (N,) -> dict of tensors
        {'units', 'durations', 'dense'} for ONE clip (the reader flattens to (1, N))."""
        wave = waveform.reshape(1, -1)
        units, n_frames, dense = self.encode_batch(wave)
        n = int(n_frames[0].item())
        units, dense = units[0, :n], dense[0, :n]
        if self.deduplicate:
            units, durations = torch.unique_consecutive(units, return_counts=True)
        else:
            durations = torch.ones_like(units)
        return {"units": units, "durations": durations, "dense": dense}


def kmeans_assign(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """KMeansQuantizer.forward: x fp32 (M,D) on CUDA, centroids (K,D) -> int64 (M) nearest centroid (first index on ties)."""
    if not x.is_cuda:
        raise _lib.DisscError("kmeans_assign runs only on CUDA")
    x = x.float().contiguous()
    c = centroids.to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty((x.shape[0],), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dissc_kmeans_assign(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(c.data_ptr()), x.shape[0],
                                                  x.shape[1], c.shape[0], ctypes.c_void_p(out.data_ptr()),
                                                  ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "dissc_kmeans_assign")
    return out
