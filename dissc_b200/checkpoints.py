"""Loaders for the real checkpoints the reference pipeline downloads -- without fairseq / old sklearn.

* ``hubert_base_ls960.pt`` is a fairseq checkpoint: ``{'args'|'cfg': ..., 'model': state_dict, ...}`` whose pickle refers
  to omegaconf / fairseq / argparse classes.  Only ``['model']`` (plain tensors) is needed, so unknown classes are
  replaced by inert stubs while unpickling.
* ``km.bin`` is a joblib pickle of ``sklearn.cluster.MiniBatchKMeans``; only ``cluster_centers_`` is needed.

Neither file exists offline, so these loaders are exercised only on synthetic look-alikes (tests/test_cli_host.py).
"""
from __future__ import annotations

import pickle

import numpy as np
import torch


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state

    def __call__(self, *a, **k):
        return self


class _StubUnpickler(pickle.Unpickler):
    _SAFE_PREFIXES = ("torch", "collections", "numpy", "builtins", "_codecs", "copyreg")

    def find_class(self, module, name):
        if module.split(".")[0] in self._SAFE_PREFIXES:
            return super().find_class(module, name)
        return type(name, (_Stub,), {})


class _StubPickleModule:
    Unpickler = _StubUnpickler
    __name__ = "pickle"

    @staticmethod
    def load(f, **kw):
        return _StubUnpickler(f, **kw).load()


def load_fairseq_hubert(path: str) -> dict:
    """-> fairseq HuBERT state dict (fp32 tensors), ignoring config objects that would need fairseq/omegaconf."""
    ck = torch.load(path, map_location="cpu", pickle_module=_StubPickleModule, weights_only=False)
    sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    return {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}


def load_kmeans_centers(path: str) -> torch.Tensor:
    """-> cluster_centers_ (K,D) fp32 from a joblib/pickle dump of a (MiniBatch)KMeans, or from a .npy / .pt tensor."""
    if path.endswith(".npy"):
        return torch.from_numpy(np.load(path)).float()
    if path.endswith(".pt"):
        return torch.as_tensor(torch.load(path, map_location="cpu")).float()
    try:
        import joblib
        km = joblib.load(path)
    except Exception:
        with open(path, "rb") as f:
            km = _StubUnpickler(f).load()
    centers = getattr(km, "cluster_centers_", None)
    if centers is None and hasattr(km, "_state"):
        centers = km._state.get("cluster_centers_")
    if centers is None:
        raise ValueError(f"{path}: no cluster_centers_ found")
    return torch.from_numpy(np.asarray(centers)).float()
