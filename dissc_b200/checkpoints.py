"""Loaders for the real checkpoints the reference pipeline downloads -- without fairseq / old sklearn.

* ``hubert_base_ls960.pt`` is a fairseq checkpoint: ``{'args'|'cfg': ..., 'model': state_dict, ...}`` whose pickle refers
  to omegaconf / fairseq / argparse classes.  Only ``['model']`` (plain tensors) is needed, so unknown classes are
  replaced by inert stubs while unpickling.
* ``km.bin`` is a joblib pickle of ``sklearn.cluster.MiniBatchKMeans``; only ``cluster_centers_`` is needed.

Neither file exists offline, so the loaders are exercised on look-alikes: a ``torch.save`` of a dict whose config
objects are instances of classes that do not exist at load time, and a REAL ``joblib.dump`` of a fitted
``sklearn.cluster.MiniBatchKMeans`` (tests/test_hubert_oracle.py).
"""
from __future__ import annotations

import pickle

import numpy as np
import torch


# Unpickling runs code: a crafted ``hubert.pt`` / ``km.bin`` could import and call anything.  Only the (module, name)
# pairs below -- what tensors, numpy arrays and plain containers need to rebuild themselves -- resolve to real objects;
# EVERY other global (fairseq / omegaconf / argparse config classes, sklearn estimators, but also ``builtins.eval``,
# ``os.system``, ``torch.load`` ...) becomes an inert stub class that merely records its state.
_TORCH_STORAGES = ("FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage", "LongStorage", "IntStorage",
                   "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage")
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch", "Size"), ("torch", "device"),
    *[("torch", n) for n in _TORCH_STORAGES], ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
    ("joblib.numpy_pickle", "NumpyArrayWrapper"), ("joblib.numpy_pickle", "NDArrayWrapper"),
    ("_codecs", "encode"), ("copyreg", "_reconstructor"), ("builtins", "object"),
    *[("builtins", n) for n in ("set", "frozenset", "list", "dict", "tuple", "int", "float", "complex", "bool", "str",
                                "bytes", "bytearray", "slice", "range")],
}


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state

    def __call__(self, *a, **k):
        return self


def _resolve(unpickler_cls, self, module, name):
    if (module, name) in _ALLOWED_GLOBALS:
        return super(unpickler_cls, self).find_class(module, name)
    return type(name, (_Stub,), {"__module__": module})


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        return _resolve(_StubUnpickler, self, module, name)


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


def _joblib_load_stubbed(path: str):
    """joblib.load with the same allowlist: joblib's NumpyUnpickler (it knows the array framing of ``joblib.dump``)
    with ``find_class`` overridden, so an sklearn estimator comes back as a stub holding its ``__dict__``."""
    from joblib import numpy_pickle as npk

    class _JoblibStubUnpickler(npk.NumpyUnpickler):
        def find_class(self, module, name):
            return _resolve(_JoblibStubUnpickler, self, module, name)

    with open(path, "rb") as f:
        with npk._validate_fileobject_and_memmap(f, path, None) as (fobj, _):   # handles joblib's compressed containers
            if isinstance(fobj, str):
                raise ValueError("pre-0.10 joblib layout (companion files) is not supported")
            try:
                return _JoblibStubUnpickler(path, fobj, ensure_native_byte_order=True).load()
            except TypeError:   # older joblib: no ensure_native_byte_order argument
                fobj.seek(0)
                return _JoblibStubUnpickler(path, fobj).load()


def load_kmeans_centers(path: str) -> torch.Tensor:
    """-> cluster_centers_ (K,D) fp32 from a joblib/pickle dump of a (MiniBatch)KMeans, or from a .npy / .pt tensor.
    No sklearn class is instantiated and no global outside ``_ALLOWED_GLOBALS`` is resolved."""
    if path.endswith(".npy"):
        return torch.from_numpy(np.load(path, allow_pickle=False)).float()
    if path.endswith(".pt"):
        return torch.as_tensor(torch.load(path, map_location="cpu", weights_only=True)).float()
    km = None
    try:
        km = _joblib_load_stubbed(path)
    except Exception:   # noqa: BLE001 -- not a joblib file (or no joblib): plain pickle
        with open(path, "rb") as f:
            km = _StubUnpickler(f).load()
    state = getattr(km, "_state", None)
    centers = getattr(km, "cluster_centers_", None)
    if centers is None and isinstance(state, dict):
        centers = state.get("cluster_centers_")
    if centers is None and isinstance(km, dict):
        centers = km.get("cluster_centers_")
    if centers is None:
        raise ValueError(f"{path}: no cluster_centers_ found")
    return torch.from_numpy(np.asarray(centers)).float()
