"""ctypes wrapper over oracle/c/dissc_oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdissc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "dissc_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def conv1d(x, w, b, stride=1, dilation=1, padding=0, groups=1):
    x, xp = _f(x); w, wp = _f(w)
    B, Cin, T = x.shape
    Cout, _, k = w.shape
    Tout = (T + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    y = np.empty((B, Cout, Tout), np.float32)
    if b is not None:
        b, bp = _f(b)
    else:
        bp = None
    lib().oracle_conv1d(xp, wp, bp, y.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, Cin, Cout, T, k,
                        stride, dilation, padding, groups)
    return y


def conv_transpose1d(x, w, b, stride, padding):
    x, xp = _f(x); w, wp = _f(w); b, bp = _f(b)
    B, Cin, T = x.shape
    _, Cout, k = w.shape
    Tout = (T - 1) * stride - 2 * padding + k
    y = np.empty((B, Cout, Tout), np.float32)
    lib().oracle_conv_transpose1d(xp, wp, bp, y.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, Cin, Cout, T, k,
                                  stride, padding)
    return y


def generator_forward(folded: dict, h: dict, code, f0, spkr):
    """folded: {name: np.ndarray} with weight-norm removed (oracle.generator_oracle.folded_state_dict)."""
    assert h["resblock"] == "1"
    code = np.ascontiguousarray(code, np.int64)
    B, T = code.shape
    E = folded["dict.weight"].shape[1]
    f0a, f0p = _f(np.asarray(f0).reshape(B, T))
    spk = np.ascontiguousarray(np.asarray(spkr).reshape(B), np.int64)
    dw, dwp = _f(folded["dict.weight"]); sw, swp = _f(folded["spkr.weight"])
    x = np.empty((B, 2 * E + 1, T), np.float32)
    i64 = ctypes.POINTER(ctypes.c_int64)
    fp = ctypes.POINTER(ctypes.c_float)
    lib().oracle_build_input(code.ctypes.data_as(i64), f0p, spk.ctypes.data_as(i64), dwp, swp,
                             x.ctypes.data_as(fp), B, T, E)
    names = ["conv_pre"]
    nk = len(h["resblock_kernel_sizes"])
    nd = len(h["resblock_dilation_sizes"][0])
    for i in range(len(h["upsample_rates"])):
        names.append(f"ups.{i}")
        for j in range(nk):
            for m in range(nd):
                names += [f"resblocks.{i * nk + j}.convs1.{m}", f"resblocks.{i * nk + j}.convs2.{m}"]
    names.append("conv_post")
    keep, ptrs = [], []
    for n in names:
        for s in (".weight", ".bias"):
            a, p = _f(folded[n + s]); keep.append(a); ptrs.append(p)
    W = (fp * len(ptrs))(*ptrs)
    rates = (ctypes.c_int * len(h["upsample_rates"]))(*h["upsample_rates"])
    ks = (ctypes.c_int * len(h["upsample_kernel_sizes"]))(*h["upsample_kernel_sizes"])
    rks = (ctypes.c_int * nk)(*h["resblock_kernel_sizes"])
    dils = (ctypes.c_int * (nk * nd))(*[d for row in h["resblock_dilation_sizes"] for d in row])
    up = int(np.prod(h["upsample_rates"]))
    y = np.empty((B, 1, T * up), np.float32)
    rc = lib().oracle_generator_forward(x.ctypes.data_as(fp), y.ctypes.data_as(fp), B, 2 * E + 1, T,
                                        h["upsample_initial_channel"], len(h["upsample_rates"]), rates, ks,
                                        nk, rks, nd, dils, W)
    assert rc == 0
    return y


def kmeans_assign(x, c):
    x, xp = _f(x); c, cp = _f(c)
    out = np.empty(x.shape[0], np.int64)
    lib().oracle_kmeans_assign(xp, cp, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), x.shape[0], x.shape[1],
                               c.shape[0])
    return out


def len_carryover(lens):
    lens, lp = _f(np.asarray(lens).reshape(-1))
    out = np.empty(lens.shape[0], np.int32)
    lib().oracle_len_carryover(lp, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), lens.shape[0])
    return out
