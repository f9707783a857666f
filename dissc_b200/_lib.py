"""ctypes binding of libdissc_b200.so (include/dissc_b200.h).

There is no fallback: if the shared library is missing or a call fails, the
product path raises.  The library is built in-tree by ``dissc_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# DISSC_LIB selects an A/B build of the same sources (dissc_b200/build.py::build_variant); default: the in-tree library
LIB_PATH = os.environ.get("DISSC_LIB") or os.path.join(HERE, "libdissc_b200.so")

MAX_STAGES, MAX_KERNELS, MAX_DILATIONS = 8, 8, 8


class DisscError(RuntimeError):
    pass


class GenCfg(ctypes.Structure):
    _fields_ = [
        ("n_up", c_int), ("up_rates", c_int * MAX_STAGES), ("up_kernels", c_int * MAX_STAGES),
        ("n_rk", c_int), ("rk", c_int * MAX_KERNELS),
        ("n_dil", c_int), ("dil", (c_int * MAX_DILATIONS) * MAX_KERNELS),
        ("c0", c_int), ("embedding_dim", c_int), ("num_embeddings", c_int), ("n_spkr_rows", c_int),
        ("model_in_dim", c_int), ("resblock", c_int), ("has_f0", c_int), ("has_spkr", c_int),
    ]


class HubertCfg(ctypes.Structure):
    _fields_ = [("n_layers", c_int), ("embed_dim", c_int), ("ffn_dim", c_int), ("n_heads", c_int), ("conv_dim", c_int),
                ("pos_kernel", c_int), ("pos_groups", c_int), ("n_clusters", c_int)]


class Tensor(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("data", POINTER(c_float)), ("numel", c_int64)]


_lib = None

# every symbol include/dissc_b200.h declares (tests/test_abi.py checks header == this list == the .so)
EXPORTS = [
    "dissc_gen_create", "dissc_gen_destroy", "dissc_gen_hop", "dissc_gen_workspace_bytes", "dissc_gen_forward",
    "dissc_gen_forward_i16", "dissc_gen_forward_host", "dissc_gen_launches_per_forward", "dissc_gen_cost",
    "dissc_gen_profile", "dissc_conv1d_fused", "dissc_conv_transpose1d", "dissc_last_error", "dissc_version",
    "dissc_gen_set_tensor_cores", "dissc_gen_tensor_core_stages", "dissc_conv1d_tc", "dissc_conv_transpose1d_tc", "dissc_tc_set_single_accumulator", "dissc_tc_set_tuning", "dissc_resblock_pair_tc",
    "dissc_pred_create", "dissc_pred_destroy", "dissc_pred_workspace_bytes", "dissc_len_forward", "dissc_pitch_forward",
    "dissc_pitch_calc_freq", "dissc_len_carryover", "dissc_dedup_units", "dissc_repeat_interleave",
    "dissc_hubert_create", "dissc_hubert_destroy", "dissc_hubert_num_frames", "dissc_hubert_workspace_bytes",
    "dissc_hubert_forward", "dissc_kmeans_assign", "dissc_gen_status", "dissc_pred_status",
    "dissc_gen_forward_host_submit", "dissc_gen_forward_host_wait", "dissc_gen_host_reserve",
    "dissc_gen_forward_ex", "dissc_gen_n_extra",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DisscError(
            f"{LIB_PATH} is missing: build it with `python -m dissc_b200.build` (nvcc, sm_100a). "
            "dissc_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.dissc_last_error.restype = c_char_p
    L.dissc_version.restype = c_char_p
    L.dissc_gen_create.argtypes = [POINTER(c_void_p), POINTER(GenCfg), POINTER(Tensor), c_int, c_int]
    L.dissc_gen_destroy.argtypes = [c_void_p]
    L.dissc_gen_destroy.restype = None
    L.dissc_gen_hop.argtypes = [c_void_p]
    L.dissc_gen_launches_per_forward.argtypes = [c_void_p]
    L.dissc_gen_workspace_bytes.argtypes = [c_void_p, c_int, c_int, POINTER(c_size_t)]
    fwd = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.dissc_gen_forward.argtypes = fwd
    L.dissc_gen_forward_i16.argtypes = fwd
    L.dissc_gen_forward_host.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                         c_void_p]
    L.dissc_gen_cost.argtypes = [c_void_p, c_int, c_int, POINTER(c_double), POINTER(c_double)]
    L.dissc_gen_profile.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_void_p, c_size_t, c_void_p, POINTER(c_float), POINTER(c_double),
                                    POINTER(c_double), c_int, POINTER(c_int)]
    L.dissc_conv1d_fused.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_float,
                                     c_float, c_void_p]
    L.dissc_conv_transpose1d.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_void_p]
    L.dissc_gen_set_tensor_cores.argtypes = [c_void_p, c_int]
    L.dissc_gen_tensor_core_stages.argtypes = [c_void_p]
    L.dissc_tc_set_single_accumulator.argtypes = [c_int]
    L.dissc_tc_set_tuning.argtypes = [c_int, c_int]
    L.dissc_conv1d_tc.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                  c_float, c_float, c_void_p]
    L.dissc_conv_transpose1d_tc.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]
    L.dissc_resblock_pair_tc.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]
    L.dissc_pred_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, POINTER(Tensor), c_int, c_int]
    L.dissc_pred_destroy.argtypes = [c_void_p]
    L.dissc_pred_destroy.restype = None
    L.dissc_pred_workspace_bytes.argtypes = [c_void_p, c_int, c_int, POINTER(c_size_t)]
    L.dissc_len_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                    c_void_p, c_size_t, c_void_p]
    L.dissc_pitch_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]
    L.dissc_pitch_calc_freq.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p]
    L.dissc_gen_forward_host_submit.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                c_void_p, c_void_p]
    L.dissc_gen_forward_host_wait.argtypes = [c_void_p, c_int]
    L.dissc_gen_host_reserve.argtypes = [c_void_p, c_int, c_int]
    L.dissc_gen_forward_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]
    L.dissc_gen_n_extra.argtypes = [c_void_p]
    L.dissc_gen_status.argtypes = [c_void_p]
    L.dissc_pred_status.argtypes = [c_void_p]
    L.dissc_len_carryover.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.dissc_dedup_units.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.dissc_repeat_interleave.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p]
    L.dissc_hubert_create.argtypes = [POINTER(c_void_p), POINTER(HubertCfg), POINTER(Tensor), c_int, c_int]
    L.dissc_hubert_destroy.argtypes = [c_void_p]
    L.dissc_hubert_destroy.restype = None
    L.dissc_hubert_num_frames.argtypes = [c_int]
    L.dissc_hubert_workspace_bytes.argtypes = [c_void_p, c_int, c_int, POINTER(c_size_t)]
    L.dissc_hubert_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]
    L.dissc_kmeans_assign.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    _lib = L
    return L


EINDEX = -6


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dissc_last_error().decode("utf-8", "replace")
        if rc == EINDEX:   # what nn.Embedding raises for an id outside its table
            raise IndexError(f"{what or 'dissc_b200 call'}: {msg}")
        raise DisscError(f"{what or 'dissc_b200 call'} failed ({rc}): {msg}")
