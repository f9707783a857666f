"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports
exactly the symbols include/dissc_b200.h declares; host-side mirrors keep the reference's
state-dict key set.  No compute calls (no GPU here)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dissc_b200 import _lib, build
    build.build()
    return _lib.lib()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dissc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dissc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from dissc_b200 import _lib
    syms = _header_symbols()
    assert syms == sorted(_lib.EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (dissc_\w+)", out))
    assert set(syms) <= exported, set(syms) - exported
    for s in syms:
        assert hasattr(lib, s)


def test_library_is_sm100a_with_tma(lib):
    from dissc_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    assert "sm_100a" in out
    sass = subprocess.check_output(["cuobjdump", "-sass", "-fun",
                                    "_ZN5dissc19conv1d_fused_kernelILi64ELi11ELi1ELi8ELb0EEEvNS_10ConvParamsE",
                                    _lib.LIB_PATH]).decode()
    assert "UBLKCP" in sass, "weight tiles must arrive by bulk TMA"
    assert "FFMA" in sass


def test_errors_without_gpu_are_loud(lib):
    from dissc_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes
    h = ctypes.c_void_p()
    cfg = _lib.GenCfg()
    arr = (_lib.Tensor * 1)()
    rc = lib.dissc_gen_create(ctypes.byref(h), ctypes.byref(cfg), arr, 0, 0)
    assert rc != 0 and lib.dissc_last_error()


def test_state_dict_keys_match_reference_checkpoint_format():
    from dissc_b200 import AttrDict, CodeGenerator
    from dissc_b200 import synthetic as syn
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG))
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    assert len(sd) == 293
    assert set(gen.state_dict().keys()) == set(sd.keys())
    res = gen.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(RuntimeError):
        gen.load_state_dict({k: v for k, v in sd.items() if k != "conv_post.bias"})
    # remove_weight_norm leaves plain `.weight`, like torch's remove_weight_norm; folding matches the oracle
    from oracle import generator_oracle as go
    folded = go.folded_state_dict(sd)
    gen.remove_weight_norm()
    mine = gen.folded_state_dict()
    assert set(mine) == set(folded)
    for k in folded:
        assert torch.allclose(mine[k], folded[k], atol=1e-7), k
    assert "conv_pre.weight" in gen.state_dict() and "conv_pre.weight_g" not in gen.state_dict()
    with pytest.raises(ValueError):
        gen.remove_weight_norm()


def test_cpu_forward_refuses():
    from dissc_b200 import AttrDict, CodeGenerator, _lib
    from dissc_b200 import synthetic as syn
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG))
    code, f0, spkr = syn.synthetic_inputs(1, 4)
    with pytest.raises(_lib.DisscError):
        gen(code=code, f0=f0, spkr=spkr)


def test_shard_by_length_balances():
    from dissc_b200.dist import shard_by_length
    lens = [300, 10, 250, 40, 200, 80, 150, 120, 7]
    shards = shard_by_length(lens, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(lens)))
    sums = [sum(lens[i] for i in s) for s in shards]
    assert max(sums) - min(sums) <= max(lens)
