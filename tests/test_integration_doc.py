"""INTEGRATION.md, Option B: the ctypes stub a reference maintainer pastes must actually bind the library.

* CPU, always: the snippet is extracted from INTEGRATION.md and executed against the built .so: its structs must have the
  layout of include/dissc_b200.h (== dissc_b200/_lib.py), every function it calls must be exported.
* CPU, `refcheck` (build container only, needs /root/reference): the snippet's ``make_handle`` is driven from the
  REFERENCE's own ``sr/models.py::CodeGenerator`` after ``load_state_dict`` + ``remove_weight_norm`` -- its
  ``state_dict()`` must carry exactly the tensors ``dissc_gen_create`` looks up, with the values dissc_b200's own fold
  produces; without a GPU the call must fail cleanly with the CUDA error in ``dissc_last_error()`` (no crash, no fallback).
* GPU: the same snippet, fed with dissc_b200's folded state dict (the reference does not travel to the GPU box), must
  reproduce ``CodeGenerator.generate_int16`` bit for bit.
"""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _option_b_namespace():
    from dissc_b200 import _lib
    _lib.lib()   # raises if the library is not built
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("### Option B"):]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    code = code.replace("/path/to/dissc-b200/dissc_b200/libdissc_b200.so", _lib.LIB_PATH)
    ns = {}
    exec(compile(code, "INTEGRATION.md::OptionB", "exec"), ns)
    return ns


def test_option_b_snippet_binds_the_library():
    from dissc_b200 import _lib
    ns = _option_b_namespace()
    assert ctypes.sizeof(ns["Cfg"]) == ctypes.sizeof(_lib.GenCfg)
    assert [f[0] for f in ns["Cfg"]._fields_] == [f[0] for f in _lib.GenCfg._fields_]
    assert ctypes.sizeof(ns["Tensor"]) == ctypes.sizeof(_lib.Tensor)
    for fn in ("dissc_gen_create", "dissc_gen_workspace_bytes", "dissc_gen_hop", "dissc_gen_forward_i16",
               "dissc_gen_status", "dissc_last_error"):
        assert hasattr(ns["L"], fn), fn


@pytest.mark.refcheck
def test_option_b_from_the_reference_state_dict():
    import _refimport
    if not _refimport.available():
        pytest.skip("needs /root/reference (build container only)")
    from _util import make_generator
    from dissc_b200 import synthetic as syn
    models, AttrDict = _refimport.sr_models()
    h = AttrDict(syn.VCTK_CONFIG)
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    ref = models.CodeGenerator(h)
    ref.load_state_dict(sd)            # sr/inference.py:120
    ref.eval()
    ref.remove_weight_norm()           # sr/inference.py:162-163
    ref_sd = ref.state_dict()
    # what dissc_gen_create looks up == what the reference's folded state dict holds, value for value
    ours = make_generator(syn.VCTK_CONFIG, sd, torch.device("cpu")).folded_state_dict()
    assert set(ours) == set(ref_sd)
    for k, v in ours.items():
        assert v.shape == ref_sd[k].shape, k
        assert torch.allclose(v, ref_sd[k].float(), rtol=0, atol=1e-6), k
    ns = _option_b_namespace()
    if torch.cuda.is_available():
        pytest.skip("the reference-side half of this test is for the GPU-less build container")
    with pytest.raises(RuntimeError) as e:          # no GPU here: a clean error, not a crash or a CPU fallback
        ns["make_handle"](ref, h, 0)
    assert "cuda" in str(e.value).lower()


@pytest.mark.gpu
def test_option_b_snippet_runs_on_the_gpu(cuda_device):
    from _util import make_generator
    from dissc_b200 import AttrDict
    from dissc_b200 import synthetic as syn
    ns = _option_b_namespace()
    sd = syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0)
    gen = make_generator(syn.VCTK_CONFIG, sd, cuda_device)

    class Folded:                      # stands in for the reference module after remove_weight_norm(): same state_dict()
        def state_dict(self):
            return gen.folded_state_dict()

    handle = ns["make_handle"](Folded(), AttrDict(syn.VCTK_CONFIG), cuda_device.index or 0)
    code, f0, spkr = (t.to(cuda_device) for t in syn.synthetic_inputs(2, 33, seed=6))
    with torch.cuda.device(cuda_device):
        got = ns["generate"](handle, code, f0.reshape(2, 33), spkr.reshape(2))
    torch.cuda.synchronize()
    want = gen.generate_int16(code, f0, spkr)
    assert torch.equal(got, want)
    ns["L"].dissc_gen_destroy.argtypes = [ctypes.c_void_p]
    ns["L"].dissc_gen_destroy(handle)
