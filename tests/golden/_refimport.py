"""Import the real reference (/root/reference) in the build container.

Only used by make_golden.py / the optional `refcheck` tests; never on the GPU
box (the reference does not travel).  Shims: matplotlib (sr/utils.py:13-17)
and tensorflow (utils.py:2) are absent here and only used for plotting /
summaries.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("DISSC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "sr", "models.py"))


def _shim():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        pylab = types.ModuleType("matplotlib.pylab")
        mpl.pylab = pylab
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pylab"] = pylab
    if "tensorflow" not in sys.modules:
        tf = types.ModuleType("tensorflow")
        tf.summary = types.ModuleType("tensorflow.summary")
        sys.modules["tensorflow"] = tf
        sys.modules["tensorflow.summary"] = tf.summary


def _import_from(path_first, name):
    """Import module `name` with `path_first` at the head of sys.path, isolated
    from same-named modules (the reference has both utils.py and sr/utils.py)."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("utils", "models", "modules", "dataset", "model", "infer")
             or k.startswith(("modules.", "dataset.", "model."))}
    sys.path.insert(0, path_first)
    try:
        mod = importlib.import_module(name)
        loaded = {k: sys.modules[k] for k in list(sys.modules)
                  if k in ("utils", "models", "modules", "dataset", "model", "infer")
                  or k.startswith(("modules.", "dataset.", "model."))}
    finally:
        sys.path.remove(path_first)
        for k in list(sys.modules):
            if k in ("utils", "models", "modules", "dataset", "model", "infer") \
                    or k.startswith(("modules.", "dataset.", "model.")):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return mod, loaded


def sr_models():
    """-> the reference's sr/models.py module and its AttrDict."""
    _shim()
    mod, loaded = _import_from(os.path.join(REF, "sr"), "models")
    return mod, loaded["utils"].AttrDict


def predictors():
    """-> (model.len_predictor, model.pitch_predictor) modules."""
    _shim()
    lp, _ = _import_from(REF, "model.len_predictor")
    pp, _ = _import_from(REF, "model.pitch_predictor")
    return lp, pp


def infer_module():
    """-> the reference's infer.py (len_carryover_correction, _infer_sample)."""
    _shim()
    mod, _ = _import_from(REF, "infer")
    return mod
