"""dissc_b200 -- B200-native (sm_100a) implementation of the DISSC inference hot path.

Python here is host glue mirroring the reference's interfaces; all model
arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
``include/dissc_b200.h`` (``libdissc_b200.so``, built by ``dissc_b200.build``).
"""
from .models import AttrDict, CodeGenerator, get_padding  # noqa: F401

__all__ = ["AttrDict", "CodeGenerator", "get_padding"]
