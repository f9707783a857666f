"""One vocoder forward at BASELINE config 2 (B=64, T=300) -- the command profiled under ncu.

    python scripts/one_forward.py [B] [T] [n_forwards]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    dev = torch.device("cuda", 0)
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
    gen.load_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    gen.eval()
    gen.remove_weight_norm()
    code, f0, spkr = (t.to(dev) for t in syn.synthetic_inputs(B, T))
    for _ in range(n):
        y = gen(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize()
    print("ok", tuple(y.shape), float(y.std()))


if __name__ == "__main__":
    main()
