"""Max-abs / RMS error of the CUDA vocoder against the oracle in fp64 and fp32 (test infrastructure: uses oracle/).

    python scripts/parity_report.py [B] [T]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402
from oracle import generator_oracle as go  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    dev = torch.device("cuda", 0)
    cfg = syn.VCTK_CONFIG
    sd = syn.synthetic_generator_state_dict(cfg, seed=0)
    gen = CodeGenerator(AttrDict(cfg)).to(dev)
    gen.load_state_dict(sd)
    gen.eval()
    gen.remove_weight_norm()
    code, f0, spkr = syn.synthetic_inputs(B, T, seed=1234)
    y = gen(code=code.to(dev), f0=f0.to(dev), spkr=spkr.to(dev)).cpu()
    ref64 = go.code_generator_forward(sd, cfg, code, f0, spkr, dtype=torch.float64)
    ref32 = go.code_generator_forward(sd, cfg, code, f0, spkr, dtype=torch.float32)
    e64 = (y.double() - ref64)
    e32 = (y - ref32)
    o64 = (ref32.double() - ref64)
    print(f"B={B} T={T} out std {y.std():.3f}  cuda-vs-fp64 max {e64.abs().max():.3e} rms {e64.pow(2).mean().sqrt():.3e} | "
          f"cuda-vs-fp32oracle max {e32.abs().max():.3e} | fp32oracle-vs-fp64 max {o64.abs().max():.3e} "
          f"rms {o64.pow(2).mean().sqrt():.3e}  single_acc={os.environ.get('DISSC_TC_SINGLE_ACC', '0')} "
          f"tc={os.environ.get('DISSC_TC', '1')}")


if __name__ == "__main__":
    main()
