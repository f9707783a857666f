"""Per-launch device times of one vocoder forward (cudaEvent around every kernel).

    python scripts/profile_layers.py [B] [T]      # default 64 300 (BASELINE config 2)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    dev = torch.device("cuda", 0)
    gen = CodeGenerator(AttrDict(syn.VCTK_CONFIG)).to(dev)
    gen.load_state_dict(syn.synthetic_generator_state_dict(syn.VCTK_CONFIG, seed=0))
    gen.eval()
    gen.remove_weight_norm()
    code, f0, spkr = (t.to(dev) for t in syn.synthetic_inputs(B, T))
    for _ in range(2):
        gen(code=code, f0=f0, spkr=spkr)
    torch.cuda.synchronize()
    rows = gen.profile(code, f0, spkr)
    tot = sum(r[1] for r in rows)
    agg = {}
    for name, ms, fl, by in rows:
        print(f"{name:16s} {ms:8.3f} ms  {fl / ms / 1e9 if ms > 0 else 0:8.2f} TFLOP/s  {by / ms / 1e6 if ms > 0 else 0:8.1f} GB/s (algorithmic)")
        key = name.split(".")[0] if name.startswith("s") else name.split(".")[0]
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += ms
        a[1] += fl
    print("---- per stage")
    for k, (ms, fl) in agg.items():
        print(f"{k:10s} {ms:8.3f} ms ({100 * ms / tot:5.1f}%)  {fl / ms / 1e9:8.2f} TFLOP/s")
    flops = sum(r[2] for r in rows)
    print(f"TOTAL {tot:.3f} ms  {flops / tot / 1e9:.2f} TFLOP/s  {B * T * 320 / tot / 1e3:.2f} Msamples/s")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/layers_B{B}_T{T}.json", "w"))


if __name__ == "__main__":
    main()
