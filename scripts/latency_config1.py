"""BASELINE configs[0]: one utterance of 50 units (16 000 samples), Generator forward only -- latency on the GPU
(eager launches and one CUDA-graph replay) next to the oracle port on the host.   python scripts/latency_config1.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200 import AttrDict, CodeGenerator  # noqa: E402
from dissc_b200 import synthetic as syn  # noqa: E402
from oracle import generator_oracle as go  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = syn.VCTK_CONFIG
    sd = syn.synthetic_generator_state_dict(cfg, seed=0)
    gen = CodeGenerator(AttrDict(cfg)).to(dev)
    gen.load_state_dict(sd)
    gen.eval()
    gen.remove_weight_norm()
    for B, T in ((1, 50), (1, 300), (8, 300)):
        code, f0, spkr = syn.synthetic_inputs(B, T, seed=1234)
        c, f, s = code.to(dev), f0.to(dev), spkr.to(dev)
        for _ in range(5):
            y = gen(code=c, f0=f, spkr=s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            y = gen(code=c, f0=f, spkr=s)
        torch.cuda.synchronize()
        gpu_ms = (time.perf_counter() - t0) / n * 1e3
        gf = gen.capture_graph(B, T, dev)
        yg = gf(code=c, f0=f, spkr=s)
        torch.cuda.synchronize()
        same = bool(torch.equal(yg.view(-1), y.view(-1)))
        t0 = time.perf_counter()
        for _ in range(n):
            gf()
        torch.cuda.synchronize()
        graph_ms = (time.perf_counter() - t0) / n * 1e3
        fsd = go.folded_state_dict(sd)
        go.code_generator_forward(fsd, cfg, code, f0, spkr)
        t0 = time.perf_counter()
        for _ in range(3):
            ref = go.code_generator_forward(fsd, cfg, code, f0, spkr)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        err = (y.cpu() - ref).abs().max().item()
        print(f"B={B} T={T}: CUDA graph {graph_ms:.3f} ms (bit-identical to eager: {same}) | eager gpu {gpu_ms:.3f} ms ({B * T * 320 / gpu_ms / 1e3:.2f} M samples/s, {gen.launches_per_forward()} launches)"
              f" | cpu oracle {cpu_ms:.1f} ms ({torch.get_num_threads()} threads) | speed-up {cpu_ms / gpu_ms:.0f}x | max-abs err {err:.1e}")


if __name__ == "__main__":
    main()
