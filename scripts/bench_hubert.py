"""Throughput of the HuBERT-base unit encoder (BASELINE configs[3] shape: 96 000-sample clips), one GPU.

    python scripts/bench_hubert.py [B] [N] [iters]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissc_b200.hubert import SpeechEncoder  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 96000
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    import torchaudio
    from oracle import hubert_oracle as ho   # weight-name mapping only (synthetic weights come from torchaudio's init)
    torch.manual_seed(0)
    sd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
    cent = torch.randn(100, 768)
    dev = torch.device("cuda", 0)
    enc = SpeechEncoder.from_state_dict(sd, cent).to(dev)
    wave = 0.1 * torch.randn(B, N, device=dev)
    for _ in range(2):
        enc.encode_batch(wave, return_dense=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        units, nf, _ = enc.encode_batch(wave, return_dense=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 59.6e9 * B * (N / 96000)
    print(f"hubert encode: B={B} N={N}: {ms:.2f} ms/batch  {B / ms * 1e3:.1f} clips/s  {B * N / ms / 1e3:.1f} M samples/s  "
          f"{flops / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)  frames {int(nf[0])}")
    # CPU oracle timing on one clip for scale
    t0 = time.perf_counter()
    ho.extract_features(sd, wave[:1].cpu(), 6)
    print(f"cpu oracle (torch, {torch.get_num_threads()} threads): {1e3 * (time.perf_counter() - t0):.0f} ms / clip")


if __name__ == "__main__":
    main()
