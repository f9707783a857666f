"""Golden vectors for the HuBERT-base unit encoder from a SECOND independent implementation of the published graph:
Hugging Face ``transformers.HubertModel`` (config = hubert-base), randomly initialised with a fixed seed.

    python tests/golden/make_golden_hubert.py

The reference's own encoder (textlesslib HEAD + fairseq@dd106d95, data/encode.py:7,21-22,32) is third-party code that is
absent offline, so it cannot generate vectors; transformers' HubertModel is the conversion target of fairseq's
hubert_base_ls960 checkpoint (convert_hubert_original_pytorch_checkpoint_to_pytorch.py) and computes the same graph.
Stored: the waveform seed, a checksum of the model's weights (detects RNG drift on the box that re-creates the model),
and rows of ``hidden_states[6]`` at 96 000 and 160 000 samples.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SEED_MODEL, SEED_WAVE = 5, 21
ROWS = {96000: [0, 1, 57, 150, 297, 298], 160000: [0, 250, 498]}


def build_model():
    from transformers import HubertConfig, HubertModel
    torch.manual_seed(SEED_MODEL)
    return HubertModel(HubertConfig()).eval()


def weights_checksum(model) -> float:
    tot = 0.0
    for k, v in sorted(model.state_dict().items()):
        tot += float(v.double().abs().sum())
    return tot


def waves():
    g = torch.Generator().manual_seed(SEED_WAVE)
    return {n: 0.1 * torch.randn(1, n, generator=g) for n in ROWS}


def main():
    torch.set_num_threads(1)
    m = build_model()
    out = {"checksum": np.float64(weights_checksum(m))}
    for n, w in waves().items():
        with torch.no_grad():
            h = m(w, output_hidden_states=True).hidden_states[6][0]
        out[f"rows_{n}"] = np.asarray(ROWS[n])
        out[f"feat_{n}"] = h[ROWS[n]].numpy()
        out[f"wave_head_{n}"] = w[0, :8].numpy()
        print(n, tuple(h.shape), float(h.std()))
    np.savez_compressed(os.path.join(HERE, "hubert_hf_small.npz"), **out)


if __name__ == "__main__":
    main()
