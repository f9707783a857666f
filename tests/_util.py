import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def tiny_config():
    from dissc_b200 import synthetic as syn
    return dict(syn.VCTK_CONFIG, upsample_initial_channel=32, embedding_dim=8, model_in_dim=17)


def tiny_state_dict(g):
    return {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}


def make_generator(cfg, sd, device):
    """The sequence sr/inference.py:114-120,162-163 performs."""
    from dissc_b200 import AttrDict, CodeGenerator
    gen = CodeGenerator(AttrDict(cfg)).to(device)
    gen.load_state_dict(sd)
    gen.eval()
    gen.remove_weight_norm()
    return gen
