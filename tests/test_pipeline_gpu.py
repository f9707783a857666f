"""End to end on the GPU: waveform -> HuBERT units (encode) -> rhythm + pitch conversion (infer) -> vocoder (inference),
each stage through its C-ABI entry point, checked stage by stage against the oracles (BASELINE configs[4] in miniature)."""
import numpy as np
import pytest
import torch

from dissc_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_encode_predict_vocode_chain(cuda_device):
    torchaudio = pytest.importorskip("torchaudio")
    from dissc_b200 import AttrDict, CodeGenerator
    from dissc_b200.hubert import SpeechEncoder
    from dissc_b200.infer import convert_batch
    from dissc_b200.predictors import LenPredictor, PitchPredictor
    from oracle import generator_oracle as go
    from oracle import hubert_oracle as ho
    from oracle import predictors_oracle as po
    dev = cuda_device
    torch.manual_seed(0)
    hsd = ho.from_torchaudio(torchaudio.models.hubert_base().eval(), 6)
    g = torch.Generator().manual_seed(11)
    lens = [20000, 14000, 9000]
    waves = [0.1 * torch.randn(n, generator=g) for n in lens]
    feats = [ho.extract_features(hsd, w.view(1, -1), 6)[0] for w in waves]
    allf = torch.cat(feats)
    cent = allf[torch.randperm(allf.shape[0], generator=g)[:100]] + 0.02 * torch.randn(100, 768, generator=g)
    # 1. encode
    enc = SpeechEncoder.from_state_dict(hsd, cent).to(dev)
    wave = torch.zeros(len(lens), max(lens))
    for b, w in enumerate(waves):
        wave[b, :len(w)] = w
    units, n_frames, _ = enc.encode_batch(wave.to(dev), torch.tensor(lens, dtype=torch.int32), return_dense=False)
    for b, f in enumerate(feats):
        d = ho.kmeans_distances(f.double(), cent.double())
        top2 = d.topk(2, dim=-1, largest=False).values
        sure = (top2[:, 1] - top2[:, 0]) / top2[:, 1].clamp(min=1e-12) > 1e-3
        assert torch.equal(units[b, :f.shape[0]].cpu()[sure], d.argmin(-1)[sure])
    # 2. prosody conversion to a target speaker (units padded with n_tokens = 100, as infer.py does)
    lm = LenPredictor(100, 108).to(dev)
    len_sd = syn.synthetic_len_predictor_state_dict(100, 108, seed=21)
    lm.load_state_dict(len_sd)
    lm.norm_mean, lm.norm_std = torch.tensor(2.5), torch.tensor(1.5)
    mean, std = syn.synthetic_pitch_stats(108, seed=22)
    pm = PitchPredictor(100, 108, id2pitch_mean=mean.to(dev), id2pitch_std=std.to(dev)).to(dev)
    psd = syn.synthetic_pitch_predictor_state_dict("new", 100, 108, seed=23)
    pm.load_state_dict(psd)
    seqs = units.clone()
    seqs[seqs < 0] = 100
    spk = torch.tensor([[5], [17], [99]], device=dev)
    out_seq, f0, out_len = convert_batch(seqs, spk, 100, lm, pm, norm_pitch=True)
    for b in range(len(lens)):
        src = units[b, :int(n_frames[b])].cpu().tolist()
        wu, wf = po.infer_sample(src, int(spk[b]), 100, len_sd, (torch.tensor(2.5), torch.tensor(1.5)), psd, "new", mean, std, True)
        n = int(out_len[b])
        assert out_seq[b, :n].cpu().tolist() == wu.tolist()
        assert (f0[b, :n].cpu() - wf).abs().max().item() < 1e-3
    # 3. vocode the converted (units, f0, speaker) with per-utterance lengths
    cfg = syn.VCTK_CONFIG
    gsd = syn.synthetic_generator_state_dict(cfg, seed=0)
    gen = CodeGenerator(AttrDict(cfg)).to(dev)
    gen.load_state_dict(gsd)
    gen.eval()
    gen.remove_weight_norm()
    code = out_seq.clone()
    code[code >= 100] = 0                                # padding rows are masked by `lengths`
    y = gen(code=code, f0=f0.unsqueeze(1), spkr=spk, lengths=out_len).cpu()
    for b in range(len(lens)):
        n = int(out_len[b])
        ref = go.code_generator_forward(gsd, cfg, out_seq[b:b + 1, :n].cpu(), f0[b:b + 1, :n].cpu().unsqueeze(1), spk[b:b + 1].cpu())
        assert (y[b, 0, :320 * n] - ref[0, 0]).abs().max().item() < 1e-4
        assert torch.all(y[b, 0, 320 * n:] == 0)
