"""Host-side mirror of the reference prosody predictors (model/len_predictor.py,
model/pitch_predictor.py) -- same constructors, state-dict keys, attributes and call
signatures as ``infer.py:68-84,30,38`` uses -- with the arithmetic in
``libdissc_b200.so`` (csrc/predictors.cu).  The modules below only HOLD parameters
(so ``load_state_dict(torch.load(dir + 'best_model.pth'))`` works unchanged); none of
their ``forward`` methods is ever called.  No CPU fallback.

Extension over the reference (which always runs B=1): ``seq`` may be a padded batch
``(B, L)`` with ``lengths=`` giving the valid tokens per row; each row then equals the
reference's B=1 result on the unpadded sequence.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib

KIND_LEN, KIND_PITCH_NEW, KIND_PITCH_BASE = 0, 1, 2
_C = 128


class _Buf(nn.Module):
    """Holder for the ``pe.pe`` buffer (PositionalEncoding, model/pitch_predictor.py:6-17)."""

    def __init__(self, d_model, max_len=850):
        super().__init__()
        up = torch.linspace(0, 1, max_len).unsqueeze(-1).repeat_interleave(d_model // 2, dim=-1)
        down = torch.linspace(1, 0, max_len).unsqueeze(-1).repeat_interleave(d_model // 2, dim=-1)
        self.register_buffer("pe", torch.cat([up, down], dim=-1).unsqueeze(0))


class _Predictor(nn.Module):
    kind = -1

    def _build(self, n_tokens, n_speakers, emb_size, spk_rows, convs, bns):
        self.n_tokens, self.n_speakers, self.emb_size = n_tokens, n_speakers, emb_size
        self.token_emb = nn.Embedding(n_tokens + 1, emb_size, padding_idx=n_tokens)
        self.spk_emb = nn.Embedding(spk_rows, emb_size, padding_idx=(n_speakers if spk_rows > n_speakers else None))
        for name, cin, cout, k in convs:
            setattr(self, name, nn.Conv1d(cin, cout, kernel_size=(k,), padding=(k - 1) // 2))
        for name in bns:
            setattr(self, name, nn.BatchNorm1d(_C))
        self.__dict__["_handle"] = None
        self.__dict__["_handle_device"] = None
        self.__dict__["_ws"] = None
        self.eval()

    # ---- handle management -------------------------------------------------
    def _drop_handle(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            _lib.lib().dissc_pred_destroy(h)
        self.__dict__["_handle"] = None
        self.__dict__["_ws"] = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._drop_handle()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._drop_handle()
        return super()._apply(fn, *a, **kw)

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("dissc_b200 predictors are inference-only (eval mode)")
        return super().train(False)

    def _ensure_handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return self._handle
        self._drop_handle()
        sd = {k: v.detach().float().cpu().contiguous() for k, v in self.state_dict().items() if v.is_floating_point()}
        arr = (_lib.Tensor * len(sd))()
        keep = []
        for i, (k, v) in enumerate(sd.items()):
            keep.append((k.encode(), v))
            arr[i].name = keep[-1][0]
            arr[i].data = ctypes.cast(v.data_ptr(), ctypes.POINTER(ctypes.c_float))
            arr[i].numel = v.numel()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().dissc_pred_create(ctypes.byref(h), self.kind, self.n_tokens, self.n_speakers, arr,
                                                    len(sd), device.index), "dissc_pred_create")
        self.__dict__["_handle"], self.__dict__["_handle_device"] = h, device
        return h

    def _prep(self, seq, spk_id, lengths):
        if not seq.is_cuda:
            raise _lib.DisscError("dissc_b200 predictors run only on CUDA (sm_100a); there is no CPU path")
        dev = seq.device
        seq = seq.to(torch.int64).contiguous()
        B, L = seq.shape
        spk = spk_id.to(device=dev, dtype=torch.int64).reshape(B).contiguous()
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        h = self._ensure_handle(dev)
        need = ctypes.c_size_t()
        _lib.check(_lib.lib().dissc_pred_workspace_bytes(h, B, L, ctypes.byref(need)))
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != dev:
            self.__dict__["_ws"] = torch.empty(need.value, dtype=torch.uint8, device=dev)
        return h, seq, spk, lengths, B, L, dev


    def check_indices(self, synchronize=True):
        """Raises ``IndexError`` if a token / speaker id outside its embedding table reached the kernels since the last
        check (``nn.Embedding`` raises at the call; here the gather stays in bounds and a device-visible flag is set)."""
        if self._handle is None:
            return
        if synchronize:
            torch.cuda.synchronize(self._handle_device)
        _lib.check(_lib.lib().dissc_pred_status(self._handle), "predictor")


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class LenPredictor(_Predictor):
    """Drop-in for ``model/len_predictor.py::LenPredictor`` (inference)."""
    kind = KIND_LEN

    def __init__(self, n_tokens=100, n_speakers=99, emb_size=32, masking_rate=0.2, norm_mean=torch.tensor(0),
                 norm_std=torch.tensor(1)):
        super().__init__()
        self.keep_rate = 1 - masking_rate
        self.norm_mean, self.norm_std = norm_mean, norm_std   # overwritten from len_norm_stats.pth (infer.py:72)
        names = ["cnn1"] + [f"cnn1{i}" for i in range(1, 7)]
        convs = [(n, 2 * emb_size if n == "cnn1" else _C, _C, 3) for n in names] + [("cnn2", _C, 1, 3)]
        self._build(n_tokens, n_speakers, emb_size, n_speakers, convs, ["bn" + n[3:] for n in names])

    def forward(self, seq, spk_id, lengths=None):
        h, seq, spk, lengths, B, L, dev = self._prep(seq, spk_id, lengths)
        out = torch.empty((B, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dissc_len_forward(h, _p(seq), _p(spk), _p(lengths), B, L, float(self.norm_mean),
                                                    float(self.norm_std), _p(out), _p(self._ws), self._ws.numel(),
                                                    _stream(dev)), "dissc_len_forward")
        return out


class _PitchCommon(_Predictor):
    def forward(self, seq, spk_id, lengths=None):
        h, seq, spk, lengths, B, L, dev = self._prep(seq, spk_id, lengths)
        cls = torch.empty((B, L), dtype=torch.float32, device=dev)
        reg = torch.empty((B, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dissc_pitch_forward(h, _p(seq), _p(spk), _p(lengths), B, L, _p(cls), _p(reg),
                                                      _p(self._ws), self._ws.numel(), _stream(dev)),
                       "dissc_pitch_forward")
        return cls, reg

    def infer_freq(self, seq, spk_id, norm=False, lengths=None):
        cls, reg = self(seq, spk_id, lengths=lengths)
        return self.calc_freq(cls, reg, spk_id, norm, lengths=lengths)

    def calc_freq(self, class_preds, reg_preds, spk_id, norm=False, lengths=None):
        dev = class_preds.device
        B, L = class_preds.shape
        out = torch.empty_like(reg_preds)
        spk = spk_id.to(device=dev, dtype=torch.int64).reshape(B).contiguous()
        mean = std = None
        if not norm:
            mean = self.id2pitch_mean.to(device=dev, dtype=torch.float32).contiguous()
            std = self.id2pitch_std.to(device=dev, dtype=torch.float32).contiguous()
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dissc_pitch_calc_freq(_p(class_preds.contiguous()), _p(reg_preds.contiguous()), _p(spk),
                                                        _p(mean), _p(std), 0 if mean is None else mean.numel(),
                                                        _p(lengths), B, L, _p(out), _stream(dev)),
                       "dissc_pitch_calc_freq")
        return out


class PitchPredictor(_PitchCommon):
    """Drop-in for ``model/pitch_predictor.py::PitchPredictor`` ("new": positional table on the speaker embedding,
    BatchNorm only after cnn2)."""
    kind = KIND_PITCH_NEW

    def __init__(self, n_tokens=100, n_speakers=199, emb_size=32, masking_rate=0.4, id2pitch_mean=None,
                 id2pitch_std=None):
        super().__init__()
        self.keep_rate = 1 - masking_rate
        self.id2pitch_mean, self.id2pitch_std = id2pitch_mean, id2pitch_std
        names = ["cnn1"] + [f"cnn1{i}" for i in range(1, 8)]
        convs = [(n, 2 * emb_size if n == "cnn1" else _C, _C, 3) for n in names]
        convs += [("cnn2", _C, _C, 3), ("cnn_class1", _C, _C, 3), ("cnn_class2", _C, 1, 1), ("cnn_reg1", _C, _C, 3),
                  ("cnn_reg2", _C, 1, 1)]
        self._build(n_tokens, n_speakers, emb_size, n_speakers + 1, convs, ["bn2"])
        self.pe = _Buf(emb_size)


class PitchPredictorBase(_PitchCommon):
    """Drop-in for ``model/pitch_predictor.py::PitchPredictorBase`` (BatchNorm after every conv but cnn2)."""
    kind = KIND_PITCH_BASE

    def __init__(self, n_tokens=100, n_speakers=199, emb_size=32, masking_rate=0.4, id2pitch_mean=None,
                 id2pitch_std=None):
        super().__init__()
        self.keep_rate = 1 - masking_rate
        self.id2pitch_mean, self.id2pitch_std = id2pitch_mean, id2pitch_std
        names = ["cnn1"] + [f"cnn1{i}" for i in range(1, 8)]
        convs = [(n, 2 * emb_size if n == "cnn1" else _C, _C, 3) for n in names]
        convs += [("cnn2", _C, _C, 3), ("cnn_class1", _C, _C, 3), ("cnn_class2", _C, 1, 1), ("cnn_reg1", _C, _C, 3),
                  ("cnn_reg2", _C, 1, 1)]
        self._build(n_tokens, n_speakers, emb_size, n_speakers + 1, convs,
                    ["bn" + n[3:] for n in names] + ["bn_c1", "bn_r1"])
