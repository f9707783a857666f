"""Host-side mirror of the reference vocoder interface (sr/models.py).

``CodeGenerator`` keeps the surface ``sr/inference.py`` uses --
``CodeGenerator(h).to(dev)``, ``load_state_dict(ckpt['generator'])`` with the
reference's exact key set (``*.weight_g`` / ``*.weight_v`` / ``*.bias``,
``dict.weight``, ``spkr.weight``), ``eval()``, ``remove_weight_norm()`` and
``generator(code=..., f0=..., spkr=...) -> (B,1,hop*T)`` (sr/inference.py:69,
:114-120,:162-163) -- but owns no arithmetic: the forward is one call into
``libdissc_b200.so`` (hand-written sm_100a kernels).  PyTorch is used for
parameter bookkeeping, device memory and the stream only.  There is no
fallback path: without the CUDA library or off-GPU, ``forward`` raises.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib

LRELU_SLOPE = 0.1  # sr/models.py:13


class AttrDict(dict):
    """sr/utils.py:77-80."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def get_padding(kernel_size, dilation=1):
    """sr/utils.py:44-45."""
    return int((kernel_size * dilation - dilation) / 2)


class _WNConv(nn.Module):
    """Parameter holder with the state-dict keys of a weight-normed conv
    (``weight_g``, ``weight_v``, ``bias``); after ``fold()`` it holds ``weight``
    instead, exactly like ``torch.nn.utils.remove_weight_norm`` leaves it."""

    def __init__(self, shape, n_bias, init_std=0.01):
        super().__init__()
        v = torch.empty(shape).normal_(0.0, init_std)  # init_weights, sr/utils.py:32-35
        self.weight_g = nn.Parameter(v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1), requires_grad=False)
        self.weight_v = nn.Parameter(v, requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(n_bias), requires_grad=False)

    def folded(self) -> torch.Tensor:
        if "weight" in self._parameters:
            return self.weight.detach()
        v, g = self.weight_v.detach(), self.weight_g.detach()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / norm)

    def fold(self):
        if "weight" in self._parameters:
            raise ValueError("weight_norm already removed")  # same failure mode as torch's remove_weight_norm
        w = self.folded()
        del self._parameters["weight_g"]
        del self._parameters["weight_v"]
        self.weight = nn.Parameter(w, requires_grad=False)


class _ResBlockParams(nn.Module):
    def __init__(self, channels, kernel_size, dilation, kind):
        super().__init__()
        mk = lambda: _WNConv((channels, channels, kernel_size), channels)
        if kind == "1":   # ResBlock1, sr/models.py:16-32
            self.convs1 = nn.ModuleList([mk() for _ in dilation])
            self.convs2 = nn.ModuleList([mk() for _ in dilation])
        else:             # ResBlock2, :50-60
            self.convs = nn.ModuleList([mk() for _ in dilation])

    def remove_weight_norm(self):
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.fold()


class _Table(nn.Module):
    def __init__(self, rows, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(rows, dim), requires_grad=False)


class CodeGenerator(nn.Module):
    """Drop-in for ``sr/models.py::CodeGenerator`` (inference only)."""

    def __init__(self, h):
        super().__init__()
        self.h = h
        get = h.get if hasattr(h, "get") else (lambda k, d=None: getattr(h, k, d))
        for key in ("lambda_commit", "lambda_commit_code", "f0_quantizer_path"):
            if get(key, None):
                raise NotImplementedError(
                    f"config key '{key}' selects a VQ branch of sr/models.py:137-156 that dissc_b200 does not implement")
        self.resblock = str(h["resblock"])
        self.rates = list(h["upsample_rates"])
        self.up_kernels = list(h["upsample_kernel_sizes"])
        self.rks = list(h["resblock_kernel_sizes"])
        self.rds = [list(d) for d in h["resblock_dilation_sizes"]]
        self.c0 = int(h["upsample_initial_channel"])
        self.num_kernels = len(self.rks)
        self.num_upsamples = len(self.rates)
        self.in_dim = int(get("model_in_dim", 128))
        self.emb_dim = int(h["embedding_dim"])
        self.f0 = get("f0", None)
        self.multispkr = get("multispkr", None)

        self.conv_pre = _WNConv((self.c0, self.in_dim, 7), self.c0)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = self.c0
        for i, (u, k) in enumerate(zip(self.rates, self.up_kernels)):
            ci, co = self.c0 // (2 ** i), self.c0 // (2 ** (i + 1))
            self.ups.append(_WNConv((ci, co, k), co))
            ch = co
            for k_r, d_r in zip(self.rks, self.rds):
                self.resblocks.append(_ResBlockParams(ch, k_r, d_r, self.resblock))
        self.conv_post = _WNConv((1, ch, 7), 1)
        self.dict = _Table(int(h["num_embeddings"]), self.emb_dim)
        if self.multispkr:
            self.spkr = _Table(200, self.emb_dim)  # sr/models.py:133
        self._handle = None
        self._handle_device = None
        self._ws = None

    # ---- reference surface -------------------------------------------------
    def remove_weight_norm(self):
        """sr/models.py:116-122."""
        for m in self.modules():
            if isinstance(m, _WNConv):
                m.fold()
        self._drop_handle()

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._drop_handle()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._drop_handle()
        return super()._apply(fn, *a, **kw)

    def _extra_channels(self, kwargs, B):
        """Every keyword argument other than code / f0 / spkr is a conditioning feature appended as channels, in
        keyword order, after nearest-repeat upsampling over time (sr/models.py:216-221; `f0_feats` configs pass
        `f0_stats` = [mean, std] per utterance, sr/inference.py:237-245).  Supported: features that are constant over
        time, i.e. (B, n) or (B, n, 1) -> (B, n_extra) fp32; model_in_dim must account for them."""
        feats = []
        for k, v in kwargs.items():
            if k in ("code", "f0", "spkr"):
                continue
            if v.dim() == 3 and v.shape[-1] == 1:
                v = v[..., 0]
            if v.dim() != 2 or v.shape[0] != B:
                raise NotImplementedError(f"conditioning feature '{k}' of shape {tuple(v.shape)}: only per-utterance "
                                          "features (B, n) / (B, n, 1) are supported (sr/models.py:216-221)")
            feats.append(v.to(torch.float32))
        n_extra = self.in_dim - self.emb_dim - (1 if self.f0 else 0) - (self.emb_dim if self.multispkr else 0)
        got = sum(f.shape[1] for f in feats)
        if got != n_extra:
            raise ValueError(f"model_in_dim={self.in_dim} leaves {n_extra} extra conditioning channel(s), got {got} "
                             f"({[k for k in kwargs if k not in ('code', 'f0', 'spkr')]})")
        return torch.cat(feats, dim=1).contiguous() if feats else None

    def forward(self, **kwargs):
        lengths = kwargs.pop("lengths", None)  # extension: per-utterance valid frames (varlen batches)
        code = kwargs["code"]
        if not code.is_cuda:
            raise _lib.DisscError("dissc_b200.CodeGenerator runs only on CUDA (sm_100a); there is no CPU path")
        if code.dtype != torch.int64:
            raise TypeError("code must be int64 unit ids (float 'code' selects the code_vq branch, unsupported)")
        f0 = kwargs.get("f0", None) if self.f0 else None
        spkr = kwargs.get("spkr", None) if self.multispkr else None
        if self.f0 and f0 is None:
            raise KeyError("f0")
        if self.multispkr and spkr is None:
            raise KeyError("spkr")
        B, T = code.shape
        if f0 is not None:
            f0 = f0.reshape(B, -1)
            Tf = f0.shape[-1]
            # _upsample semantics (sr/models.py:158-177,:207-210): nearest-repeat the shorter one
            if Tf != T:
                if T < Tf and Tf % T == 0:
                    code = code.repeat_interleave(Tf // T, dim=1)
                    T = Tf
                elif Tf < T and T % Tf == 0:
                    f0 = f0.repeat_interleave(T // Tf, dim=1)
                else:
                    raise NotImplementedError(
                        "Padding condition signal - misalignment between condition features.")
            f0 = f0.to(torch.float32).contiguous()
        code = code.contiguous()
        if spkr is not None:
            spkr = spkr.reshape(B).to(torch.int64).contiguous()
        if lengths is not None:
            lengths = lengths.to(device=code.device, dtype=torch.int32).contiguous()
        extra = self._extra_channels(kwargs, B)
        return self._run(code, f0, spkr, lengths, B, T, out_dtype=torch.float32, extra=extra).view(B, 1, -1)

    def generate_int16(self, code, f0=None, spkr=None, lengths=None, out=None, **features):
        """Fused ``generate()`` of sr/inference.py:67-76: returns int16 (B, hop*T) on the device.  ``out``: an existing
        contiguous int16 (B, hop*T) device tensor the last kernel writes into directly (e.g. the send buffer of a
        gather, ``dist.ScatterGatherPipeline``)."""
        B, T = code.shape
        f0 = None if f0 is None else f0.reshape(B, T).to(torch.float32).contiguous()
        spkr = None if spkr is None else spkr.reshape(B).to(torch.int64).contiguous()
        if lengths is not None:
            lengths = lengths.to(device=code.device, dtype=torch.int32).contiguous()
        extra = self._extra_channels(features, B)
        return self._run(code.contiguous(), f0, spkr, lengths, B, T, out_dtype=torch.int16, out=out, extra=extra)

    # ---- C-ABI plumbing ------------------------------------------------------
    def folded_state_dict(self):
        """{reference key with .weight/.bias: fp32 cpu tensor}, weight-norm folded."""
        out = {}
        for name, m in self.named_modules():
            if isinstance(m, _WNConv):
                out[name + ".weight"] = m.folded().float().cpu().contiguous()
                out[name + ".bias"] = m.bias.detach().float().cpu().contiguous()
            elif isinstance(m, _Table):
                out[name + ".weight"] = m.weight.detach().float().cpu().contiguous()
        return out

    def gen_cfg(self) -> "_lib.GenCfg":
        cfg = _lib.GenCfg()
        cfg.n_up = self.num_upsamples
        for i, (u, k) in enumerate(zip(self.rates, self.up_kernels)):
            cfg.up_rates[i], cfg.up_kernels[i] = u, k
        cfg.n_rk = self.num_kernels
        cfg.n_dil = len(self.rds[0])
        for j, (k, ds) in enumerate(zip(self.rks, self.rds)):
            if len(ds) != cfg.n_dil:
                raise NotImplementedError("resblock_dilation_sizes rows must have equal length")
            cfg.rk[j] = k
            for m, d in enumerate(ds):
                cfg.dil[j][m] = d
        cfg.c0 = self.c0
        cfg.embedding_dim = self.emb_dim
        cfg.num_embeddings = self.dict.weight.shape[0]
        cfg.n_spkr_rows = self.spkr.weight.shape[0] if self.multispkr else 0
        cfg.model_in_dim = self.in_dim
        cfg.resblock = 1 if self.resblock == "1" else 2
        cfg.has_f0 = 1 if self.f0 else 0
        cfg.has_spkr = 1 if self.multispkr else 0
        return cfg

    def _drop_handle(self):
        h = self.__dict__.get("_handle", None)
        if h is not None:
            _lib.lib().dissc_gen_destroy(h)
        self.__dict__["_handle"] = None
        self.__dict__["_ws"] = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _ensure_handle(self, device: torch.device):
        if self._handle is not None and self._handle_device == device:
            return self._handle
        self._drop_handle()
        L = _lib.lib()
        sd = self.folded_state_dict()
        arr = (_lib.Tensor * len(sd))()
        keep = []
        for i, (k, v) in enumerate(sd.items()):
            name = k.encode()
            keep.append((name, v))
            arr[i].name = name
            arr[i].data = ctypes.cast(v.data_ptr(), ctypes.POINTER(ctypes.c_float))
            arr[i].numel = v.numel()
        cfg = self.gen_cfg()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.dissc_gen_create(ctypes.byref(handle), ctypes.byref(cfg), arr, len(sd), device.index),
                       "dissc_gen_create")
        self._handle, self._handle_device = handle, device
        return handle

    @property
    def hop(self) -> int:
        n = 1
        for u in self.rates:
            n *= u
        return n

    def workspace_bytes(self, B, T, device) -> int:
        n = ctypes.c_size_t()
        _lib.check(_lib.lib().dissc_gen_workspace_bytes(self._ensure_handle(device), B, T, ctypes.byref(n)))
        return n.value

    def _run(self, code, f0, spkr, lengths, B, T, out_dtype, ws=None, out=None, extra=None):
        dev = code.device
        L = _lib.lib()
        h = self._ensure_handle(dev)
        need = self.workspace_bytes(B, T, dev)
        if ws is None:
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            ws = self._ws
        t_out = T
        for u, k in zip(self.rates, self.up_kernels):
            t_out = (t_out - 1) * u - 2 * ((k - u) // 2) + k
        if out is None:
            out = torch.empty((B, t_out), dtype=out_dtype, device=dev)
        elif out.dtype != out_dtype or out.numel() != B * t_out or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"out must be a contiguous {out_dtype} tensor of {B * t_out} elements on {dev}")
        ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if extra is not None:
                f32 = out_dtype == torch.float32
                extra = extra.to(device=dev, dtype=torch.float32).contiguous()
                _lib.check(L.dissc_gen_forward_ex(h, ptr(code), ptr(f0), ptr(spkr), ptr(extra), ptr(lengths), B, T,
                                                  ptr(out) if f32 else None, None if f32 else ptr(out), ptr(ws), ws.numel(),
                                                  stream), "dissc_gen_forward_ex")
            else:
                fn = L.dissc_gen_forward if f32_out(out_dtype) else L.dissc_gen_forward_i16
                _lib.check(fn(h, ptr(code), ptr(f0), ptr(spkr), ptr(lengths), B, T, ptr(out), ptr(ws), ws.numel(), stream),
                           "dissc_gen_forward")
        return out

    def capture_graph(self, B: int, T: int, device, int16: bool = False, varlen: bool = False):
        """CUDA-graph one (B, T) forward: the ~76 kernel launches are captured once and replayed with a single
        cudaGraphLaunch, which removes the per-launch CPU / driver latency that dominates the reference's own use case
        (batch size 1, sr/inference.py:178,247).  Returns a ``GraphedForward``: write the static input tensors
        (``.code`` int64 (B,T), ``.f0`` fp32 (B,T), ``.spkr`` int64 (B), ``.lengths`` int32 (B) if ``varlen``), call it,
        read ``.out`` ((B, hop*T) fp32 or int16).  Same kernels, same arithmetic: bit-identical to the eager call."""
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        g = GraphedForward()
        g.code = torch.zeros((B, T), dtype=torch.int64, device=dev)
        g.f0 = torch.zeros((B, T), dtype=torch.float32, device=dev) if self.f0 else None
        g.spkr = torch.zeros((B,), dtype=torch.int64, device=dev) if self.multispkr else None
        g.lengths = torch.full((B,), T, dtype=torch.int32, device=dev) if varlen else None
        dt = torch.int16 if int16 else torch.float32
        g.ws = torch.empty(self.workspace_bytes(B, T, dev), dtype=torch.uint8, device=dev)   # owned by the graph
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):   # warm-up outside the capture: handle, workspace, one-time function attributes
                self._run(g.code, g.f0, g.spkr, g.lengths, B, T, out_dtype=dt, ws=g.ws)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g.graph):
            g.out = self._run(g.code, g.f0, g.spkr, g.lengths, B, T, out_dtype=dt, ws=g.ws)
        g._owner = self   # keeps the device handle (weights) alive
        return g

    def forward_host(self, code, f0, spkr, lengths=None, out=None, int16=False, device=0):
        """End-to-end C-ABI call with HOST (ideally pinned) tensors: H2D + forward + D2H + sync."""
        out = self.forward_host_submit(0, code, f0, spkr, lengths=lengths, out=out, int16=int16, device=device)
        self.forward_host_wait(0)
        return out

    def forward_host_submit(self, slot, code, f0, spkr, lengths=None, out=None, int16=False, device=0):
        """Pipelined host entry (``dissc_gen_forward_host_submit``): enqueue H2D -> forward -> D2H for ``slot`` (0 or 1)
        and return the host output tensor WITHOUT waiting; it is valid after ``forward_host_wait(slot)``.  Submitting
        the other slot before waiting hides one batch's copy-back under the next batch's forward.  The host tensors
        of a slot must stay alive (and should be pinned) until its wait returns."""
        B, T = code.shape
        dev = torch.device("cuda", device)
        h = self._ensure_handle(dev)
        if out is None:
            out = torch.empty((B, self.hop * T), dtype=torch.int16 if int16 else torch.float32).pin_memory()
        ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        self.__dict__.setdefault("_host_keep", {})[slot] = (code, f0, spkr, lengths, out)
        _lib.check(_lib.lib().dissc_gen_forward_host_submit(
            h, slot, ptr(code), ptr(f0), ptr(spkr), ptr(lengths), B, T,
            None if int16 else ptr(out), ptr(out) if int16 else None), "dissc_gen_forward_host_submit")
        return out

    def forward_host_wait(self, slot):
        if self._handle is None:
            return
        _lib.check(_lib.lib().dissc_gen_forward_host_wait(self._handle, slot), "dissc_gen_forward_host_wait")
        self.__dict__.get("_host_keep", {}).pop(slot, None)

    def host_reserve(self, B, T, device=0):
        """Pre-size the device staging arena of the host entry points for (B, T) batches."""
        _lib.check(_lib.lib().dissc_gen_host_reserve(self._ensure_handle(torch.device("cuda", device)), B, T),
                   "dissc_gen_host_reserve")

    def check_indices(self, synchronize=True):
        """Raises ``IndexError`` if a unit / speaker id outside its embedding table reached the kernels since the last
        check.  ``nn.Embedding`` (sr/models.py:128,133) raises at the call; here the forward is asynchronous, so the
        gather stays inside the table, a device-visible flag is set and the error surfaces at this call, at the next
        forward, or when ``forward_host`` returns."""
        if self._handle is None:
            return
        if synchronize:
            torch.cuda.synchronize(self._handle_device)
        _lib.check(_lib.lib().dissc_gen_status(self._handle), "CodeGenerator")

    def set_tensor_cores(self, enable: bool, device=None) -> int:
        """Toggle the tcgen05 path (default on); returns how many stages now run on tensor cores."""
        dev = device or self._handle_device or torch.device("cuda", 0)
        h = self._ensure_handle(dev)
        _lib.check(_lib.lib().dissc_gen_set_tensor_cores(h, int(bool(enable))))
        return int(_lib.lib().dissc_gen_tensor_core_stages(h))

    def launches_per_forward(self) -> int:
        if self._handle is None:
            raise _lib.DisscError("no device handle yet: run a forward first")
        return int(_lib.lib().dissc_gen_launches_per_forward(self._handle))

    def cost(self, B, T, device=None):
        """(algorithmic FLOPs, layer-fused-model bytes) of one (B,T) forward."""
        dev = device or self._handle_device or torch.device("cuda", 0)
        fl, by = ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.lib().dissc_gen_cost(self._ensure_handle(dev), B, T, ctypes.byref(fl), ctypes.byref(by)))
        return fl.value, by.value

    def profile(self, code, f0, spkr, lengths=None):
        """Per-launch device times of one forward: list of (name, ms, flops, algorithmic bytes)."""
        B, T = code.shape
        dev = code.device
        L = _lib.lib()
        h = self._ensure_handle(dev)
        need = self.workspace_bytes(B, T, dev)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty((B, self.hop * T), dtype=torch.float32, device=dev)
        cap = 512
        names = ((ctypes.c_char * 64) * cap)()
        ms = (ctypes.c_float * cap)()
        fl = (ctypes.c_double * cap)()
        by = (ctypes.c_double * cap)()
        n = ctypes.c_int()
        ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        f0 = None if f0 is None else f0.reshape(B, T).float().contiguous()
        spkr = None if spkr is None else spkr.reshape(B).contiguous()
        torch.cuda.synchronize(dev)
        with torch.cuda.device(dev):
            _lib.check(L.dissc_gen_profile(h, ptr(code), ptr(f0), ptr(spkr), ptr(lengths), B, T, ptr(out), ptr(ws),
                                           need, names, ms, fl, by, cap, ctypes.byref(n)), "dissc_gen_profile")
        return [(names[i].value.decode(), ms[i], fl[i], by[i]) for i in range(n.value)]


def f32_out(dtype) -> bool:
    return dtype == torch.float32


class GraphedForward:
    """A captured (B, T) forward (``CodeGenerator.capture_graph``): static inputs ``code`` / ``f0`` / ``spkr`` /
    ``lengths``, static output ``out``; ``__call__`` replays the graph on the current stream and returns ``out``."""
    code = f0 = spkr = lengths = out = graph = None

    def __call__(self, code=None, f0=None, spkr=None, lengths=None):
        if code is not None:
            self.code.copy_(code.reshape(self.code.shape))
        if f0 is not None and self.f0 is not None:
            self.f0.copy_(f0.reshape(self.f0.shape))
        if spkr is not None and self.spkr is not None:
            self.spkr.copy_(spkr.reshape(self.spkr.shape))
        if lengths is not None and self.lengths is not None:
            self.lengths.copy_(lengths.reshape(self.lengths.shape))
        self.graph.replay()
        return self.out
