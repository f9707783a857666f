"""Layer-level parity: the fused sm_100a conv kernels (through the C ABI) vs the
oracle's ATen calls (F.conv1d / F.conv_transpose1d / F.leaky_relu on CPU)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-5  # fp32, different summation order than oneDNN; activations are O(1)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _conv_case(dev, B, Cin, Cout, T, k, d, pre, post, res, acc, div, lengths=None, seed=0):
    from dissc_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    r = torch.randn(B, Cout, T, generator=g) if res else None
    a = torch.randn(B, Cout, T, generator=g) if acc else None
    xin = x.clone()
    if lengths is not None:
        for i, n in enumerate(lengths):
            xin[i, :, n:] = 0
    y = F.leaky_relu(xin, 0.1) if pre else xin
    y = F.conv1d(y, w, b, padding=(k * d - d) // 2, dilation=d)
    if res:
        y = y + r
    if acc:
        y = a + y
    if div:
        y = y / div
    if post:
        y = F.leaky_relu(y, 0.01)
    xd = x.to(dev)
    if lengths is not None:  # garbage past the valid length must be ignored by the kernel
        for i, n in enumerate(lengths):
            xd[i, :, n:] = float("nan")
    out = torch.full((B, Cout, T), float("nan"), device=dev)
    ld = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=dev)
    rd = None if r is None else r.to(dev)
    ad = None if a is None else a.to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_conv1d_fused(_ptr(xd), _ptr(w), _ptr(b), _ptr(rd), _ptr(ad), _ptr(out), _ptr(ld), 1,
                                                  B, Cin, Cout, T, k, d, int(pre), 0.1, int(post), 0.01,
                                                  float(div), None))
    torch.cuda.synchronize()
    got = out.cpu()
    if lengths is not None:
        for i, n in enumerate(lengths):
            got[i, :, n:] = 0
            y[i, :, n:] = 0
    assert torch.isfinite(got).all()
    err = (got - y).abs().max().item()
    assert err < TOL, f"max abs err {err}"


@pytest.mark.parametrize("C", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("k,d", [(3, 1), (3, 5), (7, 3), (11, 1), (11, 5)])
def test_conv1d_resblock_shapes(cuda_device, C, k, d):
    T = {16: 2500, 32: 1300, 64: 700, 128: 300, 256: 300}[C]
    _conv_case(cuda_device, 2, C, C, T, k, d, pre=True, post=True, res=False, acc=False, div=0)


@pytest.mark.parametrize("k,d", [(1, 1), (5, 1), (5, 3), (7, 1), (7, 5), (11, 3), (3, 3)])
def test_conv1d_all_instantiations(cuda_device, k, d):
    _conv_case(cuda_device, 1, 24, 40, 333, k, d, pre=False, post=False, res=True, acc=False, div=0)


def test_conv1d_epilogue_modes(cuda_device):
    _conv_case(cuda_device, 3, 64, 64, 515, 7, 1, pre=False, post=False, res=True, acc=True, div=0)
    _conv_case(cuda_device, 3, 64, 64, 515, 7, 1, pre=False, post=True, res=True, acc=True, div=3.0)
    _conv_case(cuda_device, 1, 16, 16, 4100, 3, 1, pre=True, post=True, res=True, acc=False, div=0)


@pytest.mark.parametrize("Cin,Cout", [(1, 1), (2, 2), (5, 3), (257, 70), (8, 130)])
def test_conv1d_ragged_channels(cuda_device, Cin, Cout):
    _conv_case(cuda_device, 2, Cin, Cout, 97, 7, 1, pre=False, post=False, res=False, acc=False, div=0)


@pytest.mark.parametrize("T", [1, 2, 7, 255, 256, 257, 1025])
def test_conv1d_ragged_time(cuda_device, T):
    _conv_case(cuda_device, 2, 32, 32, T, 11, 5, pre=True, post=False, res=True, acc=False, div=0)


def test_conv1d_lengths_mask(cuda_device):
    _conv_case(cuda_device, 4, 64, 64, 600, 11, 3, pre=True, post=True, res=False, acc=False, div=0,
               lengths=[600, 1, 257, 433])
    _conv_case(cuda_device, 3, 16, 16, 3000, 7, 5, pre=True, post=True, res=False, acc=False, div=0,
               lengths=[3000, 1000, 17])


def _convt_case(dev, B, Cin, Cout, T, k, u, lengths=None, seed=0):
    from dissc_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5
    b = torch.randn(Cout, generator=g)
    xin = x.clone()
    if lengths is not None:
        for i, n in enumerate(lengths):
            xin[i, :, n:] = 0
    y = F.conv_transpose1d(xin, w, b, stride=u, padding=(k - u) // 2)
    xd = x.to(dev)
    if lengths is not None:
        for i, n in enumerate(lengths):
            xd[i, :, n:] = float("nan")
    out = torch.full(tuple(y.shape), float("nan"), device=dev)
    ld = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_conv_transpose1d(_ptr(xd), _ptr(w), _ptr(b), _ptr(out), _ptr(ld), 1, B, Cin, Cout,
                                                      T, k, u, None))
    torch.cuda.synchronize()
    got = out.cpu()
    if lengths is not None:
        for i, n in enumerate(lengths):
            got[i, :, n * u:] = 0
            y[i, :, n * u:] = 0
    assert torch.isfinite(got).all()
    err = (got - y).abs().max().item()
    assert err < TOL, f"max abs err {err}"


@pytest.mark.parametrize("Cin,Cout,k,u,T", [
    (512, 256, 11, 5, 50), (256, 128, 8, 4, 250), (128, 64, 8, 4, 1000), (64, 32, 4, 2, 4000),
    (32, 16, 4, 2, 8000), (64, 32, 16, 8, 130), (4, 2, 4, 2, 33), (10, 7, 11, 5, 1), (2, 1, 8, 4, 129)])
def test_conv_transpose1d(cuda_device, Cin, Cout, k, u, T):
    _convt_case(cuda_device, 2, Cin, Cout, T, k, u)


def test_conv_transpose1d_lengths(cuda_device):
    _convt_case(cuda_device, 3, 64, 32, 300, 11, 5, lengths=[300, 1, 129])
    _convt_case(cuda_device, 3, 32, 16, 700, 4, 2, lengths=[700, 512, 13])


def test_unsupported_geometry_fails_loudly(cuda_device):
    from dissc_b200 import _lib
    x = torch.zeros(1, 4, 16, device=cuda_device)
    w = torch.zeros(4, 4, 9)
    out = torch.zeros(1, 4, 16, device=cuda_device)
    rc = _lib.lib().dissc_conv1d_fused(_ptr(x), _ptr(w), None, None, None, _ptr(out), None, 1, 1, 4, 4, 16, 9, 1, 0,
                                        0.0, 0, 0.0, 0.0, None)
    assert rc == -2
    assert b"kernel_size=9" in _lib.lib().dissc_last_error()


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05, split-fp16) twin of the fused conv
# ---------------------------------------------------------------------------------------------
def _tc_case(dev, B, C, T, k, d, pre, post, res, acc, div, lengths=None, seed=0, wscale=1.0, Cin=None, tol=2e-5,
             single_acc=False, xscale=1.0):
    """xscale multiplies every activation-like input (x, bias, residual, accumulator): the op is positively homogeneous,
    so the expected output scales by the same factor and the tolerance is relative to it."""
    from dissc_b200 import _lib
    # opt-in fast mode for 256-column GEMMs: one 256-column chunk, all three MMAs into ONE accumulator (the default is
    # two 128-column chunks with separate main / cross accumulators)
    _lib.lib().dissc_tc_set_single_accumulator(int(single_acc))
    _lib.lib().dissc_tc_set_tuning(2, 0 if single_acc else 1)
    g = torch.Generator().manual_seed(seed)
    Cin = Cin or C
    x = xscale * torch.randn(B, Cin, T, generator=g)
    w = wscale * torch.randn(C, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = xscale * torch.randn(C, generator=g)
    r = xscale * torch.randn(B, C, T, generator=g) if res else None
    a = xscale * torch.randn(B, C, T, generator=g) if acc else None
    xin = x.clone()
    if lengths is not None:
        for i, n in enumerate(lengths):
            xin[i, :, n:] = 0
    y = F.leaky_relu(xin, 0.1) if pre else xin
    y = F.conv1d(y.double(), w.double(), b.double(), padding=(k * d - d) // 2, dilation=d).float()
    if res:
        y = y + r
    if acc:
        y = a + y
    if div:
        y = y / div
    raw = y.clone()
    if post:
        y = F.leaky_relu(y, 0.01)
    xd = x.to(dev)
    if lengths is not None:
        for i, n in enumerate(lengths):
            xd[i, :, n:] = float("nan")
    outs = [torch.full((B, C, T), float("nan"), device=dev) for _ in range(3)]
    ld = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=dev)
    rd = None if r is None else r.to(dev)
    ad = None if a is None else a.to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_conv1d_tc(_ptr(xd), _ptr(w), _ptr(b), _ptr(rd), _ptr(ad), _ptr(outs[0]),
                                               _ptr(outs[1]), _ptr(outs[2]), _ptr(ld), 1, B, Cin, C, T, k, d, int(pre), 0.1,
                                               int(post), 0.01, float(div), None))
    torch.cuda.synchronize()
    scale = max(1.0, wscale) * xscale
    _lib.lib().dissc_tc_set_single_accumulator(0)  # library defaults
    _lib.lib().dissc_tc_set_tuning(2, 1)
    for name, got, want in (("plain", outs[0].cpu(), y), ("raw", outs[1].cpu(), raw), ("planes", outs[2].cpu(), y)):
        got = got.clone()
        want = want.clone()
        if lengths is not None:
            for i, n in enumerate(lengths):
                if name == "planes":
                    assert torch.all(got[i, :, n:] == 0), "rows past the valid length must be stored as zeros"
                got[i, :, n:] = 0
                want[i, :, n:] = 0
        assert torch.isfinite(got).all(), name
        err = (got - want).abs().max().item()
        assert err < tol * scale, f"{name}: max abs err {err}"


@pytest.mark.parametrize("C", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("k,d", [(3, 1), (3, 5), (7, 3), (11, 1), (11, 5)])
def test_tc_conv_resblock_shapes(cuda_device, C, k, d):
    T = {16: 2500, 32: 1300, 64: 700, 128: 300, 256: 300}[C]
    _tc_case(cuda_device, 2, C, T, k, d, pre=True, post=True, res=False, acc=False, div=0)


@pytest.mark.parametrize("k,d", [(3, 1), (7, 3), (11, 5)])
def test_tc_conv_single_accumulator_n256(cuda_device, k, d):
    # N=256 OPT-IN fast mode: all three split-precision MMAs into ONE TMEM accumulator (the tensor core truncates the
    # accumulator after every MMA, so the error grows with 3 * Cin*k/16 accumulations) -- looser per-layer tolerance; end
    # to end it costs 2.5x the waveform error (1.0e-4 instead of 4e-5 on the `hot` weights), which is why it is not the
    # default.  The default (two 128-column chunks, dual accumulators) holds the tight tolerance:
    _tc_case(cuda_device, 2, 256, 300, k, d, pre=True, post=True, res=True, acc=False, div=0, tol=1e-4, single_acc=True)
    _tc_case(cuda_device, 2, 256, 300, k, d, pre=True, post=True, res=True, acc=False, div=0)


def test_tc_conv_epilogue_modes(cuda_device):
    _tc_case(cuda_device, 3, 64, 515, 7, 1, pre=False, post=False, res=True, acc=True, div=0)
    _tc_case(cuda_device, 3, 64, 515, 7, 1, pre=False, post=True, res=True, acc=True, div=3.0)
    _tc_case(cuda_device, 1, 16, 4100, 3, 1, pre=True, post=True, res=True, acc=False, div=0)


@pytest.mark.parametrize("T", [1, 2, 127, 128, 129, 1025])
def test_tc_conv_ragged_time(cuda_device, T):
    _tc_case(cuda_device, 2, 32, T, 11, 5, pre=True, post=False, res=True, acc=False, div=0)


def test_tc_conv_many_tiles_persistent(cuda_device):
    # more tiles than SMs: every CTA loops, both TMEM accumulators and all pipeline phases wrap
    _tc_case(cuda_device, 8, 128, 128 * 45, 7, 3, pre=True, post=True, res=True, acc=False, div=0)
    _tc_case(cuda_device, 8, 16, 128 * 90, 11, 5, pre=True, post=True, res=False, acc=False, div=0)


def test_tc_conv_lengths_mask(cuda_device):
    _tc_case(cuda_device, 4, 64, 600, 11, 3, pre=True, post=True, res=False, acc=False, div=0,
             lengths=[600, 1, 257, 433])


@pytest.mark.parametrize("wscale", [1e-3, 1.0, 300.0])
def test_tc_conv_weight_scaling(cuda_device, wscale):
    # the power-of-two pre-scale keeps the fp16 split accurate whatever the weight magnitude
    _tc_case(cuda_device, 1, 64, 300, 7, 1, pre=False, post=False, res=False, acc=False, div=0, wscale=wscale)


# Dynamic range of the split-fp16 ACTIVATION planes (the weights carry their own power-of-two scale): fp16 is normal
# on [6.1e-5, 65504] and the `lo` plane needs |x| >= 2^-3 for all of its 11 bits, so an unscaled plane loses relative
# precision once a whole tensor sits below ~0.1 and saturates above 6.5e4.  The layer entry points measure nothing:
# they scale the planes by the power of two `plane_scale_for(absmax)` of the tensor they pack (the model does the same
# per layer at load time, DESIGN.md "Activation scale"), so every magnitude below keeps the fp32-class relative error.
@pytest.mark.parametrize("xscale", [1e-4, 1e-2, 1.0, 1e2, 1e3, 3e4])
def test_tc_conv_activation_scale_sweep(cuda_device, xscale):
    _tc_case(cuda_device, 2, 64, 515, 7, 3, pre=True, post=True, res=True, acc=True, div=3.0, xscale=xscale)
    _tc_case(cuda_device, 1, 256, 300, 11, 1, pre=True, post=True, res=True, acc=False, div=0, xscale=xscale)
    _tc_case(cuda_device, 1, 256, 300, 11, 1, pre=True, post=True, res=True, acc=False, div=0, xscale=xscale, tol=1e-4,
             single_acc=True)
    _tc_case(cuda_device, 2, 16, 2500, 3, 1, pre=True, post=False, res=False, acc=False, div=0, xscale=xscale)


@pytest.mark.parametrize("Cin,Cout", [(257, 512), (17, 32), (40, 16), (272, 256)])
def test_tc_conv_cin_neq_cout_chunked(cuda_device, Cin, Cout):
    # conv_pre geometry: ragged Cin (zero-padded to 16), Cout > 256 split into 256-column chunks
    _tc_case(cuda_device, 2, Cout, 300, 7, 1, pre=False, post=True, res=False, acc=False, div=0, Cin=Cin)


def _tc_convt_case(dev, B, Cin, Cout, k, u, T, lengths=None, seed=0):
    from dissc_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, k, generator=g) / (Cin * k / u) ** 0.5
    b = torch.randn(Cout, generator=g)
    xin = x.clone()
    if lengths is not None:
        for i, n in enumerate(lengths):
            xin[i, :, n:] = 0
    pad = (k - u) // 2
    want = F.conv_transpose1d(xin.double(), w.double(), b.double(), stride=u, padding=pad).float()
    Tout = want.shape[-1]
    xd = x.to(dev)
    if lengths is not None:
        for i, n in enumerate(lengths):
            xd[i, :, n:] = float("nan")
    raw = torch.full((B, Cout, Tout), float("nan"), device=dev)
    pl = torch.full((B, Cout, Tout), float("nan"), device=dev)
    ld = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=dev)
    _lib.lib().dissc_tc_set_single_accumulator(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_conv_transpose1d_tc(_ptr(xd), _ptr(w), _ptr(b), _ptr(raw), _ptr(pl), _ptr(ld), 1,
                                                         B, Cin, Cout, T, k, u, 0.1, None))
    torch.cuda.synchronize()
    raw, pl = raw.cpu(), pl.cpu()
    wantp = F.leaky_relu(want, 0.1)
    if lengths is not None:
        for i, n in enumerate(lengths):
            assert torch.all(pl[i, :, n * u:] == 0), "rows past the valid length must be stored as zeros"
            raw[i, :, n * u:] = 0
            want[i, :, n * u:] = 0
            wantp[i, :, n * u:] = 0
    assert torch.isfinite(raw).all() and torch.isfinite(pl).all()
    assert (raw - want).abs().max().item() < 2e-5
    assert (pl - wantp).abs().max().item() < 2e-5


@pytest.mark.parametrize("Cin,Cout,k,u,T", [
    (512, 256, 11, 5, 300), (256, 128, 8, 4, 640), (128, 64, 8, 4, 1500), (64, 32, 4, 2, 3000), (32, 16, 4, 2, 5000),
    (32, 16, 16, 8, 257), (32, 256, 11, 5, 1), (64, 64, 8, 4, 128), (48, 16, 4, 2, 129)])
def test_tc_conv_transpose1d(cuda_device, Cin, Cout, k, u, T):
    _tc_convt_case(cuda_device, 2, Cin, Cout, k, u, T)


def test_tc_conv_transpose1d_lengths(cuda_device):
    _tc_convt_case(cuda_device, 3, 64, 32, 8, 4, 300, lengths=[300, 1, 129])
    _tc_convt_case(cuda_device, 2, 512, 256, 11, 5, 128, lengths=[128, 77])


# ---------------------------------------------------------------------------------------------
# fused ResBlock pair (resblock_tc.cuh)
# ---------------------------------------------------------------------------------------------
def _pair_case(dev, B, C, T, k, d, acc=False, div=0.0, lengths=None, seed=0, xscale=1.0):
    from dissc_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = xscale * torch.randn(B, C, T, generator=g)
    w1 = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    w2 = 0.5 * torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b1, b2 = xscale * torch.randn(C, generator=g), xscale * torch.randn(C, generator=g)
    a = xscale * torch.randn(B, C, T, generator=g) if acc else None
    outs = []
    for b in range(B):   # reference semantics: every utterance alone, unpadded
        n = T if lengths is None else lengths[b]
        xb = x[b:b + 1, :, :n].double()
        xt = F.conv1d(F.leaky_relu(xb, 0.1), w1.double(), b1.double(), padding=(k * d - d) // 2, dilation=d)
        y = F.conv1d(F.leaky_relu(xt, 0.1), w2.double(), b2.double(), padding=(k - 1) // 2) + xb
        if acc:
            y = a[b:b + 1, :, :n].double() + y
        if div:
            y = y / div
        full = torch.zeros(1, C, T, dtype=torch.float64)
        full[:, :, :n] = y
        outs.append(full)
    want = torch.cat(outs).float()
    xd = x.to(dev)
    if lengths is not None:
        for i, n in enumerate(lengths):
            xd[i, :, n:] = float("nan")
    raw = torch.full((B, C, T), float("nan"), device=dev)
    pl = torch.full((B, C, T), float("nan"), device=dev)
    ld = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=dev)
    ad = None if a is None else a.to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dissc_resblock_pair_tc(_ptr(xd), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(ad), _ptr(raw),
                                                     _ptr(pl), _ptr(ld), 1, B, C, T, k, d, float(div), 0.1, None))
    torch.cuda.synchronize()
    raw, pl = raw.cpu(), pl.cpu()
    if lengths is not None:
        for i, n in enumerate(lengths):
            assert torch.all(pl[i, :, n:] == 0), "rows past the valid length must be stored as zeros"
            raw[i, :, n:] = 0
    assert torch.isfinite(raw).all() and torch.isfinite(pl).all()
    assert (raw - want).abs().max().item() < 3e-5 * xscale
    assert (pl - F.leaky_relu(want, 0.1)).abs().max().item() < 3e-5 * xscale


@pytest.mark.parametrize("C", [16, 32, 64])
@pytest.mark.parametrize("k,d", [(3, 1), (3, 3), (3, 5), (7, 1), (7, 3), (7, 5), (11, 1), (11, 3), (11, 5)])
def test_pair_resblock_shapes(cuda_device, C, k, d):
    _pair_case(cuda_device, 2, C, 1000, k, d)


@pytest.mark.parametrize("xscale", [1e-4, 1e-2, 1e2, 1e3, 3e4])
@pytest.mark.parametrize("C", [16, 32, 64])
def test_pair_activation_scale_sweep(cuda_device, C, xscale):
    _pair_case(cuda_device, 2, C, 700, 7, 3, acc=True, div=3.0, xscale=xscale)
    _pair_case(cuda_device, 2, C, 700, 11, 1, xscale=xscale)


@pytest.mark.parametrize("T", [1, 2, 117, 118, 119, 128, 1025])
def test_pair_ragged_time(cuda_device, T):
    _pair_case(cuda_device, 2, 32, T, 11, 5)
    _pair_case(cuda_device, 3, 16, T, 3, 1, acc=True, div=3.0)


def test_pair_many_tiles_and_lengths(cuda_device):
    # far more tiles than SMs: both worker groups wrap all their pipeline phases many times
    _pair_case(cuda_device, 6, 16, 126 * 160, 3, 1, acc=True)
    _pair_case(cuda_device, 4, 32, 118 * 90 + 5, 11, 5, acc=True, div=3.0)
    _pair_case(cuda_device, 4, 32, 700, 7, 3, lengths=[700, 1, 257, 433])


# C = 64: planes in / planes out, streamed weights, residual rebuilt from the planes (resblock64_tc.cuh)
@pytest.mark.parametrize("T", [1, 2, 117, 118, 119, 128, 1025])
def test_pair64_ragged_time(cuda_device, T):
    _pair_case(cuda_device, 2, 64, T, 11, 5)
    _pair_case(cuda_device, 3, 64, T, 3, 1, acc=True, div=3.0)


def test_pair64_many_tiles_and_lengths(cuda_device):
    # far more tiles than SMs: both worker groups and the weight ring wrap their phases many times
    _pair_case(cuda_device, 6, 64, 126 * 160, 3, 1, acc=True)
    _pair_case(cuda_device, 4, 64, 118 * 90 + 5, 11, 5, acc=True, div=3.0)
    _pair_case(cuda_device, 4, 64, 122 * 77 + 1, 7, 3)
    _pair_case(cuda_device, 4, 64, 700, 7, 3, lengths=[700, 1, 257, 433])
