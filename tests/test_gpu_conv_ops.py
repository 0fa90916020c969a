"""Op-level GPU parity of the grouped implicit-GEMM convolution kernels (nef_gconv_fwd / nef_gconv_wgrad,
both the tcgen05 TF32 and the CUDA-core implementation) through the C ABI, against torch's fp32 conv1d /
autograd on the same inputs (torch is only the checker here; TF32 is disabled on its side).

Inputs are pre-rounded to TF32 (what the producing kernels of the real path do), so the only difference
between the tensor-core path and the fp32 reference is the accumulation order: tolerances are tight.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tf32(t):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


@pytest.fixture(autouse=True)
def _no_tf32_in_torch():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _set_impl(impl):
    from network import _native as N
    lib = N.init(0)
    torch.cuda.synchronize()
    lib.nef_set_conv_impl(impl)
    return lib


CASES = [
    # B, L, groups, cin_g, cout_g, taps
    (2, 40, 2, 128, 128, 7),      # tiny: one-row-tile kernel
    (3, 100, 3, 64, 128, 3),
    (2, 64, 1, 256, 128, 3),
    (5, 333, 2, 128, 64, 3),
    (4, 50, 7, 128, 64, 1),
    (64, 1250, 3, 128, 128, 7),   # enough tiles for the 4-row-tile kernel (the production configuration)
    (96, 2500, 1, 128, 64, 3),
    (200, 16, 14, 128, 128, 3),   # the z2 deflection branch shape (many groups, 16 samples)
    (16, 1250, 2, 128, 128, 3),   # k3 128 -> 128 with many row stages (the wide weight-gradient tile)
]


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("B,L,groups,cin_g,cout_g,taps", CASES)
def test_gconv_fwd_matches_conv1d(B, L, groups, cin_g, cout_g, taps, impl):
    from network import ops
    dev = torch.device("cuda:0")
    lib = _set_impl(impl)
    try:
        gen = torch.Generator(device="cpu").manual_seed(B * 1000 + L)
        x = _tf32(torch.randn(B, groups * cin_g, L, generator=gen)).to(dev)
        w = _tf32(torch.randn(groups * cout_g, cin_g, taps, generator=gen) / np.sqrt(cin_g * taps)).to(dev)
        bias = torch.randn(groups * cout_g, generator=gen).to(dev)
        res = torch.randn(B, groups * cout_g, L, generator=gen).to(dev)
        xt = ops.Cbl4(groups * cin_g, B, L, dev).from_ncl(x)
        rt = ops.Cbl4(groups * cout_g, B, L, dev).from_ncl(res)
        yt = ops.Cbl4(groups * cout_g, B, L, dev)
        wpk = ops.pack_conv_weight(w, groups)
        d = ops.conv_desc(xt, wpk, yt, groups, cin_g, cout_g, taps, relu=True, bias=bias, res=rt)
        ops.gconv_fwd(d)
        got = yt.to_ncl()
        ref = F.relu(F.conv1d(x, w, bias, padding=taps // 2, groups=groups) + res)
        err = float((got - ref).abs().max())
        assert err < 2e-4, err
        # halo rows must stay zero (the next layer's padding)
        assert float(yt.data.view(-1, 4)[: 3].abs().max()) == 0.0
    finally:
        lib.nef_set_conv_impl(1)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("B,L,groups,cin_g,cout_g,taps", CASES)
def test_gconv_dgrad_matches_autograd(B, L, groups, cin_g, cout_g, taps, impl):
    """Data gradient = the same kernel with flipped / transposed packed weights."""
    from network import ops
    if cin_g > 128:
        pytest.skip("dgrad packs at most 128 output (= forward input) channels per group")
    dev = torch.device("cuda:0")
    lib = _set_impl(impl)
    try:
        gen = torch.Generator(device="cpu").manual_seed(B * 1000 + L + 1)
        x = torch.randn(B, groups * cin_g, L, generator=gen).to(dev).requires_grad_(True)
        w = _tf32(torch.randn(groups * cout_g, cin_g, taps, generator=gen) / np.sqrt(cin_g * taps)).to(dev)
        dy = _tf32(torch.randn(B, groups * cout_g, L, generator=gen)).to(dev)
        F.conv1d(x, w, None, padding=taps // 2, groups=groups).backward(dy)
        dyt = ops.Cbl4(groups * cout_g, B, L, dev).from_ncl(dy)
        dxt = ops.Cbl4(groups * cin_g, B, L, dev)
        wpk = ops.pack_conv_weight(w, groups, dgrad=True)
        d = ops.conv_desc(dyt, wpk, dxt, groups, cout_g, cin_g, taps)
        ops.gconv_fwd(d)
        err = float((dxt.to_ncl() - x.grad).abs().max())
        assert err < 2e-4, err
    finally:
        lib.nef_set_conv_impl(1)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("B,L,groups,cin_g,cout_g,taps", CASES)
def test_gconv_wgrad_matches_autograd(B, L, groups, cin_g, cout_g, taps, impl):
    from network import ops
    dev = torch.device("cuda:0")
    lib = _set_impl(impl)
    try:
        gen = torch.Generator(device="cpu").manual_seed(B * 1000 + L + 2)
        x = _tf32(torch.randn(B, groups * cin_g, L, generator=gen)).to(dev)
        w = torch.zeros(groups * cout_g, cin_g, taps, device=dev, requires_grad=True)
        bias = torch.zeros(groups * cout_g, device=dev, requires_grad=True)
        dy = _tf32(torch.randn(B, groups * cout_g, L, generator=gen)).to(dev)
        F.conv1d(x, w, bias, padding=taps // 2, groups=groups).backward(dy)
        xt = ops.Cbl4(groups * cin_g, B, L, dev).from_ncl(x)
        dyt = ops.Cbl4(groups * cout_g, B, L, dev).from_ncl(dy)
        dw = torch.zeros_like(w)
        db = torch.zeros_like(bias)
        ops.gconv_wgrad(dyt, xt, dw, groups, cout_g, cin_g, taps, db=db)
        scale = float(w.grad.abs().max())
        assert float((dw - w.grad).abs().max()) < 5e-4 * scale
        assert float((db - bias.grad).abs().max()) < 5e-4 * float(bias.grad.abs().max())
        # gradients accumulate
        ops.gconv_wgrad(dyt, xt, dw, groups, cout_g, cin_g, taps, db=db)
        assert float((dw - 2 * w.grad).abs().max()) < 1e-3 * scale
    finally:
        lib.nef_set_conv_impl(1)


def test_bn_partial_statistics_are_deterministic():
    """Two launches over the same input write bit-identical per-tile BatchNorm partial sums."""
    from network import ops, _native as N
    dev = torch.device("cuda:0")
    lib = _set_impl(1)
    B, L, C = 32, 700, 128
    gen = torch.Generator(device="cpu").manual_seed(5)
    x = _tf32(torch.randn(B, C, L, generator=gen)).to(dev)
    w = _tf32(torch.randn(C, C, 3, generator=gen) / 20).to(dev)
    xt = ops.Cbl4(C, B, L, dev).from_ncl(x)
    yt = ops.Cbl4(C, B, L, dev)
    wpk = ops.pack_conv_weight(w, 1)
    n_rec = (xt.rows + 127) // 128
    outs = []
    for _ in range(2):
        s1 = torch.full((n_rec, C), float("nan"), device=dev)
        s2 = torch.full((n_rec, C), float("nan"), device=dev)
        d = ops.conv_desc(xt, wpk, yt, 1, C, C, 3)
        d.stat_sum, d.stat_sq = s1.data_ptr(), s2.data_ptr()
        ops.gconv_fwd(d)
        outs.append((s1.clone(), s2.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    y = yt.to_ncl()
    np.testing.assert_allclose(outs[0][0].sum(0).cpu().numpy(), y.sum((0, 2)).cpu().numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(outs[0][1].sum(0).cpu().numpy(), (y * y).sum((0, 2)).cpu().numpy(), rtol=1e-4)
