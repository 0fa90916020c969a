"""GPU, op level: the fp16-operand kernels against conv1d / autograd.

nef_gconv_wgrad_f16 (csrc/nef_wgrad_f16.cu): weight gradient from fp16 operand copies read MN-major as the bulk copy lands
them -- no re-tile pass; and the production convolution kernel with fp16 operand copies (NefConvTerm.x_f16) in the forward and
the data-gradient direction.  First run on hardware in round 2 (profiles/r02_wgrad_f16_first_hardware_run.txt)."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _half8(t):
    """CBL4 fp32 tensor (ops.Cbl4) -> fp16 copy `half8 [C/8][rows]` with the same guard rows on both sides."""
    from network import _native as N
    c4 = t.C // 4
    v = t.data.view(c4 // 2, 2, t.rows, 4).permute(0, 2, 1, 3).contiguous().view(-1).half()
    g = N.GUARD_ROWS * 8
    buf = torch.zeros(v.numel() + 2 * g, dtype=torch.float16, device=v.device)
    buf[g:g + v.numel()] = v
    return buf, buf[g:]


@pytest.mark.parametrize("B,L,groups,cin_g,taps", [(2, 122, 1, 64, 1), (3, 250, 2, 128, 3), (4, 500, 2, 128, 7), (1, 40, 3, 64, 7),
                                                   (16, 1250, 2, 128, 7), (64, 1250, 12, 128, 7), (64, 1250, 12, 64, 3)])
def test_wgrad_f16_matches_autograd(B, L, groups, cin_g, taps):
    from network import _native as N, ops
    dev = torch.device("cuda:0")
    lib = N.init(0)
    fn = lib.nef_gconv_wgrad_f16
    cout_g = 128
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gen = torch.Generator(device="cpu").manual_seed(B * 1000 + L + taps)
        x = torch.randn(B, groups * cin_g, L, generator=gen).half().float().to(dev)      # exactly representable in fp16
        dy = torch.randn(B, groups * cout_g, L, generator=gen).half().float().to(dev)
        w = torch.zeros(groups * cout_g, cin_g, taps, device=dev, requires_grad=True)
        F.conv1d(x, w, None, padding=taps // 2, groups=groups).backward(dy)
        xt = ops.Cbl4(groups * cin_g, B, L, dev).from_ncl(x)
        dyt = ops.Cbl4(groups * cout_g, B, L, dev).from_ncl(dy)
        xkeep, x16 = _half8(xt)
        ykeep, dy16 = _half8(dyt)
        dw = torch.zeros_like(w)
        d = N.NefWgradDesc()
        d.dy, d.dy_cstride, d.dy_c4_off, d.dy_c4_gstride = dyt.ptr, dyt.rows, 0, cout_g // 4
        d.x, d.x_cstride, d.x_c4_off, d.x_c4_gstride = xt.ptr, xt.rows, 0, cin_g // 4
        d.cout_g, d.cin_g, d.groups, d.taps, d.tap_off = cout_g, cin_g, groups, taps, -(taps // 2)
        d.rows = dyt.rows
        d.dw, d.sg, d.sm, d.sn, d.st = dw.data_ptr(), cout_g * cin_g * taps, cin_g * taps, taps, 1
        scale = 0.25
        scale_dev = torch.tensor([scale], device=dev)   # out_scale is a DEVICE scalar (the inverse loss scale)
        N.check(fn(C.byref(d), C.c_void_p(dy16.data_ptr()), C.c_void_p(x16.data_ptr()), N.ptr(scale_dev), N.stream_ptr()),
                "nef_gconv_wgrad_f16")
        torch.cuda.synchronize()
        ref = w.grad * scale
        tol = 5e-4 * float(ref.abs().max())
        err = float((dw - ref).abs().max())
        print("wgrad_f16 B%d L%d g%d cin%d k%d: max err %.3e (bar %.3e)" % (B, L, groups, cin_g, taps, err, tol))
        assert err < tol
        # cross-check against the production TF32 kernel on the same tensors, and accumulation (+=)
        dw2 = torch.zeros_like(w)
        ops.gconv_wgrad(dyt, xt, dw2, groups, cout_g, cin_g, taps)
        assert float((dw - scale * dw2).abs().max()) < 2 * tol
        N.check(fn(C.byref(d), C.c_void_p(dy16.data_ptr()), C.c_void_p(x16.data_ptr()), N.ptr(scale_dev), N.stream_ptr()),
                "nef_gconv_wgrad_f16")
        assert float((dw - 2 * ref).abs().max()) < 2 * tol
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _pack_f16(w, groups):
    """Conv1d weight (groups*N, K, taps) -> the fp16 operand packing of the plan's batch packer (flag bit 2):
    [g][tap][K/64][8][N][8 halves], a 16-byte slot = 8 consecutive input channels of one output channel."""
    GN, K, taps = w.shape
    n = GN // groups
    v = w.view(groups, n, K // 64, 8, 8, taps).permute(0, 5, 2, 3, 1, 4).contiguous()
    return v.half().view(-1)


@pytest.mark.parametrize("B,L,groups,cin_g,cout_g,taps", [(3, 100, 2, 128, 128, 7), (4, 333, 3, 64, 128, 3), (16, 1250, 2, 128, 128, 7)])
def test_fp16_operand_conv_and_its_data_gradient(B, L, groups, cin_g, cout_g, taps):
    """The production conv kernel with fp16 operand copies (NefConvTerm.x_f16) at op level -- forward, and the same kernel in the
    data-gradient direction (flipped / transposed weights, dY16 as the operand), which is the route the fp16 backward takes."""
    from network import _native as N, ops
    dev = torch.device("cuda:0")
    N.init(0)
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gen = torch.Generator(device="cpu").manual_seed(B + L + taps)
        x = torch.randn(B, groups * cin_g, L, generator=gen).half().float().to(dev).requires_grad_(True)
        w = (torch.randn(groups * cout_g, cin_g, taps, generator=gen) * 0.05).half().float().to(dev)
        dy = torch.randn(B, groups * cout_g, L, generator=gen).half().float().to(dev)
        yref = F.conv1d(x, w, None, padding=taps // 2, groups=groups)
        yref.backward(dy)

        def run(inp, weight, cin, cout):
            xt = ops.Cbl4(groups * cin, B, L, dev).from_ncl(inp)
            keep, x16 = _half8(xt)
            w16 = _pack_f16(weight, groups)
            yt = ops.Cbl4(groups * cout, B, L, dev)
            d = ops.conv_desc(xt, w16, yt, groups, cin, cout, taps)
            t = d.term[0]
            t.x, t.x_c4_off, t.x_c4_gstride, t.cin_g, t.x_f16 = x16.data_ptr(), 0, cin // 8, cin // 2, 1
            ops.gconv_fwd(d)
            torch.cuda.synchronize()
            return yt.to_ncl()

        y = run(x.detach(), w, cin_g, cout_g)
        tol = 5e-4 * float(yref.abs().max())
        print("fp16 conv fwd: max err %.3e (bar %.3e)" % (float((y - yref).abs().max()), tol))
        assert float((y - yref).abs().max()) < tol
        # data gradient = the same convolution of dY with wd[g*cin + n, m, t] = w[g*cout + m, n, taps-1-t]
        wd = w.view(groups, cout_g, cin_g, taps).flip(3).permute(0, 2, 1, 3).reshape(groups * cin_g, cout_g, taps).contiguous()
        dx = run(dy, wd, cout_g, cin_g)
        tol = 5e-4 * float(x.grad.abs().max())
        print("fp16 conv dgrad: max err %.3e (bar %.3e)" % (float((dx - x.grad).abs().max()), tol))
        assert float((dx - x.grad).abs().max()) < tol
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
