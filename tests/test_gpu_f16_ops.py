"""GPU, op level: the fp16-operand kernels against conv1d / autograd, through the public C ABI (nef_ncl_to_h8,
nef_pack_weights flag bit 2, NefConvTerm.x_f16, NefConvDesc.res16 / y16, nef_gconv_wgrad_f16).

nef_gconv_wgrad_f16 (csrc/nef_wgrad_f16.cu): weight gradient from fp16 operand copies read MN-major as the bulk copy lands
them -- no re-tile pass; and the production convolution kernel with fp16 operand copies in the forward and the data-gradient
direction, with the residual read from an fp16 copy and a loss-scaled fp16 output copy.  First run on hardware in round 2
(profiles/r02_wgrad_f16_first_hardware_run.txt)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _no_tf32:
    def __enter__(self):
        self.a, self.b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *e):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.a, self.b


@pytest.mark.parametrize("B,L,groups,cin_g,taps", [(2, 122, 1, 64, 1), (3, 250, 2, 128, 3), (4, 500, 2, 128, 7), (1, 40, 3, 64, 7),
                                                   (16, 1250, 2, 128, 7), (64, 1250, 12, 128, 7), (64, 1250, 12, 64, 3)])
def test_wgrad_f16_matches_autograd(B, L, groups, cin_g, taps):
    from network import _native as N, ops
    dev = torch.device("cuda:0")
    N.init(0)
    cout_g = 128
    with _no_tf32():
        gen = torch.Generator(device="cpu").manual_seed(B * 1000 + L + taps)
        x = torch.randn(B, groups * cin_g, L, generator=gen).half().float().to(dev)      # exactly representable in fp16
        dy = torch.randn(B, groups * cout_g, L, generator=gen).half().float().to(dev)
        w = torch.zeros(groups * cout_g, cin_g, taps, device=dev, requires_grad=True)
        F.conv1d(x, w, None, padding=taps // 2, groups=groups).backward(dy)
        S = 4.0                                                                            # a "loss scale" on the dY copy ...
        x16 = ops.H8(groups * cin_g, B, L, dev).from_ncl(x)
        dy16 = ops.H8(groups * cout_g, B, L, dev).from_ncl(dy, scale=S)
        scale = torch.tensor([0.25 / S], device=dev)                                     # ... removed by the device scalar
        dw = torch.zeros_like(w)
        ops.gconv_wgrad_f16(dy16, x16, dw, groups, cout_g, cin_g, taps, out_scale=scale)
        torch.cuda.synchronize()
        ref = w.grad * 0.25
        tol = 5e-4 * float(ref.abs().max())
        err = float((dw - ref).abs().max())
        print("wgrad_f16 B%d L%d g%d cin%d k%d: max err %.3e (bar %.3e)" % (B, L, groups, cin_g, taps, err, tol))
        assert err < tol
        # cross-check against the production TF32 kernel on the same tensors, and accumulation (+=)
        dw2 = torch.zeros_like(w)
        xt = ops.Cbl4(groups * cin_g, B, L, dev).from_ncl(x)
        dyt = ops.Cbl4(groups * cout_g, B, L, dev).from_ncl(dy)
        ops.gconv_wgrad(dyt, xt, dw2, groups, cout_g, cin_g, taps)
        assert float((dw - 0.25 * dw2).abs().max()) < 2 * tol
        ops.gconv_wgrad_f16(dy16, x16, dw, groups, cout_g, cin_g, taps, out_scale=scale)
        assert float((dw - 2 * ref).abs().max()) < 2 * tol


@pytest.mark.parametrize("B,L,groups,cin_g,cout_g,taps", [(3, 100, 2, 128, 128, 7), (4, 333, 3, 64, 128, 3), (16, 1250, 2, 128, 128, 7),
                                                          (40, 1250, 12, 128, 128, 7)])
def test_fp16_operand_conv_and_its_data_gradient(B, L, groups, cin_g, cout_g, taps):
    """The production conv kernel with fp16 operand copies (NefConvTerm.x_f16) at op level -- forward, and the same kernel in the
    data-gradient direction (flipped / transposed weights, loss-scaled dY16 as the operand, accumulators scaled back), which is
    the route the fp16 backward takes.  The last shape takes the persistent kernel."""
    from network import _native as N, ops
    dev = torch.device("cuda:0")
    N.init(0)
    with _no_tf32():
        gen = torch.Generator(device="cpu").manual_seed(B + L + taps)
        x = torch.randn(B, groups * cin_g, L, generator=gen).half().float().to(dev).requires_grad_(True)
        w = (torch.randn(groups * cout_g, cin_g, taps, generator=gen) * 0.05).half().float().to(dev)
        dy = torch.randn(B, groups * cout_g, L, generator=gen).half().float().to(dev)
        yref = F.conv1d(x, w, None, padding=taps // 2, groups=groups)
        yref.backward(dy)
        xt = ops.Cbl4(groups * cin_g, B, L, dev)
        yt = ops.Cbl4(groups * cout_g, B, L, dev)
        # forward
        x16 = ops.H8(groups * cin_g, B, L, dev).from_ncl(x.detach())
        d = ops.conv_desc(xt, w, yt, groups, cin_g, cout_g, taps)
        ops.use_f16_operand(d, x16, ops.pack_conv_weight(w, groups, f16=True), cin_g)
        ops.gconv_fwd(d)
        y = yt.to_ncl()
        tol = 5e-4 * float(yref.detach().abs().max())
        print("fp16 conv fwd: max err %.3e (bar %.3e)" % (float((y - yref.detach()).abs().max()), tol))
        assert float((y - yref.detach()).abs().max()) < tol
        # data gradient: operand = S * dY in fp16, accumulators * 1/S
        S = 8.0
        dy16 = ops.H8(groups * cout_g, B, L, dev).from_ncl(dy, scale=S)
        inv = torch.tensor([1.0 / S], device=dev)
        d2 = ops.conv_desc(yt, w, xt, groups, cout_g, cin_g, taps)
        ops.use_f16_operand(d2, dy16, ops.pack_conv_weight(w, groups, dgrad=True, f16=True), cout_g)
        d2.acc_scale = inv.data_ptr()
        ops.gconv_fwd(d2)
        dx = xt.to_ncl()
        tol = 5e-4 * float(x.grad.abs().max())
        print("fp16 conv dgrad: max err %.3e (bar %.3e)" % (float((dx - x.grad).abs().max()), tol))
        assert float((dx - x.grad).abs().max()) < tol


def test_residual_and_output_as_fp16_copies():
    """NefConvDesc.res16 / y16 / y = NULL: y16 = fp16(S * (conv(x16) / S' + r16 * rs)) with nothing stored in fp32."""
    from network import _native as N, ops
    dev = torch.device("cuda:0")
    N.init(0)
    B, L, groups, ch, taps = 40, 1250, 12, 128, 7       # persistent kernel
    with _no_tf32():
        gen = torch.Generator(device="cpu").manual_seed(7)
        x = torch.randn(B, groups * ch, L, generator=gen).half().float().to(dev)
        r = torch.randn(B, groups * ch, L, generator=gen).half().float().to(dev)
        w = (torch.randn(groups * ch, ch, taps, generator=gen) * 0.05).half().float().to(dev)
        ref = F.relu(F.conv1d(x, w, None, padding=taps // 2, groups=groups) * 0.5 + r * 0.25) * 2.0
        xt = ops.Cbl4(groups * ch, B, L, dev)
        x16 = ops.H8(groups * ch, B, L, dev).from_ncl(x)
        r16 = ops.H8(groups * ch, B, L, dev).from_ncl(r)
        y16 = ops.H8(groups * ch, B, L, dev)
        sc = torch.tensor([0.5, 0.25, 2.0], device=dev)
        d = ops.conv_desc(xt, w, xt, groups, ch, ch, taps, relu=True)
        ops.use_f16_operand(d, x16, ops.pack_conv_weight(w, groups, f16=True), ch)
        d.y = None
        d.y16, d.acc_scale, d.y16_scale = y16.ptr, sc.data_ptr(), sc.data_ptr() + 8
        d.res16, d.res_cstride, d.res_c4_off, d.res_c4_gstride, d.res16_scale = r16.ptr, r16.rows, 0, ch // 4, sc.data_ptr() + 4
        ops.gconv_fwd(d)
        got = y16.to_ncl()
        tol = 1.5e-3 * float(ref.abs().max())            # one fp16 rounding of the stored result
        err = float((got - ref).abs().max())
        print("res16 / y16-only: max err %.3e (bar %.3e)" % (err, tol))
        assert err < tol and float(xt.data.abs().max()) == 0.0   # nothing was stored in fp32
