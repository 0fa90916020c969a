"""GPU: the tensor-core stem (csrc/nef_stem_tc.cu) at op level -- forward (conv k15 s2 p7 grouped -> ReLU -> MaxPool(3,2,1),
resnet_1d.py:102-105 + encoder.py:35-38) and weight gradient against float64 torch on the same inputs.

Forward: the split-precision fp16 MMAs must be fp32-accurate, i.e. the stored fp16 value may differ from fp16(reference) by
at most one fp16 ulp, and on few elements; the max-pool selection codes must equal the reference's wherever the two best
conv positions of a window are not within rounding distance of each other.  Weight gradient: evaluated on the device's own
selection codes (a linear map once they are fixed) it must match to the 11-bit rounding of the routed gradient."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    from network import _native as N, ops
    return N, ops, N.init(0), torch.device("cuda:0")


def _conv64(x, w, G):
    return F.conv1d(x.double(), w.double(), stride=2, padding=7, groups=G)


def _run_fwd(N, ops, lib, dev, B, G, L, seed, amp=1.0):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, G, L, generator=gen) * amp).to(dev)
    w = (torch.randn(128 * G, 1, 15, generator=gen) * 0.05).to(dev)
    y = ops.H8(128 * G, B, L // 4, dev)
    codes = torch.zeros(32 * G * y.rows + 2 * N.GUARD_ROWS, dtype=torch.int32, device=dev)
    cptr = codes.data_ptr() + 4 * N.GUARD_ROWS
    N.check(lib.nef_stem_tc_fwd(N.ptr(x), N.ptr(w), C.c_void_p(y.ptr), C.c_void_p(cptr), B, G, L, N.stream_ptr()), "nef_stem_tc_fwd")
    # codes: one uint32 per (4-channel chunk, row): byte k = channel 4 c4 + k
    cw = codes[N.GUARD_ROWS:N.GUARD_ROWS + 32 * G * y.rows].view(32 * G, B, y.Lp)[:, :, N.HALO:N.HALO + L // 4]
    cb = torch.stack([(cw >> (8 * k)) & 0xff for k in range(4)], dim=1)              # (32G, 4, B, L4)
    code = cb.reshape(128 * G, B, L // 4).permute(1, 0, 2).contiguous().long()        # (B, 128G, L4)
    return x, w, y, codes, cptr, code


@pytest.mark.parametrize("B,G,L,amp", [(2, 1, 512, 1.0), (3, 3, 1000, 1.0), (2, 12, 5000, 1.0), (2, 2, 2048, 0.05), (1, 3, 16, 1.0)])
def test_stem_tc_forward(B, G, L, amp):
    N, ops, lib, dev = _setup()
    x, w, y, _, _, code = _run_fwd(N, ops, lib, dev, B, G, L, 11 + L, amp)
    conv = _conv64(x, w, G)                                                           # (B, 128G, L2)
    ref = F.max_pool1d(F.relu(conv), 3, 2, 1)
    got = y.to_ncl().double()
    want = ref.float().half().double()
    # one fp16 ulp of the stored value, plus the fp32-level absolute error of the convolution itself (a value the ReLU clips in
    # the reference may come out as +1e-8 here)
    tol = want.abs() * 2.0 ** -10 + 2e-6 * float(conv.abs().max())
    assert bool(((got - want).abs() <= tol).all()), "more than one fp16 ulp from the float64 reference"
    frac = float((got != want).double().mean())
    rel = float((got - want).norm() / want.norm())
    print("stem_tc_fwd B%d G%d L%d amp %g: %.4f %% of the stored values differ by one ulp, rel-L2 %.2e" % (B, G, L, amp, 100 * frac, rel))
    assert frac < 0.03 and rel < 1e-4
    # selections: windows (2j-1, 2j, 2j+1), first maximum wins, 3 = clipped
    L4 = L // 4
    pad = F.pad(conv, (1, 0), value=float("-inf"))                                    # position -1
    win = torch.stack([pad[:, :, 0:2 * L4:2], pad[:, :, 1:2 * L4 + 1:2], pad[:, :, 2:2 * L4 + 2:2]], dim=-1)   # (B, C, L4, 3)
    best = win.argmax(dim=-1)
    top2 = win.topk(2, dim=-1).values
    scale = conv.abs().max()
    clear = ((top2[..., 0] - top2[..., 1]) > 1e-5 * scale) & (top2[..., 0].abs() > 1e-5 * scale)
    want_code = torch.where(top2[..., 0] > 0, best, torch.full_like(best, 3))
    assert bool((code[clear] == want_code[clear]).all()), "a max-pool selection differs where the window has a clear winner"
    assert float(clear.double().mean()) > 0.99


@pytest.mark.parametrize("B,G,L", [(2, 1, 512), (3, 3, 1000), (2, 12, 5000)])
def test_stem_tc_weight_gradient(B, G, L):
    N, ops, lib, dev = _setup()
    x, w, y, codes, cptr, code = _run_fwd(N, ops, lib, dev, B, G, L, 23 + L)
    L4 = L // 4
    gen = torch.Generator().manual_seed(5)
    S = 1024.0
    dy = (torch.randn(B, 128 * G, L4, generator=gen) * 1e-3).to(dev)
    dy16 = ops.H8(128 * G, B, L4, dev).from_ncl(dy, scale=S)
    dyq = dy16.to_ncl().double() / S                                                  # what the kernel reads
    inv = torch.tensor([1.0 / S], device=dev)
    dw = torch.zeros(128 * G, 1, 15, device=dev)
    N.check(lib.nef_stem_tc_bwd(N.ptr(x), C.c_void_p(cptr), C.c_void_p(dy16.ptr), N.ptr(dw), N.ptr(inv), B, G, L, N.stream_ptr()),
            "nef_stem_tc_bwd")
    # reference: the pooled output on the device's selections is linear in w
    wd = w.double().requires_grad_(True)
    conv = F.conv1d(x.double(), wd, stride=2, padding=7, groups=G)
    idx = (2 * torch.arange(L4, device=dev) - 1)[None, None, :] + code.clamp(max=2)
    out = torch.gather(conv, 2, idx.clamp(min=0)) * (code < 3).double()
    out.backward(dyq)
    err = float((dw.double() - wd.grad).norm() / wd.grad.norm())
    print("stem_tc_bwd B%d G%d L%d: rel-L2 %.2e" % (B, G, L, err))
    # conv position 2j+1 collects dy[j] (code 2) and dy[j+1] (code 0); where both are set the kernel adds the two fp16 values
    # in fp16 (one more 11-bit rounding, as every back-propagated gradient of the fp16 backward carries): ~1.5e-4 overall
    assert err < 5e-4
    # and the CUDA-core kernel on the same codes and the same (fp32) gradient agrees
    dy32 = ops.Cbl4(128 * G, B, L4, dev).from_ncl(dyq.float())
    dw2 = torch.zeros_like(dw)
    N.check(lib.nef_stem_bwd(N.ptr(x), C.c_void_p(cptr), C.c_void_p(dy32.ptr), N.ptr(dw2), B, G, L, N.stream_ptr()), "nef_stem_bwd")
    assert float((dw2.double() - wd.grad).norm() / wd.grad.norm()) < 1e-5
