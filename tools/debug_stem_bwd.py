"""Which operand precision does nef_stem_tc_bwd realise?  Compares dW with float64 references built from x, from fp16(x) alone
and from fp16(x) + fp16(x - fp16(x)), and from dy exact vs dy as read."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from network import _native as N, ops
import test_gpu_stem_tc as T
dev = torch.device("cuda:0")
lib = N.init(0)
for (B, G, L) in ((2, 1, 512), (2, 12, 5000)):
    x, w, y, codes, cptr, code = T._run_fwd(N, ops, lib, dev, B, G, L, 23 + L)
    L4 = L // 4
    gen = torch.Generator().manual_seed(5)
    S = 1024.0
    dy = (torch.randn(B, 128 * G, L4, generator=gen) * 1e-3).to(dev)
    dy16 = ops.H8(128 * G, B, L4, dev).from_ncl(dy, scale=S)
    dyq = dy16.to_ncl().double() / S
    inv = torch.tensor([1.0 / S], device=dev)
    dw = torch.zeros(128 * G, 1, 15, device=dev)
    N.check(lib.nef_stem_tc_bwd(N.ptr(x), C.c_void_p(cptr), C.c_void_p(dy16.ptr), N.ptr(dw), N.ptr(inv), B, G, L, N.stream_ptr()), "bwd")
    xh = x.half().float()
    xl = (x - xh).half().float()
    idx = (2 * torch.arange(L4, device=dev) - 1)[None, None, :] + code.clamp(max=2)
    for label, xx in (("x", x.double()), ("x_hi", xh.double()), ("x_hi+x_lo", xh.double() + xl.double()), ("x_lo only", xl.double())):
        wd = w.double().requires_grad_(True)
        conv = F.conv1d(xx, wd, stride=2, padding=7, groups=G)
        out = torch.gather(conv, 2, idx.clamp(min=0)) * (code < 3).double()
        out.backward(dyq)
        print("B%d G%d L%d vs %-10s: rel-L2 %.3e   (|ref| %.3e, |dev| %.3e)" % (B, G, L, label, float((dw.double() - wd.grad).norm() / wd.grad.norm()),
                                                                             float(wd.grad.norm()), float(dw.double().norm())))
