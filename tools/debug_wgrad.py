"""One-hot probe of the tcgen05 wgrad kernel: where does dy[r0, m0] * x[r1, n0] land?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
from network import ops, _native as N
dev = torch.device("cuda:0")
lib = N.init(0)
B, L, C, taps = 2, 58, 128, 7   # Lp = 64, rows = 128: two stages, no tail
lib.nef_set_conv_impl(1)
print("dbg", os.environ.get("NEF_WG_DEBUG"))
for (b0, l0, m0, l1, n0) in [(0, 10, 0, 10, 0), (0, 10, 5, 12, 0), (0, 10, 0, 10, 7), (1, 20, 37, 18, 66), (0, 3, 127, 5, 127)]:
    dy = torch.zeros(B, C, L, device=dev); x = torch.zeros(B, C, L, device=dev)
    dy[b0, m0, l0] = 1.0; x[b0, n0, l1] = 1.0
    dyt = ops.Cbl4(C, B, L, dev).from_ncl(dy); xt = ops.Cbl4(C, B, L, dev).from_ncl(x)
    dw = torch.zeros(C, C, taps, device=dev)
    ops.gconv_wgrad(dyt, xt, dw, 1, C, C, taps)
    nz = dw.nonzero().tolist()
    print("dy(b%d,l%d,m%d) x(l%d,n%d) expect [m=%d,n=%d,t=%d] got" % (b0, l0, m0, l1, n0, m0, n0, l1 - l0 + 3),
          [(i, dw[tuple(i)].item()) for i in nz][:8])
dy = torch.ones(B, C, L, device=dev); x = torch.ones(B, C, L, device=dev)
dyt = ops.Cbl4(C, B, L, dev).from_ncl(dy); xt = ops.Cbl4(C, B, L, dev).from_ncl(x)
dw = torch.zeros(C, C, taps, device=dev)
ops.gconv_wgrad(dyt, xt, dw, 1, C, C, taps)
print("ones: dw[0,0,:]", dw[0, 0].tolist(), "min", dw.min().item(), "max", dw.max().item(), "expect", [2 * (58 - abs(t - 3)) for t in range(7)])
