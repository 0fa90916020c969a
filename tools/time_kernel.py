"""CUDA-event timing of the k7 grouped conv forward / wgrad at the bench shape (B x 1536 x 1250)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
from network import ops, _native as N
dev = torch.device("cuda:0"); lib = N.init(0)
G, B, L4 = 12, int(os.environ.get("B", 256)), 1250
C1 = 128 * G
x, y = ops.Cbl4(C1, B, L4, dev), ops.Cbl4(C1, B, L4, dev)
x.data.normal_(); y.data.normal_()
w = torch.randn(C1, 128, 7, device=dev) * 0.03
wpk = ops.pack_conv_weight(w, G)
d = ops.conv_desc(x, wpk, y, G, 128, 128, 7, relu=True, round_tf32=True)
dw = torch.zeros_like(w)
def t(fn, n=8):
    for _ in range(2): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
print("conv k7 fwd  %.3f ms" % t(lambda: ops.gconv_fwd(d)))
print("wgrad k7     %.3f ms" % t(lambda: ops.gconv_wgrad(y, x, dw, G, 128, 128, 7)))
