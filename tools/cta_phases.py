"""Phase timing of conv_tc CTAs (ns): setup, pipeline fill, main loop, epilogue, teardown."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import numpy as np, torch
from network import ops, _native as N
dev = torch.device("cuda:0"); lib = N.init(0)
G, B, L4 = 12, 256, 1250
C1 = 128 * G
x, y = ops.Cbl4(C1, B, L4, dev), ops.Cbl4(C1, B, L4, dev)
x.data.normal_()
w = torch.randn(C1, 128, 7, device=dev) * 0.03
wpk = ops.pack_conv_weight(w, G)
d = ops.conv_desc(x, wpk, y, G, 128, 128, 7, relu=True, round_tf32=True)
for _ in range(3): ops.gconv_fwd(d)
torch.cuda.synchronize()
n = 1024
buf = (C.c_ulonglong * (8 * n))()
lib.nef_tc_debug_dump(buf, n)
t = np.array(buf, dtype=np.int64).reshape(n, 8)
t0 = t[:, 0].min()
names = ["setup(0-1)", "fill(1-2)", "mainloop(2-3 issue end)", "drain(3-4)", "epilogue(4-5)", "teardown(5-6)", "total(0-6)"]
seg = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5], t[:, 6] - t[:, 0]], 1)
for lo, hi in ((0, 148), (148, 296), (296, 1024)):
    print("CTAs %d..%d (start offsets min %.1f us max %.1f us)" % (lo, hi, (t[lo:hi, 0].min() - t0) / 1e3, (t[lo:hi, 0].max() - t0) / 1e3))
    for i, nm in enumerate(names):
        print("   %-26s median %8.0f ns   p10 %8.0f   p90 %8.0f" % (nm, np.median(seg[lo:hi, i]), np.percentile(seg[lo:hi, i], 10), np.percentile(seg[lo:hi, i], 90)))
