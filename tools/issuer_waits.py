"""Where does the MMA-issuing warp of the persistent conv kernel wait? (cycles, per CTA)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import numpy as np, torch
from network import ops, _native as N
dev = torch.device("cuda:0"); lib = N.init(0)
taps = int(os.environ.get("TAPS", 7))
G, B, L4 = 12, 256, 1250
C1 = 128 * G
x, y = ops.Cbl4(C1, B, L4, dev), ops.Cbl4(C1, B, L4, dev)
x.data.normal_()
w = torch.randn(C1, 128, taps, device=dev) * 0.03
wpk = ops.pack_conv_weight(w, G)
d = ops.conv_desc(x, wpk, y, G, 128, 128, taps, relu=True, round_tf32=True)
for _ in range(3): ops.gconv_fwd(d)
torch.cuda.synchronize()
n = 148
buf = (C.c_ulonglong * (8 * n))()
lib.nef_tc_debug_dump(buf, n)  # wait counters need NEF_TC_WS=8 (timing mode)
t = np.array(buf, dtype=np.int64).reshape(n, 8)
tot = t[:, 0].astype(float)
print("taps %d  issuer cycles median %.0f ; waiting: acc_empty %.1f%%  full_x %.1f%%  full_w %.1f%%" % (
    taps, np.median(tot), 100 * np.median(t[:, 1] / tot), 100 * np.median(t[:, 2] / tot), 100 * np.median(t[:, 3] / tot)))
