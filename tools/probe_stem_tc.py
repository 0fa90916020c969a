"""Accuracy and speed of the tensor-core stem (nef_stem_tc_fwd) against an fp64 reference, at the natural operand scale and
with the operands pre-scaled by powers of two (does kind::f16 flush subnormal fp16 inputs?).
    python tools/probe_stem_tc.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
import torch.nn.functional as F
from network import _native as N, ops
dev = torch.device("cuda:0")
lib = N.init(0)


def run(B, G, L, xs, ws, seed=0):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.rand(B, G, L, generator=gen).to(dev) * xs
    w = (torch.randn(128 * G, 1, 15, generator=gen) * 0.032).to(dev) * ws
    ref = F.max_pool1d(F.relu(F.conv1d(x.double(), w.double(), stride=2, padding=7, groups=G)), 3, 2, 1)
    y = ops.H8(128 * G, B, L // 4, dev)
    am = torch.zeros(32 * G * y.rows + 2 * N.GUARD_ROWS, dtype=torch.int32, device=dev)
    N.check(lib.nef_stem_tc_fwd(N.ptr(x), N.ptr(w), C.c_void_p(y.ptr), C.c_void_p(am.data_ptr()), B, G, L, N.stream_ptr()), "nef_stem_tc_fwd")
    got = y.to_ncl().double()
    want = ref.float().half().double()   # the device stores fp16
    err = float((got - want).norm() / want.norm())
    err32 = float((got - ref).norm() / ref.norm())
    print("B%d G%d L%d x*%g w*%g: rel-L2 vs fp16(ref) %.3e, vs fp64 ref %.3e, max abs %.3e (ref max %.3e)"
          % (B, G, L, xs, ws, err, err32, float((got - want).abs().max()), float(ref.abs().max())), flush=True)


for xs, ws in ((1.0, 1.0), (64.0, 1.0), (1.0, 256.0), (64.0, 256.0)):
    run(4, 3, 512, xs, ws)
run(8, 12, 5000, 1.0, 1.0)
run(8, 12, 5000, 64.0, 256.0)
# speed at the bench shape
B, G, L = 256, 12, 5000
x = torch.rand(B, G, L, device=dev)
w = torch.randn(128 * G, 1, 15, device=dev) * 0.032
y = ops.H8(128 * G, B, L // 4, dev)
am = torch.zeros(32 * G * y.rows + 2 * N.GUARD_ROWS, dtype=torch.int32, device=dev)
for _ in range(3):
    lib.nef_stem_tc_fwd(N.ptr(x), N.ptr(w), C.c_void_p(y.ptr), C.c_void_p(am.data_ptr()), B, G, L, N.stream_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.nef_stem_tc_fwd(N.ptr(x), N.ptr(w), C.c_void_p(y.ptr), C.c_void_p(am.data_ptr()), B, G, L, N.stream_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = (y.buf.numel() * 2 + am.numel() * 4 + x.numel() * 4) / 1e9
print("stem_tc_fwd 256x12x5000: %.3f ms, %.2f GB -> %.0f GB/s" % (ms, gb, gb / ms * 1e3))

import numpy as np
lib.nef_stem_tc_debug.restype = C.c_int
lib.nef_stem_tc_debug.argtypes = [C.c_void_p]
buf = np.zeros(64, dtype=np.int64)
lib.nef_stem_tc_debug(buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(8, 8)
print("phase cycles of CTA 0 (window, build, issue, mma wait, epilogue+sync | tile total):")
for i in range(8):
    r = t[i]
    nxt = t[i + 1][0] if i < 7 else r[5]
    print("  tile %d: %6d %6d %6d %6d %6d | %6d" % (i, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - r[0]))
