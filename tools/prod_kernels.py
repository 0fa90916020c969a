"""CUDA-event time and MMA-issuer wait breakdown of the PRODUCTION conv launches of the 256 x 12 x 5000 train step, built
through the public op layer exactly as csrc/nef_plan.cu builds them:

    fwd5164  first convolution of a big block: fp16 operands, ReLU + dropout + fp16 copy + bit plane, no fp32 store
    dgrad    masked data gradient on loss-scaled fp16 gradient copies (EPI 14368)
    wgrad16  fp16 weight gradient (nef_gconv_wgrad_f16)

    python tools/prod_kernels.py [taps] [cin_g]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import numpy as np, torch
from network import ops, _native as N

dev = torch.device("cuda:0"); lib = N.init(0)
G, B, L4 = 12, int(os.environ.get("B", 256)), 1250
C1 = 128 * G


def waits(tag, ms, flops, bytes_):
    n = 148
    buf = (C.c_ulonglong * (8 * n))()
    lib.nef_tc_debug_dump(buf, n)
    t = np.array(buf, dtype=np.int64).reshape(n, 8)
    tot = t[:, 0].astype(float)
    print("%-28s %.3f ms  %6.0f TFLOP/s  %5.0f GB/s | issuer cycles %.0f ; waits: acc_empty %.1f%%  full_x %.1f%%  full_w %.1f%%" % (
        tag, ms, flops / ms / 1e9, bytes_ / ms / 1e6, np.median(tot), 100 * np.median(t[:, 1] / tot),
        100 * np.median(t[:, 2] / tot), 100 * np.median(t[:, 3] / tot)), flush=True)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def bits_plane():
    return torch.zeros((C1 // 32) * B * (L4 + 2 * N.HALO) + 2 * N.GUARD_ROWS, dtype=torch.int32, device=dev)


def fwd5164(taps, cin_g, drop=0.2):
    cin = cin_g * G
    x16 = ops.H8(cin, B, L4, dev); x16.data.normal_()
    y = ops.Cbl4(C1, B, L4, dev)
    y16 = ops.H8(C1, B, L4, dev)
    w = torch.randn(C1, cin_g, taps, device=dev) * 0.03
    wpk = ops.pack_conv_weight(w, G, f16=True)
    d = ops.conv_desc(y, wpk, y, G, cin_g, 128, taps, relu=True, round_tf32=True)
    ops.use_f16_operand(d, x16, wpk, cin_g)
    bp = bits_plane()
    d.y16 = y16.ptr
    d.out_bits = bp.data_ptr() + 4 * N.GUARD_ROWS
    d.drop_p, d.drop_seed = drop, 1234
    d.y = None
    ms = timeit(lambda: ops.gconv_fwd(d))
    waits("fwd5164 k%d cin %d drop %.1f" % (taps, cin_g, drop), ms, 2.0 * B * L4 * C1 * cin_g * taps, B * L4 * (cin + C1) * 2.0)


def dgrad(taps):
    g16 = ops.H8(C1, B, L4, dev); g16.data.normal_()
    y = ops.Cbl4(C1, B, L4, dev)
    y16 = ops.H8(C1, B, L4, dev)
    w = torch.randn(C1, 128, taps, device=dev) * 0.03
    wpk = ops.pack_conv_weight(w, G, dgrad=True, f16=True)
    d = ops.conv_desc(y, wpk, y, G, 128, 128, taps, round_tf32=True)
    ops.use_f16_operand(d, g16, wpk, 128)
    bp = bits_plane(); bp.fill_(0x55555555)
    sc = torch.tensor([64.0, 1.0 / 64.0], device=dev)
    d.y16 = y16.ptr
    d.mask_bits = bp.data_ptr() + 4 * N.GUARD_ROWS
    d.mask_mode, d.mask_scale = 1, 1.25
    d.mask, d.mask_cstride, d.mask_c4_off, d.mask_c4_gstride = y.ptr, y.rows, 0, 32
    d.acc_scale = sc.data_ptr() + 4
    d.y16_scale = sc.data_ptr()
    d.y = None
    ms = timeit(lambda: ops.gconv_fwd(d))
    waits("dgrad14368 k%d" % taps, ms, 2.0 * B * L4 * C1 * 128 * taps, B * L4 * (C1 + C1) * 2.0)


def wgrad16(taps, cin_g=128):
    cin = cin_g * G
    dy16 = ops.H8(C1, B, L4, dev); dy16.data.normal_()
    x16 = ops.H8(cin, B, L4, dev); x16.data.normal_()
    dw = torch.zeros(C1, cin_g, taps, device=dev)
    sc = torch.tensor([1.0 / 64.0], device=dev)
    ms = timeit(lambda: ops.gconv_wgrad_f16(dy16, x16, dw, G, 128, cin_g, taps, out_scale=sc))
    print("%-28s %.3f ms  %6.0f TFLOP/s  %5.0f GB/s" % ("wgrad_f16 k%d cin %d" % (taps, cin_g), ms,
          2.0 * B * L4 * C1 * cin_g * taps / ms / 1e9, B * L4 * (cin + C1) * 2.0 / ms / 1e6), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "fwd"):
        fwd5164(7, 128); fwd5164(3, 128); fwd5164(3, 64)
    if what in ("all", "dgrad"):
        dgrad(7); dgrad(3)
    if what in ("all", "wgrad"):
        wgrad16(7); wgrad16(3); wgrad16(3, 64)
