"""Thin Python handles on the single-op entry points of the C ABI (used by the unit tests and by
bench.py's per-kernel roofline measurement).  Device memory comes from torch; everything else is the
library."""
import ctypes as C

import torch

from . import _native as N


class Cbl4:
    """A CBL4 activation tensor (see include/nefnet_b200.h): C channels, B segments, L samples, with the
    guard rows the tile loads may touch on both sides."""

    def __init__(self, C_, B, L, device):
        self.C, self.B, self.L = C_, B, L
        self.Lp = L + 2 * N.HALO
        self.rows = B * self.Lp
        n = (C_ // 4) * self.rows * 4
        g = N.GUARD_ROWS * 4
        self.buf = torch.zeros(n + 2 * g, dtype=torch.float32, device=device)
        self.data = self.buf[g:g + n]

    @property
    def ptr(self):
        return self.data.data_ptr()

    def from_ncl(self, x, round_tf32=False):
        lib = N.load()
        x = x.contiguous().float()
        N.check(lib.nef_ncl_to_cbl4(N.ptr(x), C.c_void_p(self.ptr), self.B, self.C, self.L, int(round_tf32),
                                    N.stream_ptr()), "nef_ncl_to_cbl4")
        return self

    def to_ncl(self):
        lib = N.load()
        out = torch.empty((self.B, self.C, self.L), dtype=torch.float32, device=self.buf.device)
        N.check(lib.nef_cbl4_to_ncl(C.c_void_p(self.ptr), N.ptr(out), self.B, self.C, self.L, N.stream_ptr()),
                "nef_cbl4_to_ncl")
        return out


class H8:
    """fp16 operand copy of a (B, C, L) tensor: `half8 [C/8][B * Lp]` with guard rows (the layout of NefConvDesc.y16)."""

    def __init__(self, C_, B, L, device):
        self.C, self.B, self.L = C_, B, L
        self.Lp = L + 2 * N.HALO
        self.rows = B * self.Lp
        n = (C_ // 8) * self.rows * 8
        g = N.GUARD_ROWS * 8
        self.buf = torch.zeros(n + 2 * g, dtype=torch.float16, device=device)
        self.data = self.buf[g:g + n]

    @property
    def ptr(self):
        return self.data.data_ptr()

    def from_ncl(self, x, scale=1.0):
        x = x.contiguous().float()
        N.check(N.load().nef_ncl_to_h8(N.ptr(x), C.c_void_p(self.ptr), self.B, self.C, self.L, float(scale), N.stream_ptr()),
                "nef_ncl_to_h8")
        return self

    def to_ncl(self):
        out = torch.empty((self.B, self.C, self.L), dtype=torch.float32, device=self.buf.device)
        N.check(N.load().nef_h8_to_ncl(C.c_void_p(self.ptr), N.ptr(out), self.B, self.C, self.L, N.stream_ptr()), "nef_h8_to_ncl")
        return out


def pack_conv_weight(w, groups, dgrad=False, lo=False, f16=False):
    """Conv1d weight (groups*cout_g, cin_g, k) -> packed forward (or data-gradient) operand; f16: the fp16 operand packing
    of NefConvTerm.x_f16 (half the bytes; returned as a float32-typed buffer of half the length)."""
    lib = N.load()
    w = w.contiguous().float()
    cout, cin_g, k = w.shape
    cout_g = cout // groups
    out = torch.empty(w.numel() // 2 if f16 else w.numel(), dtype=torch.float32, device=w.device)
    flags = (1 if dgrad else 0) | (2 if lo else 0) | (4 if f16 else 0)
    if not dgrad:
        N.check(lib.nef_pack_weights(N.ptr(w), N.ptr(out), groups, cout_g, cin_g, k, cout_g * cin_g * k, cin_g * k, k,
                                     1, flags, N.stream_ptr()), "nef_pack_weights")
    else:
        N.check(lib.nef_pack_weights(N.ptr(w), N.ptr(out), groups, cin_g, cout_g, k, cout_g * cin_g * k, k, cin_g * k,
                                     1, flags, N.stream_ptr()), "nef_pack_weights")
    return out


def conv_desc(x: Cbl4, wpk, y: Cbl4, groups, cin_g, cout_g, taps, x_off=0, x_gs=None, y_off=0, y_gs=None, relu=False,
              round_tf32=False, bias=None, res: Cbl4 = None):
    d = N.NefConvDesc()
    d.n_terms, d.groups, d.N, d.round_tf32 = 1, groups, cout_g, int(round_tf32)
    t = d.term[0]
    t.x, t.x_cstride, t.x_c4_off = x.ptr, x.rows, x_off
    t.x_c4_gstride = cin_g // 4 if x_gs is None else x_gs
    t.cin_g, t.taps, t.tap_off, t.w = cin_g, taps, -(taps // 2), wpk.data_ptr()
    d.rows, d.Lp, d.L = x.rows, x.Lp, x.L
    d.y, d.y_cstride, d.y_c4_off = y.ptr, y.rows, y_off
    d.y_c4_gstride = cout_g // 4 if y_gs is None else y_gs
    d.y_Lp, d.y_lmul, d.y_ladd = y.Lp, 1, 0
    d.relu = int(relu)
    d.mask_scale = 1.0
    if bias is not None:
        d.bias = bias.data_ptr()
    if res is not None:
        d.res, d.res_cstride, d.res_c4_off, d.res_c4_gstride = res.ptr, res.rows, 0, cout_g // 4
    return d


def use_f16_operand(d, x16: H8, wpk16, cin_g, term=0, x_off=0, x_gs=None):
    """Switches term `term` of a descriptor built by conv_desc to the fp16 operand copy x16 and fp16-packed weights
    (offsets / strides in 4-channel chunk units, as everywhere; cin_g = real input channels per group)."""
    t = d.term[term]
    t.x, t.x_cstride = x16.ptr, x16.rows
    t.x_c4_off = x_off // 2
    t.x_c4_gstride = (cin_g // 4 if x_gs is None else x_gs) // 2
    t.cin_g, t.x_f16, t.w = cin_g // 2, 1, wpk16.data_ptr()
    return d


def gconv_wgrad_f16(dy16: H8, x16: H8, dw, groups, cout_g, cin_g, taps, out_scale=None):
    """nef_gconv_wgrad_f16: dw (groups*cout_g, cin_g, taps) += out_scale * dy16^T x16; out_scale: device scalar tensor or None."""
    lib = N.load()
    d = N.NefWgradDesc()
    d.dy_cstride, d.dy_c4_off, d.dy_c4_gstride = dy16.rows, 0, cout_g // 4
    d.x_cstride, d.x_c4_off, d.x_c4_gstride = x16.rows, 0, cin_g // 4
    d.cout_g, d.cin_g, d.groups, d.taps, d.tap_off = cout_g, cin_g, groups, taps, -(taps // 2)
    d.rows = dy16.rows
    d.dw, d.sg, d.sm, d.sn, d.st = dw.data_ptr(), cout_g * cin_g * taps, cin_g * taps, taps, 1
    N.check(lib.nef_gconv_wgrad_f16(C.byref(d), C.c_void_p(dy16.ptr), C.c_void_p(x16.ptr), N.ptr(out_scale), N.stream_ptr()),
            "nef_gconv_wgrad_f16")


def gconv_fwd(d):
    lib = N.load()
    N.check(lib.nef_gconv_fwd(C.byref(d), N.stream_ptr()), "nef_gconv_fwd")


def gconv_wgrad(dy: Cbl4, x: Cbl4, dw, groups, cout_g, cin_g, taps, db=None):
    """dw: Conv1d-layout gradient tensor (groups*cout_g, cin_g, taps), accumulated into."""
    lib = N.load()
    d = N.NefWgradDesc()
    d.dy, d.dy_cstride, d.dy_c4_off, d.dy_c4_gstride = dy.ptr, dy.rows, 0, cout_g // 4
    d.x, d.x_cstride, d.x_c4_off, d.x_c4_gstride = x.ptr, x.rows, 0, cin_g // 4
    d.cout_g, d.cin_g, d.groups, d.taps, d.tap_off = cout_g, cin_g, groups, taps, -(taps // 2)
    d.rows = dy.rows
    d.dw, d.sg, d.sm, d.sn, d.st = dw.data_ptr(), cout_g * cin_g * taps, cin_g * taps, taps, 1
    if db is not None:
        d.db = db.data_ptr()
    N.check(lib.nef_gconv_wgrad(C.byref(d), N.stream_ptr()), "nef_gconv_wgrad")
