"""Arithmetic model of the B200 path for the CPU oracle.  TEST INFRASTRUCTURE ONLY (see nefnet_oracle.py).

The reference computes in fp32 (``Solver`` calls ``.float()``, solver.py:21).  The B200 path multiplies on tensor cores:
TF32 operands (10-bit stored significand, round-to-nearest / ties away = ``cvt.rna.tf32.f32``) with fp32 accumulation, fp16
operand copies of TF32-rounded values for the six encoder k7 convolutions, and it stores most activations and
back-propagated gradients already rounded to TF32 (DESIGN.md "Precision").  A ReLU whose pre-activation lies within that
rounding distance of zero can flip, and one flipped mask element moves a gradient by a whole term -- so against the plain
fp32 oracle the production gradients differ by several percent in L2 whatever the implementation.

``B200Precision`` restates WHERE the device rounds, as hooks of ``nefnet_oracle.forward(..., prec=...)``:

  conv(kind, x, w, b)   operands as the tensor core sees them: 'tf32' -> weights rounded to TF32 (activations arrive rounded
                        from ``store``); 'fp16' -> both operands additionally through fp16 (saturating, RN-even: the encoder
                        forward; 'dec' = the same for the decoder's convolutions 2-4 when dec_f16); 'fp32' -> untouched (stem and output kernels on CUDA cores; the decoder's first
                        convolution is evaluated split-precision, x_hi w_hi + x_lo w_hi + x_hi w_lo, i.e. to ~2^-22)
  pattern(name)         optional observed (value != 0) pattern of a ReLU output, see B200Precision
  store(x)              a stored activation: rounded to TF32
  pre(x)                identity forward; the GRADIENT passing through is rounded to TF32 (the device stores the gradient of
                        every pre-activation rounded: masked data-gradient epilogues, bnbwd_apply)

All roundings are straight-through for autograd (the device's backward uses the rounded operands and ignores the rounding
itself).  Accumulation stays in the CPU's fp32, so what is left between this model and the device is accumulation order
(~1e-6 relative) and the operand TRUNCATION the tensor core applies to the few gradient tensors that are stored unrounded.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def tf32_rna(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round the magnitude to 10 explicit significand bits, ties away from zero."""
    bits = x.contiguous().view(torch.int32)
    mag = bits & 0x7FFFFFFF
    sign = bits & -0x80000000
    mag = (mag + 0x1000) & 0x7FFFE000
    return (mag | sign).view(torch.float32)


def f16_sat(x: torch.Tensor) -> torch.Tensor:
    """cvt.rn.satfinite.f16.f32 and back"""
    return x.clamp(-65504.0, 65504.0).half().float()


class _STE(torch.autograd.Function):
    """y = fn(x) forward, dy/dx = 1 backward"""

    @staticmethod
    def forward(ctx, x, which):
        return tf32_rna(x) if which == 0 else f16_sat(tf32_rna(x))

    @staticmethod
    def backward(ctx, g):
        return g, None


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return tf32_rna(g)


class B200Precision:
    """patterns: optional name -> 0/1 tensor, the (value != 0) pattern of an activation as observed on the device
    ('<block>.h', '<block>.y', 'dec<k>.<stage>.<bn>'); the oracle then multiplies by it instead of applying ReLU (and the
    dropout keep-mask, which the pattern of a block's h contains).  acts: filled with the named intermediates of a run."""

    def __init__(self, fwd_f16: bool = True, patterns=None, record: bool = False, dec_f16: bool = True, z2_f16: bool = True):
        self.fwd_f16 = fwd_f16
        self.dec_f16 = dec_f16 and fwd_f16     # decoder convolutions 2-4 on fp16 operand copies (nef_set_dec_f16)
        self.z2_f16 = z2_f16 and fwd_f16       # the z2_conv2 chain on fp16 operand copies (NEF_Z2_F16)
        self.patterns = patterns or {}
        self.acts = {} if record else None

    def pattern(self, name):
        return self.patterns.get(name)

    def observe(self, name, x):
        if self.acts is not None:
            self.acts[name] = x.detach()

    def _ops(self, kind, x, w):
        if kind == "fp32":
            return x, w
        if (kind == "fp16" and self.fwd_f16) or (kind == "dec" and self.dec_f16) or (kind == "z2" and self.z2_f16):
            return _STE.apply(x, 1), _f16w(w)
        return x, _STE.apply(w, 0)

    def conv(self, kind, x, w, b=None, **kw):
        x, w = self._ops(kind, x, w)
        return F.conv1d(x, w, b, **kw)

    def conv_transpose(self, x, w, b=None, **kw):
        if self.z2_f16:
            return F.conv_transpose1d(_STE.apply(x, 1), _f16w(w), b, **kw)
        return F.conv_transpose1d(x, _STE.apply(w, 0), b, **kw)

    def store(self, x):
        return _STE.apply(x, 0)

    def pre(self, x):
        return _RoundGrad.apply(x)


class _F16W(torch.autograd.Function):
    """the packer converts the fp32 weight straight to fp16 (cvt.rn.satfinite.f16x2.f32)"""

    @staticmethod
    def forward(ctx, w):
        return f16_sat(w)

    @staticmethod
    def backward(ctx, g):
        return g


def _f16w(w):
    return _F16W.apply(w)
