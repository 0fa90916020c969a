"""CPU probe (not a test): how far do fp16-operand data / weight gradients move the parameter gradients?

Emulates, inside the fp32 oracle, the backward the next round plans: every convolution's backward reads
fp16 copies of its operands -- the upstream gradient scaled by one global power-of-two loss scale S (saturating conversion),
the saved input and the weights -- multiplies them exactly and accumulates in fp32 (what tcgen05 kind::f16 does), then
removes S.  For comparison the same with TF32-rounded operands (today's backward).  Forward stays fp32 in all three, and the
upstream gradients are fixed random tensors (no loss discontinuity), as in tests/test_gpu_parity.py::test_oracle_fixed_upstream.
    python tests/probe_fp16_backward_accuracy.py > profiles/r01_fp16_backward_accuracy_oracle.txt"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nefnet_oracle as O  # noqa: E402

MODE = {"kind": "fp32", "S": 1.0}
_conv1d = F.conv1d


def _tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _q(t, scale=1.0):
    if MODE["kind"] == "fp16":
        return (t * scale).clamp(-65504.0, 65504.0).half().float()
    if MODE["kind"] == "tf32":
        return _tf32(t * scale)
    return t * scale


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding, groups):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, groups, b is not None)
        return _conv1d(x, w, b, stride=stride, padding=padding, groups=groups)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, padding, groups, has_b = ctx.cfg
        S = MODE["S"] if MODE["kind"] == "fp16" else 1.0
        gq, xq, wq = _q(g, S), _q(x), _q(w)
        dx = torch.nn.grad.conv1d_input(x.shape, wq, gq, stride=stride, padding=padding, groups=groups) / S
        dw = torch.nn.grad.conv1d_weight(xq, w.shape, gq, stride=stride, padding=padding, groups=groups) / S
        db = g.sum(dim=(0, 2)) if has_b else None
        return dx, dw, db, None, None, None


def conv1d(x, w, b=None, stride=1, padding=0, groups=1):
    if MODE["kind"] == "fp32" or w.shape[1] < 64:      # stem and the 64 -> 1 output conv stay on their fp32 kernels
        return _conv1d(x, w, b, stride=stride, padding=padding, groups=groups)
    return QConv.apply(x, w, b, stride, padding, groups)


def grads(kind, S, B, G, L, seed):
    MODE["kind"], MODE["S"] = kind, S
    F.conv1d = conv1d
    try:
        P = O.make_params(G, seed)
        inp = O.make_inputs(B, G, L, seed)
        names = O.live_param_names(G)
        for n in names:
            P[n].requires_grad_(True)
        gen = torch.Generator().manual_seed(seed)
        ups = [torch.randn(B, 1, L, generator=gen) / (B * L) for _ in range(3)]   # the magnitude of the L1 loss's gradient
        outs = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(0, 1 % G))
        torch.autograd.backward(outs, ups)
        return {n: P[n].grad.detach().double() for n in names if n not in O.ZERO_GRAD_PARAMS}
    finally:
        F.conv1d = _conv1d


def main():
    B, G, L, seed = 4, 12, 1000, 3
    S = 2.0 ** round(math.log2(419.0 * B * L))        # S ~ 2^24 at B * L = 8 * 5000
    ref = grads("fp32", 1.0, B, G, L, seed)
    print("B=%d G=%d L=%d, fixed upstream gradients ~ N(0, 1) / (B L); loss scale S = 2^%d" % (B, G, L, round(math.log2(S))))
    print("%-44s %12s %12s   %12s %12s" % ("parameter", "tf32 relL2", "tf32 1-cos", "fp16 relL2", "fp16 1-cos"))
    got = {k: grads(k, S, B, G, L, seed) for k in ("tf32", "fp16")}
    worst = {"tf32": (0.0, 0.0), "fp16": (0.0, 0.0)}
    for n, r in ref.items():
        row = []
        for k in ("tf32", "fp16"):
            g = got[k][n]
            rel = float((g - r).norm() / (r.norm() + 1e-300))
            cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-300))
            row += [rel, 1.0 - cos]
            worst[k] = (max(worst[k][0], rel), max(worst[k][1], 1.0 - cos))
        print("%-44s %12.3e %12.3e   %12.3e %12.3e" % (n, *row))
    for k in ("tf32", "fp16"):
        print("worst %s: rel-L2 %.3e, 1 - cosine %.3e" % (k, *worst[k]))
    for s_exp in (-6, -3, 3, 6):                      # sensitivity: S off by 2^-6 .. 2^6
        g = grads("fp16", S * 2.0 ** s_exp, B, G, L, seed)
        w = max(float((g[n] - ref[n]).norm() / (ref[n].norm() + 1e-300)) for n in ref)
        print("fp16 with S * 2^%+d: worst rel-L2 %.3e" % (s_exp, w))


if __name__ == "__main__":
    main()
