"""Standin-Learning loss on the device (reference: network/loss/losses.py:21-50).

    loss = f0 * L1(predict.detach(), predict_shuffle_p) + f1 * L1(predict.detach(), predict_shuffle_l)
         + f2 * reg(predict, target),            reg = L1 or MSE (cfg.SOLVER.reg_loss)

One fused reduction kernel forward and one elementwise kernel backward (nef_loss_fwd / nef_loss_bwd).
"""
import ctypes as C

import torch

from .. import _native as N


def _prep(t, like):
    return t.detach().to(device=like.device, dtype=torch.float32).contiguous()


class _StandinLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predict, *rest):
        if predict.device.type != "cuda":
            raise RuntimeError("losswrapper (B200) needs CUDA tensors; there is no CPU path")
        with N.guard(predict):       # launches go to the tensors' device, whatever device is current
            return _StandinLoss._fwd(ctx, predict, *rest)

    @staticmethod
    def backward(ctx, dlosses):
        with N.guard(ctx.saved_tensors[0]):
            return _StandinLoss._bwd(ctx, dlosses)

    @staticmethod
    def _fwd(ctx, predict, pred_p, pred_l, target, gt1, gt2, use_mse, factors, mask):
        lib = N.init(N.device_index(predict))
        o = _prep(predict, predict)
        p, l, t = _prep(pred_p, predict), _prep(pred_l, predict), _prep(target, predict)
        n = o.numel()
        if not (p.numel() == n and l.numel() == n and t.numel() == n):
            raise ValueError("losswrapper: predictions and target must have the same number of elements")
        f = (C.c_float * 3)(*[float(v) for v in factors])
        sums = torch.empty(3, dtype=torch.float64, device=o.device)
        losses = torch.empty(4, dtype=torch.float32, device=o.device)
        if gt1 is None and gt2 is None:
            N.check(lib.nef_loss_fwd(N.ptr(o), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, mask, N.ptr(sums),
                                     N.ptr(losses), N.stream_ptr()), "nef_loss_fwd")
        else:  # explicit references for the two stand-in terms: one masked launch per term
            g1 = _prep(gt1, predict) if gt1 is not None else o
            g2 = _prep(gt2, predict) if gt2 is not None else o
            parts = torch.zeros(3, 4, dtype=torch.float32, device=o.device)
            for k, (a_, m_) in enumerate(((g1, 1), (g2, 2), (o, 4))):
                if mask & m_:
                    N.check(lib.nef_loss_fwd(N.ptr(a_), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, m_,
                                             N.ptr(sums), N.ptr(parts[k]), N.stream_ptr()), "nef_loss_fwd")
            losses = parts.sum(0)
            ctx.gts = (g1, g2)
        ctx.save_for_backward(o, p, l, t)
        ctx.meta = (use_mse, tuple(float(v) for v in factors), mask, gt1 is None and gt2 is None)
        return losses

    @staticmethod
    def _bwd(ctx, dlosses):
        o, p, l, t = ctx.saved_tensors
        use_mse, factors, mask, fused = ctx.meta
        lib = N.load()
        f = (C.c_float * 3)(*factors)
        n = o.numel()
        dl = dlosses.detach().to(torch.float32).contiguous()
        d_o, d_p, d_l = torch.empty_like(o), torch.empty_like(p), torch.empty_like(l)
        if fused:
            N.check(lib.nef_loss_bwd(N.ptr(o), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, mask, N.ptr(dl),
                                     N.ptr(d_o), N.ptr(d_p), N.ptr(d_l), N.stream_ptr()), "nef_loss_bwd")
        else:
            g1, g2 = ctx.gts
            N.check(lib.nef_loss_bwd(N.ptr(g1), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, mask & 1, N.ptr(dl),
                                     None, N.ptr(d_p), None, N.stream_ptr()), "nef_loss_bwd")
            N.check(lib.nef_loss_bwd(N.ptr(g2), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, mask & 2, N.ptr(dl),
                                     None, None, N.ptr(d_l), N.stream_ptr()), "nef_loss_bwd")
            N.check(lib.nef_loss_bwd(N.ptr(o), N.ptr(p), N.ptr(l), N.ptr(t), n, int(use_mse), f, mask & 4, N.ptr(dl),
                                     N.ptr(d_o), None, None, N.stream_ptr()), "nef_loss_bwd")
        return d_o.view_as(o), d_p.view_as(p), d_l.view_as(l), None, None, None, None, None, None


def pair_loss(a, b, use_mse):
    """mean |a - b| (or squared): the unsupervised validation term (losses.py:47-49); no gradient."""
    with N.guard(a):
        lib = N.init(N.device_index(a))
        a_, b_ = _prep(a, a), _prep(b, a)
        sums = torch.empty(3, dtype=torch.float64, device=a.device)
        res = torch.empty(1, dtype=torch.float32, device=a.device)
        N.check(lib.nef_pair_loss(N.ptr(a_), N.ptr(b_), a_.numel(), int(use_mse), N.ptr(sums), N.ptr(res), N.stream_ptr()),
                "nef_pair_loss")
    return res[0]


def losswrapper(predict, predict_shuffle_p, predict_shuffle_l, target, cfg, rest_out=None, rest_view=None,
                loss1_gt=None, loss2_gt=None):
    """Same signature and return tuple as the reference (losses.py:21-50)."""
    if cfg.SOLVER.reg_loss == 'l2_loss':
        use_mse = True
    elif cfg.SOLVER.reg_loss == 'l1_loss':
        use_mse = False
    else:
        raise NotImplementedError
    using = cfg.SOLVER.loss_using
    mask = (1 if 1 in using else 0) | (2 if 2 in using else 0) | (4 if 3 in using else 0)
    factor = cfg.SOLVER.loss_factor
    losses = _StandinLoss.apply(predict, predict_shuffle_p, predict_shuffle_l, target, loss1_gt, loss2_gt, use_mse,
                                (factor[0], factor[1], factor[2]), mask)
    out = (losses[0], losses[1], losses[2], losses[3])
    if rest_out is not None and rest_view is not None:  # val
        with torch.no_grad():
            out = out + (pair_loss(rest_out, rest_view, use_mse),)
    return out
