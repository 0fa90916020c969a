"""GPU: the 1e-3 output bar away from the default initialisation (VERDICT r1 "what's weak" 3).

No trained checkpoint can be obtained offline, so "trained-scale" weights are emulated two ways: (a) every weight tensor
rescaled by its own factor in [0.6, 1.7], BatchNorm gamma in [0.5, 1.5] and beta in [-0.3, 0.3], biases x 3; (b) the weights
after eight SGD steps of the CPU oracle at a learning rate (3.0) that moves them by percents.  Outputs of the production arithmetic (fp16 / TF32
tensor cores) must stay within 1e-3 relative of the fp32 oracle.  A third case drives the fp16 operand copies into
saturation (stem weights x 1e7): conversions saturate to the largest finite value instead of overflowing, so everything
stays finite (outputs and every gradient); the difference to the arithmetic-model oracle is reported."""
import random

import numpy as np
import pytest
import torch

from oracle import nefnet_oracle as O
from oracle.b200_precision import B200Precision
from test_gpu_parity import _model, _to

pytestmark = pytest.mark.gpu

OUT_RTOL = 1e-3


def _rescaled(P, seed):
    gen = torch.Generator().manual_seed(seed)
    Q = {}
    for k, v in P.items():
        if "running_" in k or "num_batches" in k:
            Q[k] = v.clone()
        elif k.endswith("double_conv.1.weight") or k.endswith("double_conv.4.weight"):      # BatchNorm gamma
            Q[k] = 0.5 + torch.rand(v.shape, generator=gen)
        elif k.endswith("double_conv.1.bias") or k.endswith("double_conv.4.bias"):          # BatchNorm beta
            Q[k] = 0.6 * torch.rand(v.shape, generator=gen) - 0.3
        elif k.endswith(".bias"):
            Q[k] = v * 3.0
        else:
            Q[k] = v * float(0.6 + 1.1 * torch.rand((), generator=gen))
    return Q


def _after_sgd(P, G, L, steps, seed):
    Q = {k: v.clone() for k, v in P.items()}
    mom = {}
    for it in range(steps):
        inp = O.make_inputs(4, G, L, seed + it)
        # lr far above the reference's 0.1: eight steps must move the weights by percents, as a long training run would
        O.train_step(Q, inp, lead_choice=(it % G, (it + 1) % G), lr=3.0, momentum=0.9, momentum_buf=mom)
    return {k: v.detach().clone() for k, v in Q.items()}


def _worst_rel(dev_outs, ref_outs):
    return max(float(((a.detach().cpu() - b.detach()).abs() / b.detach().abs()).max()) for a, b in zip(dev_outs, ref_outs))


@pytest.mark.parametrize("case", ["rescaled_a", "rescaled_b", "after_sgd"])
def test_outputs_within_1e3_at_trained_scale(case):
    dev = torch.device("cuda:0")
    G, B, L, seed = 12, 4, 1000, 41
    P = O.make_params(G, seed)
    if case == "after_sgd":
        P0 = P
        P = _after_sgd(P, G, 500, 8, seed)
        moved = sorted(float((P[k] - P0[k]).norm() / (P0[k].norm() + 1e-12)) for k in O.live_param_names(G))
        print("after_sgd: median relative weight change %.3f, max %.2f" % (moved[len(moved) // 2], moved[-1]))
    else:
        P = _rescaled(P, seed + (1 if case == "rescaled_b" else 0))
    inp = O.make_inputs(B, G, L, seed + 7)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    m = _model(G, P, dev)
    random.seed(seed)
    d = _to(inp, dev)
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    with torch.no_grad():
        ref = O.forward({k: v.clone() for k, v in P.items()}, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                        phase="train", lead_choice=(c1, c2))
    worst = _worst_rel(outs, ref)
    spread = float(torch.stack([r.max() - r.min() for r in ref]).max())
    print("trained-scale %s: worst output error %.3e relative (bar %.0e); output range %.3f" % (case, worst, OUT_RTOL, spread))
    assert worst < OUT_RTOL


def test_fp16_saturation_stays_finite():
    dev = torch.device("cuda:0")
    G, B, L, seed = 3, 2, 512, 43
    P = O.make_params(G, seed)
    P["W_encoder.conv1.weight"] = P["W_encoder.conv1.weight"] * 1.0e7      # stem outputs (~0.03 at the default scale) far beyond 65504
    inp = O.make_inputs(B, G, L, seed)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    m = _model(G, P, dev)
    random.seed(seed)
    d = _to(inp, dev)
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    torch.autograd.backward(outs, [torch.ones_like(o) / o.numel() for o in outs])
    assert all(bool(torch.isfinite(o).all()) for o in outs)
    named = dict(m.named_parameters())
    assert all(bool(torch.isfinite(named[n].grad).all()) for n in O.live_param_names(G))
    stem = m.export_activation("stem").cpu()
    assert float(stem.max()) == 65504.0, "the stem's fp16 copy must have saturated in this test"
    with torch.no_grad():
        ref = O.forward({k: v.clone() for k, v in P.items()}, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                        phase="train", lead_choice=(c1, c2), prec=B200Precision())
    worst = _worst_rel(outs, ref)
    # reported, not asserted: the device also reads identity residuals from the saturated fp16 copies, where the model adds
    # the unsaturated fp32 tensor -- beyond fp16's range the two are different (finite) computations
    print("saturated fp16 operands: worst output difference vs the arithmetic-model oracle %.3e" % worst)
