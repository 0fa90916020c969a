"""CPU: the B200 arithmetic model of the oracle (oracle/b200_precision.py) -- rounding primitives against known values, and the
effect it exists to expose: TF32-level operand rounding leaves the outputs within the north-star 1e-3 but moves the
gradients of this ReLU network by percents (mask flips), which is why GPU gradient parity is judged against the model."""
import random

import torch

from oracle import nefnet_oracle as O
from oracle.b200_precision import B200Precision, f16_sat, tf32_rna


def test_tf32_rna_known_values():
    x = torch.tensor([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, 1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -11), 3.0e-39, 0.0])
    y = tf32_rna(x)
    assert y[0] == 1.0
    assert y[1] == 1.0 + 2.0 ** -10          # tie rounds away from zero
    assert y[2] == 1.0
    assert y[3] == 1.0 + 2.0 ** -10
    assert y[4] == -(1.0 + 2.0 ** -10)
    assert y[6] == 0.0
    r = torch.randn(100000)
    assert float(((tf32_rna(r) - r).abs() / r.abs()).max()) <= 2.0 ** -11 + 1e-9
    assert torch.equal(tf32_rna(tf32_rna(r)), tf32_rna(r))
    assert float(f16_sat(torch.tensor([1e6]))[0]) == 65504.0
    t = tf32_rna(r)
    big = t.abs() > 2.0 ** -14
    assert torch.equal(f16_sat(t)[big], t[big])   # TF32-rounded values in fp16's normal range are exact in fp16


def test_model_keeps_outputs_and_moves_gradients():
    B, G, L, seed = 2, 2, 256, 3
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed)
    random.seed(seed)
    gen = torch.Generator().manual_seed(seed)
    ups = [torch.randn(B, 1, L, generator=gen) for _ in range(3)]

    def run(prec):
        Po = {k: v.clone() for k, v in P.items()}
        for n in O.live_param_names(G):
            Po[n].requires_grad_(True)
        oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(0, 1),
                       prec=prec)
        torch.autograd.backward(oo, ups)
        return [o.detach() for o in oo], {n: Po[n].grad for n in O.live_param_names(G)}

    o0, g0 = run(None)
    o1, g1 = run(B200Precision())
    assert max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(o1, o0)) < 1e-3
    errs = [float((g1[n] - g0[n]).norm() / g0[n].norm()) for n in g0 if n not in O.ZERO_GRAD_PARAMS]
    assert max(errs) < 0.2 and max(errs) > 1e-3, max(errs)
    o2, g2 = run(B200Precision())
    assert all(torch.equal(a, b) for a, b in zip(o1, o2))
