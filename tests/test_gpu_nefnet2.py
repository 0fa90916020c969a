"""GPU parity of ``Model_nefnet2`` (reference network/model_nefnet2.py:63-203) against its CPU oracle
(oracle/nefnet2_oracle.py, pinned to the unmodified reference class by tests/golden/nefnet2_*.npz) and against those golden
vectors themselves.  Same tiers as tests/test_gpu_parity.py: the fp32 CUDA-core tier pins the dataflow (shared weights read
by every lead group, weight gradients accumulated over the leads, single_conv_z1 / z2 applied after the lead mean), the
tensor-core tier is the production arithmetic."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import nefnet_oracle as O
from oracle import nefnet2_oracle as O2
from test_gpu_parity import MODES, mode, _to, EXACT_OUT_RTOL, OUT_RTOL, EXACT_GRAD_REL_L2, TF32_GRAD_REL_L2, TF32_GRAD_COS
from test_oracle_golden import sample_idx

pytestmark = pytest.mark.gpu

NEFNET2 = ("nefnet2_train_b2_g3_l256", "nefnet2_test_b2_g2_l128_v3")


def _model(G, P, dev, train=True):
    from network.model_nefnet2 import Model_nefnet2
    m = Model_nefnet2(theta_encoder_len=1, lead_num=G)
    sd = m.state_dict()
    assert list(sd.keys()) == list(P.keys())
    m.load_state_dict({k: v.clone() for k, v in P.items()})
    m = m.float().to(dev)
    m.train(train)
    m.dropout_p = 0.0
    return m


def _grad_check(named, ref, exact, what):
    worst = (0.0, "")
    for n, r in ref.items():
        if n in O.ZERO_GRAD_PARAMS:
            assert bool(torch.isfinite(named[n].grad).all()), n
            continue
        got, r = named[n].grad.detach().cpu().double(), r.double()
        err = float((got - r).norm() / (r.norm() + 1e-30))
        cos = float((got * r).sum() / (got.norm() * r.norm() + 1e-30))
        worst = max(worst, (err, n))
        if exact:
            assert err < EXACT_GRAD_REL_L2, (what, n, err)
        else:
            assert err < TF32_GRAD_REL_L2 and cos > TF32_GRAD_COS, (what, n, err, cos)
    print(what, "worst grad rel-L2 %.3e (%s)" % worst)


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("B,G,L,seed,ragged", [(3, 3, 512, 21, False), (2, 12, 1000, 22, False), (4, 2, 264, 23, True)])
def test_oracle_fixed_upstream(B, G, L, seed, ragged, mode_name):
    dev = torch.device("cuda:0")
    exact = MODES[mode_name][1] == 1
    P = O2.make_params(seed)
    inp = O.make_inputs(B, G, L, seed, ragged_rois=ragged)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    gen = torch.Generator().manual_seed(seed)
    ups = [torch.randn(B, 1, L, generator=gen) for _ in range(3)]
    with mode(mode_name):
        m = _model(G, P, dev)
        random.seed(seed)
        d = _to(inp, dev)
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        torch.autograd.backward(outs, [u.to(dev) for u in ups])
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        named = dict(m.named_parameters())
    Po = {k: v.clone() for k, v in P.items()}
    for n in O2.live_param_names():
        Po[n].requires_grad_(True)
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    oo = O2.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(c1, c2),
                    stats_out=stats)
    torch.autograd.backward(oo, ups)
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().numpy(), rtol=EXACT_OUT_RTOL if exact else OUT_RTOL, atol=0)
    _grad_check(named, {n: Po[n].grad for n in O2.live_param_names()}, exact, "nefnet2 %s B%d G%d L%d" % (mode_name, B, G, L))
    for n in O2.UNUSED_PARAMS:
        assert named[n].grad is None
    for k, v in stats.items():
        np.testing.assert_allclose(sd[k].numpy(), v.numpy(), rtol=1e-4 if exact else 2e-3, atol=1e-5 if exact else 2e-4)


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("name", NEFNET2)
def test_golden(name, mode_name, golden_dir, cfg):
    """Vectors produced by the unmodified reference class (oracle/make_golden_nefnet2.py)."""
    from network import build_loss
    dev = torch.device("cuda:0")
    exact = MODES[mode_name][1] == 1
    out_rtol = EXACT_OUT_RTOL if exact else OUT_RTOL
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, G, L, seed, V = (int(g[k]) for k in ("B", "G", "L", "seed", "V"))
    lead = tuple(int(v) for v in g["lead_choice"])
    P = O2.make_params(seed)
    inp = _to(O.make_inputs(B, G, L, seed, V=V, ragged_rois=bool(int(g["ragged"]))), dev)
    loss_fn = build_loss(cfg)
    train = "train" in name
    with mode(mode_name):
        m = _model(G, P, dev, train=train)
        # the golden run drew its two leads from random.seed(seed) (oracle/make_golden_nefnet2.py); find a seed state that
        # reproduces them is unnecessary: patch the draws
        draws = iter(lead)
        orig = random.randint
        random.randint = lambda a, b: next(draws)
        try:
            if train:
                outs = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
            else:
                outs = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"], phase="test")
        finally:
            random.randint = orig
        if train:
            losses = loss_fn(outs[0], outs[1], outs[2], inp["target"], cfg)
            losses[0].backward()
            named = dict(m.named_parameters())
            # Gradients through the L1 loss are discontinuous in the outputs (sign(out - target)): only the exact tier is held
            # to the golden gradient samples; the TF32 tiers' gradient bars are test_oracle_fixed_upstream's (smooth upstream).
            for n in O2.live_param_names() if exact else ():
                if n in O.ZERO_GRAD_PARAMS:
                    continue
                gr = named[n].grad.detach().cpu()
                norm_ref = float(g["gn/" + n][0])
                got = gr.flatten()[sample_idx(gr.numel())].double().numpy()
                ref = g["gs/" + n].astype(np.float64)
                assert np.linalg.norm(got - ref) <= 5e-3 * (np.linalg.norm(ref) + 1e-3 * norm_ref), (n, mode_name)
                assert abs(float(gr.double().norm()) - norm_ref) <= 2e-3 * norm_ref, (n, mode_name)
            assert all(bool(torch.isfinite(p.grad).all()) for n, p in named.items() if n in O2.live_param_names())
        else:
            losses = loss_fn(outs[0], outs[1], outs[2], inp["target"], cfg, outs[3], inp["rest_view"])
        sd = m.state_dict()
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().cpu().numpy(), g[f"out{i}"], rtol=out_rtol, atol=0)
    np.testing.assert_allclose(np.array([float(v.detach()) for v in losses]), g["losses"], rtol=1e-5 if exact else 2e-3, atol=1e-5)
    for k in g.files:
        if k.startswith("bn/"):
            np.testing.assert_allclose(sd[k[3:]].cpu().numpy(), g[k], rtol=1e-4 if exact else 2e-3, atol=1e-5 if exact else 2e-4)


def test_unsupported_entry_points_say_so():
    from network.model_nefnet2 import Model_nefnet2
    m = Model_nefnet2(1, 2)
    with pytest.raises(NotImplementedError):
        m.gen_ecg(None, None, None, None)
    with pytest.raises(NotImplementedError):
        m(None, None, None, None, phase="gen")
