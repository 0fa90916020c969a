"""GPU parity of the CUDA path (through the reference-facing module surface and the C ABI) against
(1) the golden vectors the unmodified reference produced and (2) the CPU oracle on fresh seeded inputs.

Tolerances: outputs within 1e-3 relative (north star); gradients are compared in relative L2 per
tensor (the convolutions multiply in TF32, as the reference's own cuDNN path does by default)."""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import nefnet_oracle as O
from oracle.make_golden import sample_idx

pytestmark = pytest.mark.gpu

GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
OUT_RTOL = 1e-3
GRAD_REL_L2 = 1e-2


def _model(G, P, dev, train=True, dropout=0.0):
    import network
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=G)
    m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    m = m.float().to(dev)
    m.train(train)
    m.dropout_p = dropout
    return m


def _to(inp, dev):
    return {k: v.to(dev) for k, v in inp.items()}


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("name", GOLDEN)
def test_golden(name, impl, golden_dir, cfg):
    from network import _native as N
    from network import build_loss
    dev = torch.device("cuda:0")
    N.init(0).nef_set_conv_impl(impl)
    try:
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        B, G, L, seed, V = (int(g[k]) for k in ("B", "G", "L", "seed", "V"))
        P = O.make_params(G, seed)
        inp = _to(O.make_inputs(B, G, L, seed, V=V, ragged_rois=bool(int(g["ragged"]))), dev)
        train = name.startswith("train")
        m = _model(G, P, dev, train=train)
        loss_fn = build_loss(cfg)
        random.seed(seed)
        if train:
            outs = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
            losses = loss_fn(outs[0], outs[1], outs[2], inp["target"], cfg)
            losses[0].backward()
        else:
            outs = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"],
                     phase="test")
            losses = loss_fn(outs[0], outs[1], outs[2], inp["target"], cfg, outs[3], inp["rest_view"])
            random.seed(seed)
            z1, z2 = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
            gen = m.gen_ecg(z1, z2, inp["rest_theta"], inp["rois"])
            np.testing.assert_allclose(gen.cpu().numpy(), g["gen_ecg"], rtol=OUT_RTOL, atol=0)
            zz1 = z1.cpu().flatten()[sample_idx(z1.numel(), 256)].numpy()
            zz2 = z2.cpu().flatten()[sample_idx(z2.numel(), 256)].numpy()
            np.testing.assert_allclose(zz1, g["gen_z1_sample"], rtol=5e-3, atol=5e-3 * float(np.abs(g["gen_z1_sample"]).max()))
            np.testing.assert_allclose(zz2, g["gen_z2_sample"], rtol=5e-3, atol=5e-3 * float(np.abs(g["gen_z2_sample"]).max()))
        for i, o in enumerate(outs):
            np.testing.assert_allclose(o.cpu().numpy(), g[f"out{i}"], rtol=OUT_RTOL, atol=0)
        np.testing.assert_allclose(np.array([float(v) for v in losses]), g["losses"], rtol=2e-3, atol=1e-6)
        sd = m.state_dict()
        for k in g.files:
            if k.startswith("bn/"):
                np.testing.assert_allclose(sd[k[3:]].cpu().numpy(), g[k], rtol=2e-3, atol=2e-4)
        if train:
            named = dict(m.named_parameters())
            for n in O.UNUSED_PARAMS:
                assert named[n].grad is None
            worst = 0.0
            for n in O.live_param_names(G):
                gr = named[n].grad
                assert gr is not None, n
                gr = gr.cpu()
                if n in O.ZERO_GRAD_PARAMS:
                    assert float(gr.abs().max()) < 1e-4, n
                    continue
                norm_ref = float(g["gn/" + n][0])
                got = gr.flatten()[sample_idx(gr.numel())].numpy()
                ref = g["gs/" + n]
                err = float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30))
                nerr = abs(float(gr.double().norm()) - norm_ref) / norm_ref
                worst = max(worst, err, nerr)
                assert err < 3 * GRAD_REL_L2 and nerr < GRAD_REL_L2, (n, err, nerr)
            print(name, "impl", impl, "worst grad err", worst)
    finally:
        N.load().nef_set_conv_impl(1)


@pytest.mark.parametrize("B,G,L,seed", [(4, 3, 512, 11), (2, 12, 1000, 12), (5, 2, 264, 13)])
def test_oracle_fresh_inputs(B, G, L, seed, cfg):
    """Fresh seeded inputs, full forward/backward against the CPU oracle (same weights)."""
    from network import build_loss
    dev = torch.device("cuda:0")
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed)
    m = _model(G, P, dev)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    random.seed(seed)
    d = _to(inp, dev)
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    losses = build_loss(cfg)(outs[0], outs[1], outs[2], d["target"], cfg)
    losses[0].backward()
    Po = {k: v.clone() for k, v in P.items()}
    for n in O.live_param_names(G):
        Po[n].requires_grad_(True)
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                   lead_choice=(c1, c2), stats_out=stats)
    ol = O.standin_loss(*oo, inp["target"])
    ol[0].backward()
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.cpu().numpy(), b.detach().numpy(), rtol=OUT_RTOL, atol=0)
    assert abs(float(losses[0]) - float(ol[0])) < 2e-3 * abs(float(ol[0]))
    named = dict(m.named_parameters())
    for n in O.live_param_names(G):
        if n in O.ZERO_GRAD_PARAMS:
            continue
        ref = Po[n].grad
        got = named[n].grad.cpu()
        err = float((got - ref).norm() / (ref.norm() + 1e-30))
        assert err < GRAD_REL_L2, (n, err)
    sd = m.state_dict()
    for k, v in stats.items():
        np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=2e-3, atol=2e-4)


def test_dropout_statistics():
    """Dropout on: ~20% of the block activations are zeroed and survivors scaled; outputs stay finite and
    two different steps draw different masks."""
    dev = torch.device("cuda:0")
    G, B, L = 2, 3, 256
    P = O.make_params(G, 5)
    d = _to(O.make_inputs(B, G, L, 5), dev)
    m = _model(G, P, dev, dropout=0.2)
    a = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")[0]
    b = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")[0]
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert float((a - b).abs().max()) > 0
    m.eval()
    with torch.no_grad():
        c = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")[0]
        e = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")[0]
    # eval: no dropout; the two lead draws differ but out (mean latents) does not depend on them
    assert float((c - e).abs().max()) == 0.0


def test_cpu_input_raises():
    import network
    m = network.Model_nefnet(1, 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64), torch.zeros(1, 1, 2), torch.zeros(1, 2), torch.zeros(1, 7, 2, dtype=torch.long))
