"""Pins oracle/nefnet_oracle.py to the vectors the unmodified reference produced
(tests/golden/*.npz, written by oracle/make_golden.py inside the build container)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import nefnet_oracle as O
from oracle.make_golden import sample_idx

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "t*_b*.npz")))  # the hot-path vectors (data_*.npz: test_data_oracle.py)


def test_golden_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_vectors(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, G, L, seed, V = (int(g[k]) for k in ("B", "G", "L", "seed", "V"))
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed, V=V, ragged_rois=bool(int(g["ragged"])))
    lead = tuple(int(v) for v in g["lead_choice"])
    stats = {k: v for k, v in P.items() if "running_" in k or "num_batches" in k}
    train = name.startswith("train")
    if train:
        for n in O.live_param_names(G):
            P[n].requires_grad_(True)
        outs = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                         lead_choice=lead, stats_out=stats)
        losses = O.standin_loss(*outs, inp["target"])
        losses[0].backward()
    else:
        with torch.no_grad():
            outs = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                             rest_theta=inp["rest_theta"], phase="test", lead_choice=lead, bn_training=False,
                             stats_out=stats)
            losses = O.standin_loss(*outs[:3], inp["target"], rest_out=outs[3], rest_view=inp["rest_view"])
            z1, z2 = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
            gen = O.gen_ecg(P, z1, z2, inp["rest_theta"], inp["rois"])
        np.testing.assert_allclose(gen.numpy(), g["gen_ecg"], rtol=2e-6, atol=0)
        np.testing.assert_allclose(z1.flatten()[sample_idx(z1.numel(), 256)].numpy(), g["gen_z1_sample"], rtol=1e-5,
                                   atol=1e-6)
        np.testing.assert_allclose(z2.flatten()[sample_idx(z2.numel(), 256)].numpy(), g["gen_z2_sample"], rtol=1e-5,
                                   atol=1e-6)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().numpy(), g[f"out{i}"], rtol=2e-6, atol=0)  # north-star bar is 1e-3
    np.testing.assert_allclose(np.array([float(v.detach()) for v in losses]), g["losses"], rtol=1e-6, atol=1e-8)
    for k, v in stats.items():
        np.testing.assert_allclose(v.numpy(), g["bn/" + k], rtol=1e-5, atol=1e-6)
    if train:
        assert int(stats["decoder.1.double_conv.1.num_batches_tracked"]) == 3  # three decoder passes per step
        for n in O.live_param_names(G):
            gr = P[n].grad
            if n in O.ZERO_GRAD_PARAMS:
                assert float(gr.abs().max()) < 1e-5
                continue
            norm_ref = float(g["gn/" + n][0])
            got = gr.flatten()[sample_idx(gr.numel())].numpy()
            # ReLU-mask flips at |pre-activation| < 1 ulp bound the agreement (see make_golden.py)
            np.testing.assert_allclose(got, g["gs/" + n], rtol=2e-3, atol=2e-3 * norm_ref / np.sqrt(gr.numel()) + 1e-9)
            assert abs(float(gr.double().norm()) - norm_ref) <= 5e-4 * norm_ref
        for n in O.UNUSED_PARAMS:
            assert P[n].grad is None


def test_state_dict_contract():
    """Key set / shapes / parameter counts of SURVEY 8(b)."""
    for G, n in ((1, 2702081), (3, 7626369), (12, 29785665)):
        shapes = O.param_shapes(G)
        total = sum(int(np.prod(s)) for k, s in shapes.items() if "running_" not in k and "num_batches" not in k)
        assert total == n
    assert O.param_shapes(3)["z2_conv2.1.weight"] == (2688, 64, 2)


def test_roi_edge_cases():
    """Empty rois contribute nothing; truncated lengths must tile L/4 (roi_pooling_1d.py:85-98)."""
    rois = torch.tensor([[[0, 16], [16, 16], [16, 30], [30, 33], [33, 36], [36, 36], [36, 64]]])
    z = torch.arange(7 * 32, dtype=torch.float32).view(1, 1, 7, 32)
    out = O.roi_reverse(z, rois)
    assert out.shape == (1, 1, 16)
    a = O.roi_align_center(torch.ones(1, 2, 16), rois)
    assert a.shape == (1, 2, 7, 16)
    assert float(a.max()) <= 1.0 and float(a.min()) >= 0.5


NEFNET2 = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "nefnet2_*.npz")))


@pytest.mark.parametrize("name", NEFNET2)
def test_nefnet2_oracle_matches_reference_vectors(name, golden_dir):
    """the shared-trunk variant (model_nefnet2.py:63-203) restated as the 1-lead trunk over leads
    folded into the batch, pinned to the unmodified reference class (oracle/make_golden_nefnet2.py)."""
    from oracle import nefnet2_oracle as O2
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, G, L, seed, V = (int(g[k]) for k in ("B", "G", "L", "seed", "V"))
    P = O2.make_params(seed)
    inp = O.make_inputs(B, G, L, seed, V=V, ragged_rois=bool(int(g["ragged"])))
    lead = tuple(int(v) for v in g["lead_choice"])
    stats = {k: v for k, v in P.items() if "running_" in k or "num_batches" in k}
    if "train" in name:
        for n in O2.live_param_names():
            P[n].requires_grad_(True)
        outs = O2.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                          lead_choice=lead, stats_out=stats)
        losses = O.standin_loss(*outs, inp["target"])
        losses[0].backward()
        for n in O2.live_param_names():
            gr = P[n].grad
            if n in O.ZERO_GRAD_PARAMS:
                assert float(gr.abs().max()) < 1e-5
                continue
            norm_ref = float(g["gn/" + n][0])
            np.testing.assert_allclose(gr.flatten()[sample_idx(gr.numel())].numpy(), g["gs/" + n], rtol=2e-3,
                                       atol=2e-3 * norm_ref / np.sqrt(gr.numel()) + 1e-9)
            assert abs(float(gr.double().norm()) - norm_ref) <= 5e-4 * norm_ref
        for n in O2.UNUSED_PARAMS:
            assert P[n].grad is None
        assert int(stats["decoder.1.double_conv.1.num_batches_tracked"]) == 3
    else:
        with torch.no_grad():
            outs = O2.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                              rest_theta=inp["rest_theta"], phase="test", lead_choice=lead, bn_training=False,
                              stats_out=stats)
            losses = O.standin_loss(*outs[:3], inp["target"], rest_out=outs[3], rest_view=inp["rest_view"])
            z1, z2 = O2.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
        np.testing.assert_allclose(z1.numpy(), g["gen_z1"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(z2.numpy(), g["gen_z2"], rtol=1e-5, atol=1e-6)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().numpy(), g[f"out{i}"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(np.array([float(v.detach()) for v in losses]), g["losses"], rtol=1e-6, atol=1e-8)
    for k, v in stats.items():
        np.testing.assert_allclose(v.numpy(), g["bn/" + k], rtol=1e-5, atol=1e-6)


def test_nefnet2_state_dict_contract():
    """Key set is independent of lead_num (one shared single-lead trunk): 2 702 081 + 2 x (128*128*3 + 128)."""
    from oracle import nefnet2_oracle as O2
    shapes = O2.param_shapes()
    total = sum(int(np.prod(s)) for k, s in shapes.items() if "running_" not in k and "num_batches" not in k)
    assert total == 2702081 + 2 * (128 * 128 * 3 + 128)
    keys = list(shapes)
    assert keys.index("single_conv_z1.0.weight") == keys.index("decoder.1.double_conv.0.weight") - 4
    assert len(NEFNET2) >= 2
