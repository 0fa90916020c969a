"""Generate tests/golden/nefnet2_*.npz by running the UNMODIFIED reference ``Model_nefnet2``
(/root/reference/codes/network/model_nefnet2.py, importable only in the build container) and cross-check
oracle/nefnet2_oracle.py against it.

    python oracle/make_golden_nefnet2.py

TEST INFRASTRUCTURE ONLY.  The vectors are committed; nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("NEF_REFERENCE", "/root/reference/codes")

from oracle import nefnet_oracle as O  # noqa: E402
from oracle import nefnet2_oracle as O2  # noqa: E402
from oracle.make_golden import _Cfg, sample_idx  # noqa: E402

CASES = [
    # name, B, lead_num, L, seed, phase, V, ragged
    ("nefnet2_train_b2_g3_l256", 2, 3, 256, 11, "train", 0, True),
    ("nefnet2_test_b2_g2_l128_v3", 2, 2, 128, 12, "test", 3, False),
]


def run_case(name, B, G, L, seed, phase, V, ragged):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from network.loss.losses import losswrapper  # noqa: reference
    from network.model_nefnet2 import Model_nefnet2  # the reference, unmodified

    P = O2.make_params(seed)
    inp = O.make_inputs(B, G, L, seed, V=V, ragged_rois=ragged)
    torch.manual_seed(0)
    m = Model_nefnet2(theta_encoder_len=1, lead_num=G).float()
    m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert list(m.state_dict().keys()) == list(O2.param_shapes().keys()), "state_dict key order differs"
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()  # exact parity is defined with dropout off
    if phase == "test":
        m.eval()
    random.seed(seed)
    c1 = random.randint(0, G - 1)
    c2 = random.randint(0, G - 1)
    random.seed(seed)
    rec = dict(B=B, G=G, L=L, seed=seed, V=V, ragged=int(ragged), lead_choice=np.array([c1, c2]))
    Po = {k: v.clone() for k, v in P.items()}
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    errs = {}
    if phase == "train":
        ref_out = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
        losses = losswrapper(*ref_out, inp["target"], _Cfg)
        losses[0].backward()
        for n in O2.live_param_names():
            Po[n].requires_grad_(True)
        oo = O2.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                        lead_choice=(c1, c2), stats_out=stats)
        ol_ = O.standin_loss(*oo, inp["target"])
        ol_[0].backward()
        gerr = 0.0
        for n, p in m.named_parameters():
            if n in O2.UNUSED_PARAMS:
                assert p.grad is None, n
                continue
            g = p.grad
            rec["gs/" + n] = g.flatten()[sample_idx(g.numel())].numpy()
            rec["gn/" + n] = np.array([float(g.double().norm()), float(g.double().sum())])
            if n in O.ZERO_GRAD_PARAMS:
                assert float(g.abs().max()) < 1e-5 and float(Po[n].grad.abs().max()) < 1e-5, n
                continue
            gerr = max(gerr, float((Po[n].grad - g).norm() / float(g.norm())))
        errs["grad_rel_l2_max"] = gerr
    else:
        with torch.no_grad():
            ref_out = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                        rest_theta=inp["rest_theta"], phase="test")
            losses = losswrapper(*ref_out[:3], inp["target"], _Cfg, ref_out[3], inp["rest_view"])
            z1, z2 = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
            oo = O2.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                            rest_theta=inp["rest_theta"], phase="test", lead_choice=(c1, c2), bn_training=False,
                            stats_out=stats)
            ol_ = O.standin_loss(*oo[:3], inp["target"], rest_out=oo[3], rest_view=inp["rest_view"])
            z1o, z2o = O2.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
        rec["gen_z1"] = z1.numpy()
        rec["gen_z2"] = z2.numpy()
        errs["gen_z1"] = float((z1o - z1).abs().max() / z1.abs().max())
        errs["gen_z2"] = float((z2o - z2).abs().max() / z2.abs().max())
    for i, t in enumerate(ref_out):
        rec[f"out{i}"] = t.detach().numpy()
    rec["losses"] = np.array([float(v.detach()) for v in losses])
    sd = m.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            rec["bn/" + k] = v.numpy().copy()
    for i, (a, b) in enumerate(zip(oo, ref_out)):
        errs[f"out{i}_maxrel"] = float(((a - b).abs() / b.abs()).max())
    errs["loss"] = float(abs(float(ol_[0]) - float(losses[0])))
    for k, v in stats.items():
        errs["bn"] = max(errs.get("bn", 0.0), float((v.double() - sd[k].double()).abs().max()))
    print(f"{name}: lead_choice=({c1},{c2}) oracle-vs-reference " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert all(v < (5e-4 if k.startswith("grad") else 2e-5) for k, v in errs.items()), errs
    return rec


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for case in CASES:
        rec = run_case(*case)
        np.savez_compressed(os.path.join(out_dir, case[0] + ".npz"), **rec)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
