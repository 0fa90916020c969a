"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference/codes, which only exists in the build container) on the seeded synthetic inputs
and weights of oracle/nefnet_oracle.py, and cross-check the oracle restatement against it.

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference errors

TEST INFRASTRUCTURE ONLY.  The vectors are committed; nothing at test/bench time reads
/root/reference.
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

CASES = [
    # name, B, G, L, seed, phase, V, ragged
    ("train_b2_g1_l128", 2, 1, 128, 1, "train", 0, False),
    ("train_b3_g3_l512_ragged", 3, 3, 512, 2, "train", 0, True),
    ("train_b1_g12_l5000", 1, 12, 5000, 3, "train", 0, False),
    ("test_b2_g3_l512_v4", 2, 3, 512, 4, "test", 4, False),
    ("train_b4_g2_l64", 4, 2, 64, 5, "train", 0, True),
]


def sample_idx(numel: int, n: int = 64) -> np.ndarray:
    return ((np.arange(n, dtype=np.int64) * 2654435761 + 12345) % max(numel, 1)).astype(np.int64)


class _Cfg:  # minimal stand-in for the yacs node losswrapper reads (losses.py:26-44)
    class SOLVER:
        reg_loss = "l1_loss"
        loss_using = [1, 2, 3]
        loss_factor = [0.5, 0.5, 1]


def build_reference(G, P):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from network.model_nefnet import Model_nefnet  # the reference, unmodified

    torch.manual_seed(0)
    m = Model_nefnet(theta_encoder_len=1, lead_num=G).float()
    missing = m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(m.state_dict().keys()) == list(O.param_shapes(G).keys()), "state_dict key order differs"
    return m


def run_case(name, B, G, L, seed, phase, V, ragged):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from network.loss.losses import losswrapper  # noqa: reference

    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed, V=V, ragged_rois=ragged)
    m = build_reference(G, P)
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
    # --- reference
    if phase == "train":
        out, op, ol = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
        losses = losswrapper(out, op, ol, inp["target"], _Cfg)
        losses[0].backward()
        ref_out = (out, op, ol)
    else:
        with torch.no_grad():
            out, op, ol, rest = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                                  rest_theta=inp["rest_theta"], phase="test")
            losses = losswrapper(out, op, ol, inp["target"], _Cfg, rest, inp["rest_view"])
            random.seed(seed)
            z1, z2 = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
            gen = m.gen_ecg(z1, z2, inp["rest_theta"], inp["rois"])
        ref_out = (out, op, ol, rest)
        rec["gen_z1_sample"] = z1.flatten()[sample_idx(z1.numel(), 256)].numpy()
        rec["gen_z2_sample"] = z2.flatten()[sample_idx(z2.numel(), 256)].numpy()
        rec["gen_ecg"] = gen.numpy()
    for i, t in enumerate(ref_out):
        rec[f"out{i}"] = t.detach().numpy()
    rec["losses"] = np.array([float(v.detach()) for v in losses])
    sd = m.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            rec["bn/" + k] = v.numpy().copy()
    # --- oracle, same inputs
    Po = {k: v.clone() for k, v in P.items()}
    names = O.live_param_names(G)
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    errs = {}
    if phase == "train":
        for n in names:
            Po[n].requires_grad_(True)
        oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                       lead_choice=(c1, c2), stats_out=stats)
        ol_ = O.standin_loss(*oo, inp["target"])
        ol_[0].backward()
        gerr = 0.0
        for n, p in m.named_parameters():
            if n in O.UNUSED_PARAMS:
                assert p.grad is None, n
                continue
            g = p.grad
            rec["gs/" + n] = g.flatten()[sample_idx(g.numel())].numpy()
            rec["gn/" + n] = np.array([float(g.double().norm()), float(g.double().sum())])
            if n in O.ZERO_GRAD_PARAMS:  # true gradient is exactly 0; both sides hold rounding noise
                assert float(g.abs().max()) < 1e-5 and float(Po[n].grad.abs().max()) < 1e-5, n
                continue
            e = float((Po[n].grad - g).norm() / float(g.norm()))
            gerr = max(gerr, e)
        errs["grad_rel_l2_max"] = gerr
    else:
        with torch.no_grad():
            oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                           rest_theta=inp["rest_theta"], phase="test", lead_choice=(c1, c2), bn_training=False,
                           stats_out=stats)
            ol_ = O.standin_loss(*oo[:3], inp["target"], rest_out=oo[3], rest_view=inp["rest_view"])
            z1o, z2o = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen")
            geno = O.gen_ecg(Po, z1o, z2o, inp["rest_theta"], inp["rois"])
        errs["gen_ecg"] = float((geno - gen).abs().max())
        errs["gen_z2"] = float((z2o - z2).abs().max() / z2.abs().max())
    for i, (a, b) in enumerate(zip(oo, ref_out)):
        errs[f"out{i}_maxrel"] = float(((a - b).abs() / b.abs()).max())
    errs["loss"] = float(abs(float(ol_[0]) - float(losses[0])))
    for k, v in stats.items():
        errs["bn"] = max(errs.get("bn", 0.0), float((v.double() - sd[k].double()).abs().max()))
    print(f"{name}: lead_choice=({c1},{c2}) oracle-vs-reference " +
          " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    # gradients: a ReLU pre-activation within one ulp of 0 can flip its mask between two fp32
    # evaluation orders (seen on train_b4_g2_l64: 1 flip in 229k -> 2.4e-4 on the upstream weights,
    # while an fp64 run of the oracle agrees with its own fp32 run to 7e-7), hence the looser bound
    assert all(v < (5e-4 if k.startswith('grad') else 2e-5) for k, v in errs.items()), errs
    return rec


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for case in CASES:
        rec = run_case(*case)
        np.savez_compressed(os.path.join(out_dir, case[0] + ".npz"), **rec)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
