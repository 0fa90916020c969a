"""GPU parity of the CUDA path (through the reference-facing module surface and the C ABI) against
(1) the golden vectors the unmodified reference produced and (2) the CPU oracle on fresh seeded inputs.

Two tiers:
  * production (tcgen05 TF32 convolutions, activations stored TF32-rounded): synthesized waveforms within
    1e-3 relative of the fp32 reference (the north-star bar).  Gradients of a ReLU network move by a few
    percent in L2 between TF32 and fp32 whatever the implementation (ReLU masks within rounding distance
    of zero flip; measured 2-8 % with a TF32-rounding CPU emulation of the oracle), so they are held to
    cosine >= 0.99 / relative L2 <= 0.15 here ...
  * ... and the backward LOGIC is pinned in the exact tier: CUDA-core convolutions with all TF32 rounding
    switched off (nef_set_exact_fp32) must reproduce the fp32 oracle's gradients to 2e-3 relative L2.
"""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import nefnet_oracle as O
from oracle.make_golden import sample_idx

pytestmark = pytest.mark.gpu

GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "t*_b*.npz")))  # the hot-path vectors (data_*.npz: test_data_oracle.py)
OUT_RTOL = 1e-3          # north star
EXACT_OUT_RTOL = 2e-5
EXACT_GRAD_REL_L2 = 5e-3  # a single ReLU-mask flip at a pre-activation within 1 ulp of 0 costs ~2e-3 (see make_golden.py)
TF32_GRAD_REL_L2 = 0.15
TF32_GRAD_COS = 0.99
GOLDEN_TF32_COS = 0.8   # golden L1-loss gradients at TF32: statistical (sign flips), see test_golden

MODES = {"exact_simt": (0, 1), "tf32_simt": (0, 0), "tf32_tc": (1, 0)}


class mode:
    def __init__(self, name):
        self.impl, self.exact = MODES[name]

    def __enter__(self):
        from network import _native as N
        lib = N.init(0)
        torch.cuda.synchronize()
        lib.nef_set_conv_impl(self.impl)
        N.check(lib.nef_set_exact_fp32(self.exact), "nef_set_exact_fp32")

    def __exit__(self, *a):
        from network import _native as N
        lib = N.load()
        torch.cuda.synchronize()
        lib.nef_set_conv_impl(1)
        lib.nef_set_exact_fp32(0)


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


def _grad_check(named, ref_grads, exact, what):
    worst = (0.0, "")
    for n, ref in ref_grads.items():
        if n in O.ZERO_GRAD_PARAMS:  # exactly zero in exact arithmetic; both sides hold rounding noise
            assert bool(torch.isfinite(named[n].grad).all()), n
            continue
        got = named[n].grad.detach().cpu().double()
        ref = ref.double()
        err = float((got - ref).norm() / (ref.norm() + 1e-30))
        cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
        if err > worst[0]:
            worst = (err, n)
        if exact:
            assert err < EXACT_GRAD_REL_L2, (what, n, err)
        else:
            assert err < TF32_GRAD_REL_L2 and cos > TF32_GRAD_COS, (what, n, err, cos)
    print(what, "worst grad rel-L2 %.3e (%s)" % worst)


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("name", GOLDEN)
def test_golden(name, mode_name, golden_dir, cfg):
    """Vectors produced by the unmodified reference (oracle/make_golden.py)."""
    from network import build_loss
    dev = torch.device("cuda:0")
    exact = MODES[mode_name][1] == 1
    out_rtol = EXACT_OUT_RTOL if exact else OUT_RTOL
    with mode(mode_name):
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
            np.testing.assert_allclose(gen.cpu().numpy(), g["gen_ecg"], rtol=out_rtol, atol=0)
            ztol = 2e-5 if exact else 5e-3
            for z, key in ((z1, "gen_z1_sample"), (z2, "gen_z2_sample")):
                zz = z.cpu().flatten()[sample_idx(z.numel(), 256)].numpy()
                np.testing.assert_allclose(zz, g[key], rtol=ztol, atol=ztol * float(np.abs(g[key]).max()))
        for i, o in enumerate(outs):
            np.testing.assert_allclose(o.detach().cpu().numpy(), g[f"out{i}"], rtol=out_rtol, atol=0)
        np.testing.assert_allclose(np.array([float(v.detach()) for v in losses]), g["losses"],
                                   rtol=1e-5 if exact else 2e-3, atol=1e-5)  # atol: BN statistics are summed
        # with atomics, so two decoder calls on identical latents (G = 1) agree to ~1e-6, not bit-exactly
        sd = m.state_dict()
        for k in g.files:
            if k.startswith("bn/"):
                np.testing.assert_allclose(sd[k[3:]].cpu().numpy(), g[k], rtol=1e-4 if exact else 2e-3,
                                           atol=1e-5 if exact else 2e-4)
        if train:
            assert int(sd["decoder.1.double_conv.1.num_batches_tracked"]) == 3
            named = dict(m.named_parameters())
            for n in O.UNUSED_PARAMS:
                assert named[n].grad is None
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
                nerr = abs(float(gr.double().norm()) - norm_ref) / norm_ref
                if exact:
                    # The golden gradients come from the L1 Standin loss: d|a - b| = sign(a - b) flips wherever two
                    # predictions agree to within fp32 rounding (and ReLU masks flip at |pre-activation| < 1 ulp),
                    # each flip moving a 1-D gradient (a bias: one sum over positions) by ~1/L of its norm.
                    # The flip-free check of the backward logic is test_oracle_fixed_upstream (2e-3 relative L2).
                    np.testing.assert_allclose(got, ref, rtol=1e-2, atol=1e-2 * norm_ref / np.sqrt(gr.numel()) + 1e-9)
                    # (a bias gradient is ONE sum over positions: two sign flips at B = 1 move it by 0.5 %)
                    assert nerr < (1e-2 if gr.dim() == 1 else 5e-3), (n, nerr)
                else:      # L1 loss: sign(out - target) flips make the comparison statistical (B = 1: few samples);
                    # the TF32-tier gradient bar proper is test_oracle_fixed_upstream (cosine >= 0.99)
                    assert nerr < (0.3 if gr.dim() == 1 else 0.15), (n, nerr)   # a bias gradient is ONE sum over positions
                    cos = float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
                    if os.environ.get("NEF_TEST_VERBOSE"):
                        print("golden-cos", name, n, "%.4f" % cos)
                    assert cos > GOLDEN_TF32_COS, (n, cos)


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("B,G,L,seed,ragged", [(4, 3, 512, 11, False), (2, 12, 1000, 12, False), (5, 2, 264, 13, True),
                                               (3, 1, 16, 14, False)])
def test_oracle_fixed_upstream(B, G, L, seed, ragged, mode_name):
    """Fresh seeded inputs; backward driven by fixed upstream gradients (no loss discontinuity)."""
    dev = torch.device("cuda:0")
    exact = MODES[mode_name][1] == 1
    P = O.make_params(G, seed)
    if L == 16:
        inp = O.make_inputs(B, G, 64, seed)
        inp = {k: (v[..., :16].contiguous() if k in ("x", "target") else v) for k, v in inp.items()}
        inp["rois"] = torch.tensor([[0, 4], [4, 4], [4, 8], [8, 8], [8, 12], [12, 12], [12, 16]]).repeat(B, 1, 1)
    else:
        inp = O.make_inputs(B, G, L, seed, ragged_rois=ragged)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    gen = torch.Generator().manual_seed(seed)
    ups = [torch.randn(B, 1, inp["x"].shape[-1], generator=gen) for _ in range(3)]
    with mode(mode_name):
        m = _model(G, P, dev)
        random.seed(seed)
        d = _to(inp, dev)
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        torch.autograd.backward(outs, [u.to(dev) for u in ups])
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        named = dict(m.named_parameters())
    Po = {k: v.clone() for k, v in P.items()}
    for n in O.live_param_names(G):
        Po[n].requires_grad_(True)
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                   lead_choice=(c1, c2), stats_out=stats)
    torch.autograd.backward(oo, ups)
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().numpy(), rtol=EXACT_OUT_RTOL if exact else OUT_RTOL,
                                   atol=0)
    _grad_check(named, {n: Po[n].grad for n in O.live_param_names(G)}, exact, "%s B%d G%d L%d" % (mode_name, B, G, L))
    for k, v in stats.items():
        np.testing.assert_allclose(sd[k].numpy(), v.numpy(), rtol=1e-4 if exact else 2e-3, atol=1e-5 if exact else 2e-4)


@pytest.mark.parametrize("mode_name", ["exact_simt", "tf32_tc"])
def test_eval_mode_backward(mode_name):
    """module.eval() + phase='train' under autograd (a caller fine-tuning with frozen BatchNorm statistics): the forward uses
    the running statistics, so the BatchNorm gradient has no batch-mean terms (dc = gamma * invstd * g) and the convolution
    biases in front of it DO get gradients.  Exact tier against the oracle with bn_training=False."""
    dev = torch.device("cuda:0")
    exact = MODES[mode_name][1] == 1
    B, G, L, seed = 3, 2, 256, 19
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    gen = torch.Generator().manual_seed(seed)
    ups = [torch.randn(B, 1, L, generator=gen) for _ in range(3)]
    with mode(mode_name):
        m = _model(G, P, dev, train=False)
        random.seed(seed)
        d = _to(inp, dev)
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        torch.autograd.backward(outs, [u.to(dev) for u in ups])
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        named = dict(m.named_parameters())
    Po = {k: v.clone() for k, v in P.items()}
    for n in O.live_param_names(G):
        Po[n].requires_grad_(True)
    oo = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(c1, c2),
                   bn_training=False)
    torch.autograd.backward(oo, ups)
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().numpy(), rtol=EXACT_OUT_RTOL if exact else OUT_RTOL, atol=0)
    for n in O.live_param_names(G):   # including the conv biases in front of the BatchNorms (ZERO_GRAD_PARAMS in training mode)
        got, ref = named[n].grad.detach().cpu().double(), Po[n].grad.double()
        err = float((got - ref).norm() / (ref.norm() + 1e-30))
        assert err < (EXACT_GRAD_REL_L2 if exact else TF32_GRAD_REL_L2), (n, err)
    for k in P:   # running statistics untouched
        if "running_" in k or "num_batches" in k:
            assert torch.equal(sd[k], P[k]), k


def test_long_sequence_config4_shape():
    """BASELINE config 4 shape (12 leads x 20000 samples, PTB rate) at a batch the CPU oracle finishes in seconds:
    training-mode outputs of the production path within the north-star tolerance."""
    dev = torch.device("cuda:0")
    B, G, L, seed = 1, 12, 20000, 21
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    with mode("tf32_tc"):
        m = _model(G, P, dev)
        random.seed(seed)
        d = _to(inp, dev)
        with torch.no_grad():
            outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        outs = [o.cpu() for o in outs]
    with torch.no_grad():
        oo = O.forward({k: v.clone() for k, v in P.items()}, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                       phase="train", lead_choice=(c1, c2))
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=OUT_RTOL, atol=0)


def test_panorama_sweep_config5_shape():
    """BASELINE config 5 shape: encode once, decode 24 Angular-Encoding query views (eval mode, running statistics),
    through forward(phase='test') and through gen_ecg on the 'gen' latents; both against the oracle."""
    dev = torch.device("cuda:0")
    B, G, L, V, seed = 2, 12, 5000, 24, 22
    P = O.make_params(G, seed)
    for k in P:   # non-trivial running statistics, as after training
        if k.endswith("running_mean"):
            P[k] = 0.05 * torch.randn_like(P[k])
        if k.endswith("running_var"):
            P[k] = 0.5 + torch.rand_like(P[k])
    inp = O.make_inputs(B, G, L, seed, V=V)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    with mode("tf32_tc"):
        m = _model(G, P, dev, train=False)
        random.seed(seed)
        d = _to(inp, dev)
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], rest_theta=d["rest_theta"], phase="test")
        z1, z2 = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="gen")
        gen = m.gen_ecg(z1, z2, d["rest_theta"], d["rois"]).cpu()
        outs = [o.cpu() for o in outs]
    with torch.no_grad():
        oo = O.forward({k: v.clone() for k, v in P.items()}, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"],
                       rest_theta=inp["rest_theta"], phase="test", lead_choice=(c1, c2), bn_training=False)
    assert outs[3].shape == (B, V, L)
    for a, b in zip(outs, oo):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=OUT_RTOL, atol=0)
    np.testing.assert_allclose(gen.numpy(), oo[3].numpy(), rtol=OUT_RTOL, atol=0)


def test_standin_loss_kernels(cfg):
    """nef_loss_fwd / nef_loss_bwd against torch on identical inputs (exact op-level check)."""
    from network import losswrapper
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    for reg in ("l1_loss", "l2_loss"):
        class c2(cfg):
            class SOLVER:
                reg_loss = reg
                loss_using = [1, 2, 3]
                loss_factor = [0.5, 0.25, 1.5]
        ts = [torch.rand(5, 1, 333, generator=gen) for _ in range(4)]
        a = [t.clone().to(dev).requires_grad_(i < 3) for i, t in enumerate(ts)]
        b = [t.clone().requires_grad_(i < 3) for i, t in enumerate(ts)]
        got = losswrapper(a[0], a[1], a[2], a[3], c2)
        ref = O.standin_loss(b[0], b[1], b[2], b[3], factor=(0.5, 0.25, 1.5), reg_loss="l1_loss" if reg == "l1_loss" else "mse")
        np.testing.assert_allclose([float(v.detach()) for v in got], [float(v.detach()) for v in ref], rtol=2e-6)
        got[0].backward()
        ref[0].backward()
        for x, y in zip(a[:3], b[:3]):
            np.testing.assert_allclose(x.grad.cpu().numpy(), y.grad.numpy(), rtol=1e-6, atol=1e-12)
        rest, view = torch.rand(5, 4, 333, generator=gen), torch.rand(5, 4, 333, generator=gen)
        got5 = losswrapper(a[0], a[1], a[2], a[3], c2, rest.to(dev), view.to(dev))
        ref5 = O.standin_loss(b[0], b[1], b[2], b[3], factor=(0.5, 0.25, 1.5),
                              reg_loss="l1_loss" if reg == "l1_loss" else "mse", rest_out=rest, rest_view=view)
        assert len(got5) == 5
        np.testing.assert_allclose(float(got5[4]), float(ref5[4]), rtol=2e-6)


def test_dropout_statistics():
    """Dropout on (train mode): masks differ between steps, the kept fraction of the hidden activations is
    ~0.8 in effect (outputs move), eval mode is deterministic."""
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
    # eval: no dropout; the lead draws differ but `out` (mean latents) does not depend on them
    assert float((c - e).abs().max()) == 0.0


def test_sgd_step_matches_torch():
    from network import _native as N
    dev = torch.device("cuda:0")
    lib = N.init(0)
    gen = torch.Generator().manual_seed(1)
    n = 100003
    p, g = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([ref_p], lr=0.1, momentum=0.9)
    pd, md = p.to(dev), torch.zeros(n, device=dev)
    for it in range(3):
        gi = g * (it + 1)
        ref_p.grad = gi.clone()
        opt.step()
        gd = gi.to(dev)
        N.check(lib.nef_sgd_step(N.ptr(pd), N.ptr(gd), N.ptr(md), n, 0.1, 0.9, 1.0, N.stream_ptr()), "nef_sgd_step")
    np.testing.assert_allclose(pd.cpu().numpy(), ref_p.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_flat_sgd_with_reference_scheduler(cfg):
    """FlatSGD is a torch Optimizer: the reference's StepLR / MultiStepLR (optim_scheduler.py:13-18) drive its learning
    rate, and the trajectory equals torch.optim.SGD(momentum=0.9) + the same scheduler on the same gradients."""
    from network.optim import FlatSGD
    from torch.optim.lr_scheduler import MultiStepLR
    dev = torch.device("cuda:0")
    G, B, L, seed = 2, 2, 64, 5
    P = O.make_params(G, seed)
    inp = _to(O.make_inputs(B, G, L, seed), dev)
    with mode("exact_simt"):
        m = _model(G, P, dev)
        opt = FlatSGD(m, lr=0.1, momentum=0.9)
        sch = MultiStepLR(opt, [1, 2], gamma=0.1)
        ref = {n: torch.nn.Parameter(p.detach().clone()) for n, p in m.named_parameters() if n in O.live_param_names(G)}
        ropt = torch.optim.SGD(list(ref.values()), lr=0.1, momentum=0.9)
        rsch = MultiStepLR(ropt, [1, 2], gamma=0.1)
        lrs = []
        for it in range(3):
            random.seed(seed + it)
            outs = m(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
            network_loss = __import__("network").losswrapper(outs[0], outs[1], outs[2], inp["target"], cfg)[0]
            network_loss.backward()
            for n, p in m.named_parameters():
                if n in ref:
                    ref[n].grad = p.grad.detach().clone()
            lrs.append(opt.lr)
            opt.step(); ropt.step()
            opt.zero_grad(); ropt.zero_grad()
            sch.step(); rsch.step()
        assert lrs == pytest.approx([0.1, 0.01, 0.001])
        got = dict(m.named_parameters())
        for n, r in ref.items():
            np.testing.assert_allclose(got[n].detach().cpu().numpy(), r.detach().cpu().numpy(), rtol=2e-5, atol=1e-6, err_msg=n)
        sd = opt.state_dict()
        assert float(sd["flat_momentum"].abs().max()) > 0 and sd["param_groups"][0]["lr"] == pytest.approx(0.001)


def test_cpu_input_raises():
    import network
    m = network.Model_nefnet(1, 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64), torch.zeros(1, 1, 2), torch.zeros(1, 2), torch.zeros(1, 7, 2, dtype=torch.long))


def test_full_size_segment_independence():
    """BASELINE.json configs[1] size (256 x 12 x 5000): in eval mode every segment is independent, so the synthesized lead of
    a segment must not depend on which batch, row tile or CTA it was computed in.  The same 4 segments are run alone
    (small grids, one-row-tile kernels) and inside the full batch at scattered positions (persistent 256-row-tile
    kernels); outputs must agree to accumulation-order noise, and a batch permutation must permute the outputs exactly."""
    dev = torch.device("cuda:0")
    G, L, B = 12, 5000, 256
    P = O.make_params(G, 3)
    m = _model(G, P, dev, train=False)
    small = O.make_inputs(4, G, L, 17)
    big = {k: v.repeat(*([B // 4] + [1] * (v.dim() - 1))).contiguous() for k, v in small.items()}
    gen = torch.Generator().manual_seed(5)
    big["x"] = torch.rand(B, G, L, generator=gen)
    pos = [3, 77, 130, 255]
    for i, p_ in enumerate(pos):
        for k in ("x", "input_thetas", "query_theta", "rois"):
            big[k][p_] = small[k][i]
    with torch.no_grad():
        random.seed(1)
        ds = _to(small, dev)
        ref = m(ds["x"], ds["input_thetas"], ds["query_theta"], ds["rois"], phase="train")[0].cpu()
        random.seed(1)
        db = _to(big, dev)
        out = m(db["x"], db["input_thetas"], db["query_theta"], db["rois"], phase="train")[0].cpu()
        perm = torch.randperm(B, generator=gen)
        dp = {k: v[perm].contiguous() for k, v in db.items()}
        random.seed(1)
        outp = m(dp["x"], dp["input_thetas"], dp["query_theta"], dp["rois"], phase="train")[0].cpu()
    assert torch.isfinite(out).all()
    np.testing.assert_allclose(out[pos].numpy(), ref.numpy(), rtol=2e-5, atol=0)
    assert torch.equal(outp, out[perm])
