"""GPU: the call sequences of the reference's callers (solver/solver.py Solver, demo.ipynb Generator) replayed line by
line on this package's `network` and checked against the CPU oracle driven the same way.  The reference itself cannot
travel to the GPU box; tests/test_dropin_reference.py runs its unmodified Solver on this `network` in the build container.

The assertions below were first validated on the CPU with the REFERENCE's own Model_nefnet / losswrapper substituted for
this package's (a scratch script that registers a shim `network` module and replaces "cuda:0" by "cpu"): the reference
passes every check with loss trajectories equal to the oracle's to 1e-7, so a failure here is the CUDA path's."""
import random

import numpy as np
import pytest
import torch

from oracle import data_oracle as D
from oracle import nefnet_oracle as O

pytestmark = pytest.mark.gpu

OUT_RTOL = 1e-3   # north star


def _batch(B, G, L, seed, V, dev):
    inp = O.make_inputs(B, G, L, seed, V=V)
    return inp, {k: v.to(dev) for k, v in inp.items()}


def test_solver_shaped_epochs(cfg, tmp_path):
    """Solver.__init__ (solver.py:20-38), train epoch (:141-235 with the reference's own SGD factory,
    optim_scheduler.py:10), CheckPointer.save / load (checkpointer.py:17-70), test epoch (:171-205) -- in that order."""
    import network
    from utils import mertic as M
    dev = torch.device("cuda:0")
    G, B, L, V, seed = 3, 4, 512, 8, 31
    torch.manual_seed(0)
    model = network.build_model(cfg).float()                       # solver.py:20
    loss_fn = network.build_loss(cfg)                              # :21
    model.to(dev)                                                  # :38
    P = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}   # the oracle starts from the same weights
    optim = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)          # optim_scheduler.py:10, nef_net.yml lr
    sched = torch.optim.lr_scheduler.MultiStepLR(optim, [50, 100], gamma=0.1)  # :17-18
    model.dropout_p = 0.0     # exact parity is defined with dropout off (the oracle takes no mask here)
    mom = {}
    model.train()                                                  # :143
    for it in range(3):
        inp, d = _batch(B, G, L, seed + it, V, dev)
        random.seed(seed + it)
        c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
        random.seed(seed + it)
        result = model(d["x"], d["input_thetas"], d["query_theta"], d["rois"], rest_theta=d["rest_theta"], phase="train")  # :171
        out, out_p, out_l = result                                 # :177
        loss, l1, l2, l3 = loss_fn(out, out_p, out_l, d["target"], cfg, loss1_gt=None, loss2_gt=None)   # :187
        vals = [loss.item(), l1.item(), l2.item(), l3.item()]      # :189
        assert all(np.isfinite(vals)) and abs(vals[0] - (vals[1] + vals[2] + vals[3])) < 1e-5
        loss.backward()                                            # :233
        optim.step()                                               # :234
        optim.zero_grad()                                          # :235
        ref_loss, _ = O.train_step(P, inp, lead_choice=(c1, c2), lr=0.1, momentum=0.9, momentum_buf=mom)
        rel = abs(vals[0] - ref_loss) / ref_loss
        print("step %d loss %.6f oracle %.6f rel %.2e" % (it, vals[0], ref_loss, rel))
        assert rel < (2e-3 if it == 0 else 2e-2), (it, vals[0], ref_loss)   # step 0: forward parity; later steps also carry
        # the TF32-vs-fp32 difference of the applied gradients
    sched.step()
    sd = model.state_dict()
    assert int(sd["decoder.1.double_conv.1.num_batches_tracked"]) == 9          # three decoder passes per step
    for n in O.UNUSED_PARAMS:
        assert dict(model.named_parameters())[n].grad is None
    for k in sd:                                                   # 9 running-statistics updates on updated weights
        if "running_" in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), P[k].numpy(), rtol=1e-2, atol=1e-3, err_msg=k)
    assert all(bool(torch.isfinite(v).all()) for v in sd.values())

    path = tmp_path / "best_valid.pkl"                              # checkpointer.py:17-35
    torch.save({"optimizer": optim.state_dict(), "scheduler": sched.state_dict(), "model": model.state_dict(),
                "epoch": 0, "best_test_psnr_gen": 1.0}, path)

    def test_epoch(m):                                              # solver.py:145 + :171-205
        m.eval()
        inp, d = _batch(B, G, L, seed + 10, V, dev)
        random.seed(seed + 10)
        with torch.no_grad():                                       # :121
            out, out_p, out_l, rest_out = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"],
                                            rest_theta=d["rest_theta"], phase="test")
            losses = loss_fn(out, out_p, out_l, d["target"], cfg, rest_out[:, -4:, :], d["rest_view"][:, -4:, :])   # :192
        assert len(losses) == 5 and all(np.isfinite(v.item()) for v in losses)
        psnr_gen = M.PSNR(rest_out[:, -4:, :].contiguous().cpu().detach().numpy(),                                   # :211
                          d["rest_view"][:, -4:, :].contiguous().cpu().detach().numpy(), d["rois"].cpu().detach().numpy())
        return inp, (out, out_p, out_l, rest_out), losses, psnr_gen

    inp, outs, losses, psnr_gen = test_epoch(model)
    Pg = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    random.seed(seed + 10)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    with torch.no_grad():
        ref = O.forward(Pg, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"],
                        phase="test", lead_choice=(c1, c2), bn_training=False)
        ref_unsup = O.standin_loss(*ref[:3], inp["target"], rest_out=ref[3][:, -4:, :], rest_view=inp["rest_view"][:, -4:, :])[4]
    for a, b in zip(outs, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=OUT_RTOL, atol=0)
    assert abs(losses[4].item() - float(ref_unsup)) < 2e-3 * float(ref_unsup)
    rows = D.psnr_rows(outs[3][:, -4:, :].cpu().numpy(), inp["rest_view"][:, -4:, :].numpy(), inp["rois"].numpy())
    assert abs(psnr_gen - float(np.mean(rows))) < 1e-3

    model2 = network.build_model(cfg).float()                       # a fresh process resuming (checkpointer.py:37-70)
    model2.to(dev)
    ck = torch.load(path)
    model2.load_state_dict(ck.pop("model"))
    optim2 = torch.optim.SGD(model2.parameters(), lr=0.1, momentum=0.9)
    optim2.load_state_dict(ck.pop("optimizer"))
    assert ck["epoch"] == 0 and len(optim2.state_dict()["state"]) == len(O.live_param_names(G))
    _, outs2, _, psnr2 = test_epoch(model2)
    for a, b in zip(outs, outs2):
        assert torch.equal(a, b)                                    # same weights, eval mode: bit-identical
    assert psnr2 == psnr_gen


def test_generator_flow_of_the_demo_notebook():
    """demo.ipynb `Generator.valid`: phase='test' on a module that was never switched to eval() and outside no_grad -- the
    decoder BatchNorms use batch statistics and update their running statistics once per decoder call (3 + V)."""
    import network
    dev = torch.device("cuda:0")
    G, B, L, V, seed = 3, 2, 512, 5, 41
    P = O.make_params(G, seed)
    inp, d = _batch(B, G, L, seed, V, dev)
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=G)
    m.load_state_dict({k: v.clone() for k, v in P.items()})
    m = m.float().to(dev)
    assert m.training
    m.dropout_p = 0.0
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    random.seed(seed)
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], rest_theta=d["rest_theta"], phase="test")
    assert len(outs) == 4 and outs[3].shape == (B, V, L)
    Po = {k: v.clone() for k, v in P.items()}
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    with torch.no_grad():
        ref = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"],
                        phase="test", lead_choice=(c1, c2), bn_training=True, stats_out=stats)
    for a, b in zip(outs, ref):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.numpy(), rtol=OUT_RTOL, atol=0)
    sd = m.state_dict()
    assert int(sd["decoder.3.double_conv.4.num_batches_tracked"]) == 3 + V
    for k, v in stats.items():
        np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=2e-3, atol=2e-4)
