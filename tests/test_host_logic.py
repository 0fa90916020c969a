"""CPU: host-side logic of the package that needs no GPU -- the optimiser's torch.optim surface, the synthetic batch
generator, record packing."""
import math

import numpy as np
import pytest
import torch


class _Dummy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3))
        self.flat_params = None
        self.flat_grads = None


def test_flat_sgd_is_a_torch_optimizer_driven_by_the_reference_schedulers():
    """optim_scheduler.py:13-18 builds StepLR(optim, 50, 0.1) / MultiStepLR(optim, lr_step, 0.1) around the optimiser."""
    from network.optim import FlatSGD, get_lr_scheduler, get_optimizer

    class Cfg:
        class SOLVER:
            optim = "sgd"
            lr = 0.1
            scheduler = "MultiStep"
            lr_step = [2, 4]
    opt = get_optimizer(Cfg, _Dummy())
    assert isinstance(opt, FlatSGD) and isinstance(opt, torch.optim.Optimizer)
    sch = get_lr_scheduler(Cfg, opt)
    lrs = []
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # "scheduler.step() before optimizer.step()": no GPU step here
        for _ in range(5):
            lrs.append(opt.lr)
            sch.step()
    assert lrs == pytest.approx([0.1, 0.1, 0.01, 0.01, 0.001])
    sd = opt.state_dict()
    assert sd["flat_momentum"] is None and sd["param_groups"][0]["momentum"] == 0.9
    with pytest.raises(RuntimeError):
        opt.step()                        # no flat buffers yet: loud, not silent
    Cfg.SOLVER.scheduler = "steplr"
    assert get_lr_scheduler(Cfg, get_optimizer(Cfg, _Dummy())).step_size == 50


def test_synthetic_batches_have_the_contract_shapes():
    """SURVEY 8(b) input contract: x in [0, 1], int64 ROIs tiling [0, L] in input-sample units, radians."""
    from dataset.synthetic import make_inputs
    d = make_inputs(5, 12, 5000, seed=3, V=24)
    assert d["x"].shape == (5, 12, 5000) and d["x"].dtype == torch.float32
    assert 0.0 <= float(d["x"].min()) and float(d["x"].max()) <= 1.0
    r = d["rois"]
    assert r.dtype == torch.int64 and r.shape == (5, 7, 2)
    assert bool((r[:, 0, 0] == 0).all()) and bool((r[:, -1, 1] == 5000).all()) and bool((r[:, 1:, 0] == r[:, :-1, 1]).all())
    assert bool((r % 4 == 0).all())
    assert d["rest_theta"].shape == (5, 24, 2) and float(d["rest_theta"].abs().max()) <= math.pi + 1e-6
    assert d["input_thetas"].shape == (5, 12, 2) and d["query_theta"].shape == (5, 2) and d["target"].shape == (5, 1, 5000)
    again = make_inputs(5, 12, 5000, seed=3, V=24)
    assert all(torch.equal(d[k], again[k]) for k in d)


def test_pack_records_offsets():
    from dataset.tianchi import LEAD_THETA, pack_records
    recs = [np.arange(8 * t, dtype=np.int64).reshape(8, t) for t in (5, 9, 4)]
    raw, off, lens = pack_records(recs, "cpu")
    assert raw.dtype == torch.float64 and raw.numel() == 8 * (5 + 9 + 4)
    assert off.tolist() == [0, 40, 112] and lens.tolist() == [5, 9, 4]
    assert float(raw[off[1] + 9 * 2 + 3]) == float(recs[1][2, 3])
    assert LEAD_THETA.shape == (12, 2)
