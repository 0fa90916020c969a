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


def test_checkpoint_interchange_with_the_reference_state_dict(tmp_path):
    """checkpointer.py:28,79-81 saves / restores `model.state_dict()` with torch.save / load_state_dict: a reference-shaped
    checkpoint loads strictly, survives .float() (solver.py:22) and a save / load round trip bit for bit."""
    import network
    from oracle import nefnet_oracle as O
    P = O.make_params(2, seed=9)
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=2).float()
    res = m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    path = tmp_path / "best_valid.pkl"
    torch.save({"model": m.state_dict(), "epoch": 3}, path)
    back = torch.load(path, map_location="cpu")
    m2 = network.Model_nefnet(theta_encoder_len=1, lead_num=2)
    m2.load_state_dict(back["model"])
    for k, v in m2.state_dict().items():
        assert v.dtype == P[k].dtype and torch.equal(v, P[k]), k
    assert m2.state_dict()["decoder.1.double_conv.1.num_batches_tracked"].dtype == torch.long
    with pytest.raises(RuntimeError):   # a 3-lead checkpoint does not fit a 2-lead model: same loud failure as nn.Module
        m2.load_state_dict(O.make_params(3, seed=9))


def test_default_initialisation_scales():
    """resnet_1d.py:114-120: encoder convs ~ N(0, sqrt(2 / (k * k * out_channels))); PyTorch defaults elsewhere
    (uniform within 1 / sqrt(fan_in)); BatchNorm gamma 1, beta 0, running stats (0, 1)."""
    import network
    torch.manual_seed(0)
    G = 2
    sd = network.Model_nefnet(theta_encoder_len=1, lead_num=G).state_dict()
    w = sd["W_encoder.layer1.0.conv1.weight"]
    assert float(w.std()) == pytest.approx(math.sqrt(2.0 / (7 * 7 * 128 * G)), rel=0.03) and abs(float(w.mean())) < 1e-4
    w = sd["W_encoder.conv1.weight"]
    assert float(w.std()) == pytest.approx(math.sqrt(2.0 / (15 * 15 * 128 * G)), rel=0.06)
    for name, fan_in in (("w_conv.0.conv1.weight", 128 * 3), ("z1_conv.0.residual_conv.weight", 64), ("z1_conv.0.residual_conv.bias", 64),
                         ("z2_conv2.1.weight", 64 * 2), ("z2_conv2.1.bias", 64 * 2), ("decoder.1.double_conv.0.weight", 256 * 3),
                         ("decoder.4.bias", 64 * 3), ("mlp1.weight", 12), ("mlp2.bias", 12)):
        b = 1.0 / math.sqrt(fan_in)
        assert float(sd[name].abs().max()) <= b and (sd[name].numel() < 100 or float(sd[name].abs().max()) > 0.9 * b), name
    assert bool((sd["decoder.3.double_conv.4.weight"] == 1).all()) and bool((sd["decoder.3.double_conv.4.bias"] == 0).all())
    assert bool((sd["decoder.1.double_conv.1.running_var"] == 1).all()) and int(sd["decoder.1.double_conv.1.num_batches_tracked"]) == 0


def test_sync_param_grads_writes_the_flat_buffer_back():
    """Autograd stores copies of the flat gradient views in p.grad; after the flat buffer changed (all-reduce) a stock
    optimiser needs them refreshed.  Host logic only: the flat buffers are built on the CPU here."""
    import network
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=1)
    with pytest.raises(RuntimeError):
        m.sync_param_grads()
    m._flatten(torch.device("cpu"))
    params = dict(m.named_parameters())
    assert all(p.data_ptr() >= m.flat_params.data_ptr() for p in params.values())   # parameters are views of the flat buffer
    params["mlp1.weight"].grad = torch.zeros(128, 12)
    m.flat_grads.fill_(0.5)
    m.sync_param_grads()
    assert bool((params["mlp1.weight"].grad == 0.5).all()) and params["mlp2.weight"].grad is None
    assert m.flat_grads.numel() % 4 == 0 and m.flat_grads.numel() >= 2702081


def test_data_parallel_replicas_fail_loudly():
    """solver.py:32-34 wraps the model in nn.DataParallel when it sees several GPUs; the B200 path is one process per GPU,
    and a replica must say so instead of running on buffers that belong to another device."""
    import network
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=1)
    replica = m._replicate_for_data_parallel()
    with pytest.raises(RuntimeError, match="DataParallel"):
        replica(torch.zeros(1, 1, 64), torch.zeros(1, 1, 2), torch.zeros(1, 2), torch.zeros(1, 7, 2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 64), torch.zeros(1, 1, 2), torch.zeros(1, 2), torch.zeros(1, 7, 2, dtype=torch.long))
