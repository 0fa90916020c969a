"""Data-parallel host logic on CPU (gloo, world_size 2): the single flat-gradient all-reduce that replaces
nn.DataParallel's gradient reduction (reference: solver/solver.py:32-34).  The device kernels are not involved;
per-shard gradients come from the CPU oracle, so this pins the exchange semantics SURVEY 8(e) states:
after the all-reduce and the 1/world scale, every rank holds the MEAN of the per-shard gradients (per-replica
BatchNorm statistics, as DataParallel computes them), laid out in the flat buffer the optimiser kernel walks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nefnet_oracle as O

G, L, B_GLOBAL, SEED = 2, 64, 4, 21


class _FlatHolder:
    """Stands in for Model_nefnet's flat-buffer surface (flat_grads) without touching CUDA."""

    def __init__(self, grads, names):
        self.offsets, total = {}, 0
        for n in names:
            self.offsets[n] = total
            total += (grads[n].numel() + 3) // 4 * 4   # same 16-byte slot rule as Model_nefnet._flatten
        self.flat_grads = torch.zeros(total)
        for n in names:
            self.flat_grads[self.offsets[n]:self.offsets[n] + grads[n].numel()] = grads[n].flatten()


def _shard_grads(rank, world):
    P = O.make_params(G, SEED)
    inp = O.make_inputs(B_GLOBAL, G, L, SEED)
    per = B_GLOBAL // world
    sl = slice(rank * per, (rank + 1) * per)
    for n in O.live_param_names(G):
        P[n].requires_grad_(True)
    outs = O.forward(P, inp["x"][sl], inp["input_thetas"][sl], inp["query_theta"][sl], inp["rois"][sl], phase="train",
                     lead_choice=(1, 0))
    O.standin_loss(*outs, inp["target"][sl])[0].backward()
    return {n: P[n].grad.detach().clone() for n in O.live_param_names(G)}


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "electrocardio-panorama_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from network.optim import allreduce_gradients
        torch.set_num_threads(2)
        names = O.live_param_names(G)
        mine = _shard_grads(rank, world)
        holder = _FlatHolder(mine, names)
        allreduce_gradients(holder)                      # ONE collective over the flat buffer
        mean = holder.flat_grads / world                 # the scale FlatSGD folds into nef_sgd_step (gscale)
        expect = [_shard_grads(r, world) for r in range(world)]
        worst = 0.0
        for n in names:
            e = sum(g[n] for g in expect) / world
            got = mean[holder.offsets[n]:holder.offsets[n] + e.numel()].view_as(e)
            worst = max(worst, float((got - e).abs().max() / (e.abs().max() + 1e-30)))
        # the stock-optimiser route: the real module's flat buffers (host logic only, CPU tensors), p.grad holding COPIES
        # of the flat views as autograd leaves them; average=True must write the mean back into every p.grad
        import network
        m = network.Model_nefnet(theta_encoder_len=1, lead_num=G)
        m._flatten(torch.device("cpu"))
        params = dict(m.named_parameters())
        for n in names:
            m._grad_views[n].copy_(mine[n])
            params[n].grad = m._grad_views[n].clone()
        allreduce_gradients(m, average=True)
        for n in names:
            e = sum(g[n] for g in expect) / world
            assert params[n].grad.data_ptr() != m._grad_views[n].data_ptr()
            worst = max(worst, float((params[n].grad - e).abs().max() / (e.abs().max() + 1e-30)))
        assert all(params[n].grad is None for n in O.UNUSED_PARAMS)
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_flat_gradient_allreduce_is_the_mean_of_shard_gradients():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-6, (r, ret[r])


def test_allreduce_is_a_noop_without_a_process_group():
    import sys
    from network.optim import allreduce_gradients
    h = _FlatHolder({"a": torch.ones(5)}, ["a"])
    allreduce_gradients(h)
    assert float(h.flat_grads.sum()) == 5.0


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """bench.py --impl reference launched the way the driver launches it for N > 1: rank 0 alone times the CPU arm and prints
    ONE JSON line, the other rank exits 0 without work (and without importing torch.distributed)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--length", "512", "--ref-batch", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1 and d["unit"] == "segments/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "NCCL" in d["config"]["workload"] and d["higher_is_better"] is True
