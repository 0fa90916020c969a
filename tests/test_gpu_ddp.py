"""GPU, needs 2 devices: N-GPU gradient equivalence on hardware of the module's built-in data-parallel exchange
(SURVEY 8e): torchrun, one process per GPU, NCCL.  Skipped on a one-GPU box (run it with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", ["0", "1"])   # one all-reduce behind the backward (default) / two buckets, the first overlapped
def test_builtin_allreduce_equals_mean_of_shard_gradients(overlap):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "tests", "ddp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, NEF_DDP_OVERLAP=overlap))
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "DDP_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_module_on_a_device_that_is_not_current():
    """a module on cuda:1 called while cuda:0 is the current device must launch on cuda:1 (device guard around the
    native calls) and give what the same module gives on cuda:0."""
    import random
    import network
    from oracle import nefnet_oracle as O
    cfg = type("Cfg", (), {"SOLVER": type("S", (), {"reg_loss": "l1_loss", "loss_using": [1, 2, 3], "loss_factor": [0.5, 0.5, 1]})})
    G, B, L = 3, 2, 256
    inp = O.make_inputs(B, G, L, 3)
    torch.manual_seed(0)
    ref = network.Model_nefnet(1, G)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    res = []
    torch.cuda.set_device(0)
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        m = network.Model_nefnet(1, G)
        m.load_state_dict(sd)
        m = m.to(dev).train()
        m.dropout_p = 0.0
        d = {k: v.to(dev) for k, v in inp.items()}
        random.seed(5)
        assert torch.cuda.current_device() == 0
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        loss = network.losswrapper(outs[0], outs[1], outs[2], d["target"], cfg)[0]
        loss.backward()
        from network.optim import FlatSGD
        opt = FlatSGD(m, lr=0.1, momentum=0.9)
        opt.step()
        torch.cuda.synchronize(dev)
        assert all(o.device == dev for o in outs)
        res.append(([o.detach().cpu() for o in outs], float(loss), {n: p.detach().cpu() for n, p in m.named_parameters()}))
    for a, b in zip(res[0][0], res[1][0]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    assert abs(res[0][1] - res[1][1]) < 1e-6
    for n in res[0][2]:
        torch.testing.assert_close(res[0][2][n], res[1][2][n], rtol=1e-4, atol=1e-6)
