"""Worker of tests/test_gpu_ddp.py (one process per GPU under torchrun): the module's built-in data-parallel exchange.

Every rank builds the model from a DIFFERENT seed (the first forward must broadcast rank 0's parameters), runs one training
step on its shard with the built-in all-reduce (backward() averages the gradients over the ranks: one all-reduce, or with
NEF_DDP_OVERLAP=1 two buckets, the first overlapped on a side stream), then re-runs EVERY shard locally with the exchange off and checks
    gradients after backward()  ==  mean over shards of the local gradients
to fp32 summation noise.  Dropout is off (per-rank dropout seeds differ by design); BatchNorm statistics are per replica, as
under nn.DataParallel (solver.py:32-34)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "electrocardio-panorama_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import network
    from dataset.synthetic import make_inputs
    G, L, per = 3, 512, 4
    torch.manual_seed(100 + rank)          # different initial parameters per rank on purpose
    m = network.Model_nefnet(1, G).to(dev).train()
    m.dropout_p = 0.0
    full = make_inputs(per * world, G, L, seed=5)
    gen = torch.Generator().manual_seed(9)
    ups = [torch.randn(per * world, 1, L, generator=gen) for _ in range(3)]

    def run(shard, ddp):
        sl = slice(shard * per, (shard + 1) * per)
        m.ddp_allreduce = ddp
        for p in m.parameters():
            p.grad = None
        random.seed(3)
        d = {k: v[sl].to(dev) for k, v in full.items()}
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
        torch.autograd.backward(outs, [u[sl].to(dev) for u in ups])
        return m.flat_grads.clone()

    g_ddp = run(rank, True)
    # parameters were broadcast from rank 0 at the first forward
    ref = m.flat_params.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, m.flat_params), "rank %d: parameters differ from rank 0's after the first forward" % rank
    bn0 = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k}
    local = [run(s, False) for s in range(world)]
    mean = sum(local) / world
    err = float((g_ddp - mean).norm() / mean.norm())
    amax = float((g_ddp - mean).abs().max() / mean.abs().max())
    print("rank %d: |ddp - mean(local)| rel-L2 %.3e, max %.3e" % (rank, err, amax), flush=True)
    assert err < 1e-5 and amax < 1e-4, (err, amax)
    # p.grad aliases the flat buffer (a stock optimiser sees the averaged gradients without copies)
    g_again = run(rank, True)
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert p.grad.data_ptr() == m._grad_views[n].data_ptr(), n
    assert float((g_again - mean).norm() / mean.norm()) < 1e-5
    # all ranks agree bit for bit after the exchange
    other = g_again.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(other, g_again), "rank %d: averaged gradients differ from rank 0's" % rank
    dist.barrier()
    if rank == 0:
        print("DDP_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
