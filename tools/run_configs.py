"""Runs the BASELINE.json configurations 4 (L=20000, batch 64 train step) and 5 (24-view panorama sweep, eval) once on cuda:0 and
reports time and finiteness (the parity of these code paths is covered at small sizes by tests/test_gpu_parity.py)."""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
import network, bench
from network.optim import FlatSGD
from dataset import synthetic as O  # synthetic input generator
dev = torch.device("cuda:0")
G = 12
torch.manual_seed(0); random.seed(0)
model = network.Model_nefnet(1, G).to(dev).train()
opt = FlatSGD(model)
def rep(host, B):
    r = (B + host["x"].shape[0] - 1) // host["x"].shape[0]
    return {k: v.repeat(*([r] + [1] * (v.dim() - 1)))[:B].contiguous().to(dev) for k, v in host.items()}
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out
# config 4
B, L = 64, 20000
inp = rep(O.make_inputs(8, G, L, seed=0), B)
def step():
    outs = model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
    loss = network.losswrapper(outs[0], outs[1], outs[2], inp["target"], bench.Cfg)[0]
    loss.backward(); opt.step(); opt.zero_grad()
    return loss
dt, loss = timed(step)
print("config 4: train step B=%d L=%d  %.1f ms  %.1f segments/s  loss %.5f finite=%s" % (B, L, dt * 1e3, B / dt, float(loss), bool(torch.isfinite(loss))))
# config 5
model.eval()
B, L, V = 64, 5000, 24
inp = rep(O.make_inputs(8, G, L, seed=1, V=V), B)
def sweep():
    with torch.no_grad():
        return model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"], phase="test")
dt, outs = timed(sweep)
print("config 5: 24-view sweep B=%d (one GPU shard)  %.1f ms  %.0f views/s  rest_out %s finite=%s" % (
    B, dt * 1e3, B * V / dt, tuple(outs[3].shape), bool(torch.isfinite(outs[3]).all())))
