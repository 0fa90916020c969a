"""In-situ kernel time vs step time (torch profiler / CUPTI): where are the idle gaps of a train step?"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import network, bench
from network.optim import FlatSGD
from dataset import synthetic as O  # synthetic input generator
dev = torch.device("cuda:0")
G, L, B = 12, 5000, 256
torch.manual_seed(0); random.seed(0)
model = network.Model_nefnet(1, G).to(dev).train()
opt = FlatSGD(model)
host = O.make_inputs(16, G, L, seed=0)
inp = {k: v.repeat(*([16] + [1] * (v.dim() - 1)))[:B].contiguous().to(dev) for k, v in host.items()}
def step():
    outs = model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
    loss = network.losswrapper(outs[0], outs[1], outs[2], inp["target"], bench.Cfg)[0]
    loss.backward(); opt.step(); opt.zero_grad()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("3 steps: span %.2f ms, kernel+memop busy %.2f ms, idle %.2f ms, %d device events" % ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(evs)))
gaps = []
for a, b in zip(evs, evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 15: gaps.append((g, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
print("gaps > 15 us: %d, total %.2f ms" % (len(gaps), sum(g[0] for g in gaps) / 1e3))
for g in gaps[:14]: print("  %7.0f us  after %-50s before %s" % g)
