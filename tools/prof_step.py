"""Profiling driver (run under ncu): a few train steps of the B200 path at the bench configuration.
    python tools/prof_step.py [--batch 256] [--steps 2] [--mode step|conv|wgrad]"""
import argparse, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--length", type=int, default=5000)
ap.add_argument("--mode", default="step")
a = ap.parse_args()
import network
from network import _native as N, ops
from network.optim import FlatSGD
from dataset import synthetic as O  # synthetic input generator
dev = torch.device("cuda:0")
lib = N.init(0)
G, L, B = 12, a.length, a.batch
if a.mode == "step":
    import bench
    torch.manual_seed(0); random.seed(0)
    model = network.Model_nefnet(1, G).to(dev).train()
    opt = FlatSGD(model)
    host = O.make_inputs(min(B, 16), G, L, seed=0)
    reps = (B + host["x"].shape[0] - 1) // host["x"].shape[0]
    inp = {k: v.repeat(*([reps] + [1] * (v.dim() - 1)))[:B].contiguous().to(dev) for k, v in host.items()}
    for _ in range(a.steps):
        outs = model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
        loss = network.losswrapper(outs[0], outs[1], outs[2], inp["target"], bench.Cfg)[0]
        loss.backward()
        opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    print("loss", float(loss))
else:
    C1, L4 = 128 * G, L // 4
    x, y = ops.Cbl4(C1, B, L4, dev), ops.Cbl4(C1, B, L4, dev)
    x.data.normal_(); y.data.normal_()
    w = torch.randn(C1, 128, 7, device=dev) * 0.03
    if a.mode == "conv":
        wpk = ops.pack_conv_weight(w, G)
        d = ops.conv_desc(x, wpk, y, G, 128, 128, 7, relu=True, round_tf32=True)
        for _ in range(a.steps):
            ops.gconv_fwd(d)
    else:
        dw = torch.zeros_like(w)
        for _ in range(a.steps):
            ops.gconv_wgrad(y, x, dw, G, 128, 128, 7)
    torch.cuda.synchronize()
