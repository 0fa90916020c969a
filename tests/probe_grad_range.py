"""CPU probe (not a test): dynamic range of the gradient dY at the output of every convolution of the path, from the fp32
oracle at B = 8 x 12 x 5000 with the Standin L1 loss -- the data behind the loss scale an fp16-operand backward needs
(DESIGN.md section 7).  With the L1 loss the upstream gradient is exactly +-factor / (B * L), so magnitudes scale as 1 / B.
    python tests/probe_grad_range.py > profiles/r01_grad_dynamic_range_oracle.txt"""
import sys, math, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch.nn.functional as F
from oracle import nefnet_oracle as O
torch.manual_seed(0)
G, L, B = 12, 5000, 8
P = O.make_params(G, 0)
inp = O.make_inputs(B, G, L, 0)
for n in O.live_param_names(G): P[n].requires_grad_(True)
stats = []
orig = F.conv1d
def conv1d(x, w, b=None, **kw):
    y = orig(x, w, b, **kw)
    tag = "conv%d k%d cout%d groups%d L%d" % (len(stats), w.shape[2], w.shape[0], kw.get("groups", 1), y.shape[-1])
    rec = {"tag": tag, "act_amax": float(y.detach().abs().max())}
    stats.append(rec)
    def hook(g, rec=rec):
        a = g.abs().flatten()
        nz = a[a > 0]
        rec["amax"] = float(a.max()); rec["nz_frac"] = float(nz.numel()) / a.numel()
        if nz.numel():
            s = nz[torch.randint(0, nz.numel(), (min(nz.numel(), 2_000_000),))]
            q = torch.quantile(s.double(), torch.tensor([0.001, 0.01, 0.5, 0.99], dtype=torch.double))
            rec["q"] = [float(v) for v in q]
    if y.requires_grad: y.register_hook(hook)
    return y
F.conv1d = conv1d
outs = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(0, 1))
loss = O.standin_loss(*outs, inp["target"])[0]
loss.backward()
print("B=%d: dY of every convolution output (|g| amax, nonzero fraction, quantiles 0.1%% 1%% 50%% 99%% of nonzero |g|)" % B)
for r in stats:
    if "amax" in r:
        print("%-44s act_amax %8.3g | dY amax %9.3g nz %.2f q %s" % (r["tag"], r["act_amax"], r["amax"], r["nz_frac"], " ".join("%.2e" % v for v in r.get("q", []))))
