"""GPU: where does the production path differ from the oracle under the B200 arithmetic model?  Compares the 'gen' latents
(z1 after z1_conv, z2 after the z2_conv2 chain) and the three outputs, dropout off, for the production path and its
precision switches.  python tests/probe_model_vs_device.py"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import torch
import network
from network import _native as N
from oracle import nefnet_oracle as O
from oracle.b200_precision import B200Precision

dev = torch.device("cuda:0")
lib = N.init(0)
B, G, L, seed = 4, 12, 5000, 41
P = O.make_params(G, seed)
inp = O.make_inputs(B, G, L, seed)
random.seed(seed)
c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)


def oracle(prec):
    with torch.no_grad():
        z1, z2 = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="gen", prec=prec)
        outs = O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(c1, c2),
                         prec=prec)
    return z1, z2, outs


def gpu():
    m = network.Model_nefnet(1, G)
    m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    m = m.float().to(dev).train()
    m.dropout_p = 0.0
    d = {k: v.to(dev) for k, v in inp.items()}
    with torch.no_grad():
        z1, z2 = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="gen")
        random.seed(seed)
        outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    return z1.cpu(), z2.cpu(), [o.cpu() for o in outs]


def cmp(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


refs = {"fp32": oracle(None), "model": oracle(B200Precision()), "model_no_f16": oracle(B200Precision(fwd_f16=False))}
for name, setup in (("production", lambda: None), ("fwd_f16=0", lambda: lib.nef_set_fwd_f16(0)),
                    ("conv_impl=0 (CUDA-core TF32)", lambda: lib.nef_set_conv_impl(0))):
    torch.cuda.synchronize()
    setup()
    g = gpu()
    torch.cuda.synchronize()
    lib.nef_set_fwd_f16(1); lib.nef_set_conv_impl(1)
    for rn, r in refs.items():
        z1e, z2e = cmp(g[0], r[0]), cmp(g[1], r[1].reshape(g[1].shape))
        oe = max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(g[2], r[2]))
        print("%-30s vs %-13s z1 relL2 %.2e max %.2e | z2 relL2 %.2e max %.2e | out max-rel %.2e"
              % (name, rn, z1e[0], z1e[1], z2e[0], z2e[1], oe), flush=True)
print("model vs fp32: z1 %.2e z2 %.2e out %.2e" % (cmp(refs["model"][0], refs["fp32"][0])[0], cmp(refs["model"][1], refs["fp32"][1])[0],
      max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(refs["model"][2], refs["fp32"][2]))))
