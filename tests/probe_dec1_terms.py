"""Worst relative output error of the production (tcgen05 TF32) path against the fp32 oracle for 1, 2 and 3
split-precision terms in the decoder's first convolution (nef_set_dec1_terms).  Run on the GPU box:
    python tests/probe_dec1_terms.py"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import numpy as np, torch
import network
from network import _native as N
from oracle import nefnet_oracle as O
dev = torch.device("cuda:0")
lib = N.init(0)
cases = [(2, 12, 5000, 0), (4, 12, 5000, 1), (8, 3, 512, 2), (3, 12, 1000, 3), (16, 12, 2000, 4)]
for B, G, L, seed in cases:
    P = O.make_params(G, seed)
    inp = O.make_inputs(B, G, L, seed)
    random.seed(seed)
    c1, c2 = random.randint(0, G - 1), random.randint(0, G - 1)
    Po = {k: v.clone() for k, v in P.items()}
    stats = {k: v for k, v in Po.items() if "running_" in k or "num_batches" in k}
    with torch.no_grad():
        ref = O.forward(Po, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train",
                        lead_choice=(c1, c2), stats_out=stats)
    for terms in (1, 2, 3):
        N.check(lib.nef_set_dec1_terms(terms), "terms")
        m = network.Model_nefnet(1, G)
        m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
        m = m.float().to(dev).train()
        m.dropout_p = 0.0
        random.seed(seed)
        with torch.no_grad():
            outs = m(inp["x"].to(dev), inp["input_thetas"].to(dev), inp["query_theta"].to(dev), inp["rois"].to(dev), phase="train")
        errs = [float(((a.cpu() - b).abs() / b.abs()).max()) for a, b in zip(outs, ref)]
        print("B%d G%d L%d seed%d terms=%d  max-rel out/out_p/out_l = %.2e %.2e %.2e" % (B, G, L, seed, terms, *errs), flush=True)
lib.nef_set_dec1_terms(3)
