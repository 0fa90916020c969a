"""CUDA-event timing of the two callers next to the hot path at the bench shapes, against their HBM traffic:
    python tools/bench_data.py
  prepare_segments: 256 records of 8 x 5000 float64 -> (256, 12, 5000) fp32 + rois   (reads raw twice: min/max pass + write pass)
  PSNR:             (256, 24, 5000) fp32 predictions vs targets"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "electrocardio-panorama_b200"))
import numpy as np, torch
from dataset import tianchi as T
from utils import mertic as M
dev = torch.device("cuda:0")
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6457.4


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B, L, G = 256, 5000, 12
rng = np.random.default_rng(0)
recs = [rng.integers(-400, 600, size=(8, L)).astype(np.int64) for _ in range(B)]
raw, off, lens = T.pack_records(recs, dev)
marks = torch.tensor([[0, 700, 900, 1100, 1300, 1700, L]] * B)
sel = list(range(12))
ms = timed(lambda: T.prepare_segments(raw, off, lens, marks, L=L, select_index=sel, target_index=[3] * B))
byt = 2 * B * 8 * L * 8 + 2 * B * 12 * L * 4 + B * L * 4   # raw read twice (min/max + write pass), ori + data + target written
print(json.dumps({"op": "prepare_segments", "B": B, "L": L, "ms": ms, "segments_per_s": B / ms * 1e3,
                  "algorithmic_gb": byt / 1e9, "gbs": byt / 1e9 / ms * 1e3, "hbm_frac": byt / 1e9 / ms * 1e3 / peak}))
V = 24
gt = torch.rand(B, V, L, device=dev)
pred = gt + 0.02 * torch.randn_like(gt)
acc = M.PsnrAccumulator(dev)
ms = timed(lambda: acc.update(pred, gt))
byt = 2 * B * V * L * 4
print(json.dumps({"op": "psnr", "B": B, "V": V, "L": L, "ms": ms, "views_per_s": B * V / ms * 1e3,
                  "algorithmic_gb": byt / 1e9, "gbs": byt / 1e9 / ms * 1e3, "hbm_frac": byt / 1e9 / ms * 1e3 / peak}))
