"""CPU: the numpy oracle of the two callers next to the hot path (oracle/data_oracle.py) against vectors produced by the
unmodified reference (oracle/make_golden_data.py)."""
import os

import numpy as np

from oracle import data_oracle as D


def test_prepare_segment_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "data_tianchi.npz"))
    n = int(g["n"])
    assert n >= 8
    for i in range(n):
        rec = g["record%d" % int(g["s%d/record" % i])]
        o = D.prepare_segment(rec, g["s%d/marks" % i], L=512, select_index=range(12), target_index=int(g["s%d/target_index" % i]))
        for k in ("ori_data", "data", "target_view", "rois"):   # bit-exact: same float64 arithmetic, one cast to fp32
            assert np.array_equal(o[k], g["s%d/%s" % (i, k)]), (i, k)


def test_psnr_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "data_psnr.npz"))
    assert D.psnr(g["pred"], g["gt"], g["rois"]) == float(g["psnr_rois"])
    assert D.psnr(g["pred"], g["gt"]) == float(g["psnr_full"])
    rows = D.psnr_rows(g["pred"], g["gt"], g["rois"])
    assert rows[2 * 4 + 1] == 100.0   # exact match row (mertic.py:16-17)
