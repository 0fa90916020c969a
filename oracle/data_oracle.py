"""CPU restatement (numpy) of the two callers next to the hot path.  TEST INFRASTRUCTURE ONLY -- imported by tests/
(and never by the product package).  Pinned to outputs of the unmodified reference by oracle/make_golden_data.py
(tests/golden/data_*.npz, checked on every CPU run by tests/test_data_oracle.py)."""
import math

import numpy as np


def psnr_rows(pred, gt, rois=None):
    """utils/mertic.py:7-19: the per-(segment, view) values before the final np.mean (:21)."""
    out = []
    for i in range(pred.shape[0]):
        end_point = rois[i, -1, 0] if rois is not None else pred.shape[2]          # :11
        for j in range(pred.shape[1]):
            d = pred[i, j, :end_point] - gt[i, j, :end_point]                       # :13-14
            rmse = math.sqrt(np.mean(d ** 2))                                       # :15
            out.append(100 if rmse == 0 else 20 * np.log10(1.0 / rmse))            # :16-19
    return np.array(out, dtype=np.float64)


def psnr(pred, gt, rois=None):
    return float(np.mean(psnr_rows(pred, gt, rois)))                                # :21


def prepare_segment(record, marks, L=512, select_index=None, target_index=None):
    """dataset/tianchi.py:84-111, 212-225 for one record (8, T) and one heartbeat."""
    src = np.asarray(record).astype(np.float64)                                     # :85
    III = src[1:2] - src[0:1]                                                       # :88-93
    aVR = -0.5 * (src[0:1] + src[1:2])
    aVL = src[0:1] - 0.5 * src[1:2]
    aVF = src[1:2] - 0.5 * src[0:1]
    src = np.concatenate([src, III, aVR, aVL, aVF], axis=0)
    p_on, p_off, r_on, r_off, t_on, t_off, end_point = (int(v) for v in marks)
    rois = np.array([[p_on, p_off], [p_off, r_on], [r_on, r_off], [r_off, t_on], [t_on, t_off], [t_off, end_point],
                     [end_point, L + p_on]])                                        # :103-105 (512 -> L)
    rois -= p_on                                                                    # :106
    src = src[:, p_on:end_point]                                                    # :107
    mx, mn = np.max(src), np.min(src)                                               # :110-111
    src = (src - mn) / (mx - mn)

    def fit(a):                                                                     # :212-219
        n = a.shape[-1]
        if n < L:
            return np.pad(a, [(0, 0)] * (a.ndim - 1) + [(0, L - n)], mode="constant")
        return a[..., :L]
    out = {"ori_data": fit(src).astype(np.float32), "rois": rois.astype(np.int64)}
    if select_index is not None:
        out["data"] = fit(src[list(select_index)]).astype(np.float32)               # :209, :220
    if target_index is not None:
        out["target_view"] = fit(src[int(target_index)]).astype(np.float32)         # :203, :223
    return out
