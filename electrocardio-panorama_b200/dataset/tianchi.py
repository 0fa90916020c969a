"""Whole-record preprocessing of ``EcgTianChiInterval.__getitem__`` (dataset/tianchi.py:84-111, 212-225) for a batch of
records on the device: lead derivation, heartbeat crop, min-max normalisation, padding to L, ROI table, lead selection.
The random draws of the reference (which heartbeat, which target lead, the angle jitter noise) stay with the caller and
come in as index / noise tensors, so a seeded run can be compared with the reference sample by sample."""
import math

import numpy as np
import torch

from network import _native as N

# tianchi.py:55-67: (theta, phi) of leads I, II, V1..V6, III, aVR, aVL, aVF
LEAD_THETA = np.array([[math.pi / 2, math.pi / 2], [math.pi * 5 / 6, math.pi / 2], [math.pi / 2, -math.pi / 18],
                       [math.pi / 2, math.pi / 18], [math.pi * (19 / 36), math.pi / 12], [math.pi * (11 / 20), math.pi / 6],
                       [math.pi * (16 / 30), math.pi / 3], [math.pi * (16 / 30), math.pi / 2], [math.pi * (5 / 6), -math.pi / 2],
                       [math.pi * (1 / 3), -math.pi / 2], [math.pi * (1 / 3), math.pi / 2], [math.pi * 1, math.pi / 2]])


def pack_records(records, device):
    """records: list of (8, T_b) arrays (any numeric dtype; the reference casts to float64, :85).  Returns the packed
    float64 device buffer with the per-record element offsets and lengths."""
    lens = [int(r.shape[1]) for r in records]
    offs = np.concatenate([[0], np.cumsum([8 * t for t in lens])[:-1]]).astype(np.int64)
    flat = np.concatenate([np.ascontiguousarray(r, dtype=np.float64).reshape(-1) for r in records])
    return (torch.from_numpy(flat).to(device), torch.from_numpy(offs).to(device),
            torch.tensor(lens, dtype=torch.int32, device=device))


def prepare_segments(raw, rec_off, rec_len, marks, L=512, select_index=None, target_index=None, want_ori=True):
    """raw / rec_off / rec_len from pack_records; marks (B, 7) int64 = p_on, p_off, r_on, r_off, t_on, t_off, end_point
    (:96-102).  Returns dict(data (B, G, L) | None, target_view (B, 1, L) | None, ori_data (B, 12, L) | None,
    rois (B, 7, 2) int64) -- the keys of the reference's ``meta`` that feed the model (:220-232)."""
    if not raw.is_cuda:
        raise RuntimeError("prepare_segments: CUDA tensors required (there is no CPU path)")
    dev = raw.device
    lib = N.init(N.device_index(dev))
    marks = marks.to(device=dev, dtype=torch.int64).contiguous()
    B = marks.shape[0]
    ori = torch.empty((B, 12, L), dtype=torch.float32, device=dev) if want_ori else None
    data = sel = None
    G = 0
    if select_index is not None:
        sel = torch.as_tensor(select_index).to(device=dev, dtype=torch.int32).contiguous()
        if sel.dim() == 1:
            sel = sel.unsqueeze(0).expand(B, -1).contiguous()
        G = sel.shape[1]
        data = torch.empty((B, G, L), dtype=torch.float32, device=dev)
    tgt = tidx = None
    if target_index is not None:
        tidx = torch.as_tensor(target_index).to(device=dev, dtype=torch.int32).contiguous()
        tgt = torch.empty((B, 1, L), dtype=torch.float32, device=dev)
    rois = torch.empty((B, 7, 2), dtype=torch.int64, device=dev)
    scratch = torch.empty(lib.nef_prepare_scratch_bytes(B) // 8, dtype=torch.float64, device=dev)
    with N.guard(dev):
        N.check(lib.nef_prepare_segments(N.ptr(raw), N.ptr(rec_off), N.ptr(rec_len), N.ptr(marks), B, L, N.ptr(sel), G,
                                         N.ptr(tidx), N.ptr(scratch), N.ptr(ori), N.ptr(data), N.ptr(tgt), N.ptr(rois),
                                         N.stream_ptr()),
                "nef_prepare_segments")
    return {"data": data, "target_view": tgt, "ori_data": ori, "rois": rois}


def angle_jitter(theta, jitter_factor, noise):
    """tianchi.py:77-82 with the normal draws supplied by the caller (noise ~ N(0, 1), same shape as theta)."""
    return theta + noise * (jitter_factor / 180.0 * math.pi)
