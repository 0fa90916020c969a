"""Host restatement of the device's counter-based dropout keep-masks (csrc/nef_common.cuh: drop_bits / mix64,
csrc/nef_plan.cu: nef_forward's per-block seeds), so that the CPU oracle can be run with dropout ON using exactly the
masks the CUDA path applied (nefnet_oracle.forward(..., keeps=...)).  Test infrastructure only.

The hidden activation h of every residual block is dropped: element (row, 4-channel chunk c4, lane j) is kept iff the
j-th 16-bit field of mix64(seed ^ row * K1 ^ (c4 << 40) ^ c4 * K2) is >= floor(p * 65536); row = b * (L + 6) + 3 + l in the
block's CBL4 row space, c4 = channel / 4 of the (groups * 128)-channel tensor, seed = drop_seed * 16 + block index.
"""
import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
HALO = 3
BLOCKS = ("W_encoder.layer1.0", "W_encoder.layer1.1", "W_encoder.layer1.2", "w_conv.0", "z1_conv.0", "z2_conv1.0",
          "z2_conv2.0", "z2_conv2.2")   # seed index = position (nef_plan.cu: seed + i)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def keep_mask(seed, B, C, L, p):
    """(B, C, L) bool keep-mask of a CBL4 tensor with C channels, B segments, L samples."""
    with np.errstate(over="ignore"):
        rows = (np.arange(B, dtype=np.uint64)[:, None] * np.uint64(L + 2 * HALO) + np.uint64(HALO)
                + np.arange(L, dtype=np.uint64)[None, :]).reshape(-1)                       # (B*L,)
        c4 = np.arange(C // 4, dtype=np.uint64)
        key = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (rows[None, :] * np.uint64(0x9E3779B97F4A7C15))
               ^ (c4[:, None] << np.uint64(40)) ^ (c4[:, None] * np.uint64(0xD1B54A32D192ED03)))
        bits = _mix64(key)                                                                     # (C/4, B*L)
    thr = np.uint64(int(np.float32(p) * np.float32(65536.0)))
    lanes = np.stack([((bits >> np.uint64(16 * j)) & np.uint64(0xFFFF)) >= thr for j in range(4)], axis=1)  # (C/4, 4, B*L)
    return torch.from_numpy(lanes.reshape(C, B, L).transpose(1, 0, 2).copy())


def centre_window(L4):
    """csrc/nef_elem.cu: centre_window -- the columns of z2_conv1 the device evaluates."""
    y0 = int(np.floor((L4 - 1) * 0.5))
    lo, hi = max(y0 - 2, 0), min(y0 + 4, L4)
    return lo, hi - lo


def keeps_for(drop_seed, B, G, L, p=0.2):
    """name -> keep mask for nefnet_oracle.forward(keeps=...), as nef_forward generated them for `drop_seed`."""
    L4 = L // 4
    base = (drop_seed * 16) & 0xFFFFFFFFFFFFFFFF
    shapes = {"W_encoder.layer1.0": (128 * G, L4), "W_encoder.layer1.1": (128 * G, L4), "W_encoder.layer1.2": (128 * G, L4),
              "w_conv.0": (128 * G, L4), "z1_conv.0": (128 * G, L4), "z2_conv2.0": (896 * G, 16), "z2_conv2.2": (896 * G, 32)}
    out = {}
    for i, name in enumerate(BLOCKS):
        seed = (base + i) & 0xFFFFFFFFFFFFFFFF
        if name == "z2_conv1.0":   # evaluated on the centre window only; the other columns never reach an output
            w0, Lw = centre_window(L4)
            m = torch.ones(B, 128 * G, L4, dtype=torch.bool)
            m[:, :, w0:w0 + Lw] = keep_mask(seed, B, 128 * G, Lw, p)
            out[name] = m
        else:
            C, Lr = shapes[name]
            out[name] = keep_mask(seed, B, C, Lr, p)
    return out
