"""Host restatement of the device's counter-based dropout keep-masks (csrc/nef_common.cuh: drop_bits / hash32,
csrc/nef_plan.cu: nef_forward's per-block seeds), so that the CPU oracle can be run with dropout ON using exactly the
masks the CUDA path applied (nefnet_oracle.forward(..., keeps=...)).  Test infrastructure only.

The hidden activation h of every residual block is dropped: element (row, 4-channel chunk c4, lane j) is kept iff the
j-th 16-bit field of (hash32(k), hash32(k ^ K3)), k = hash32(row ^ seedkey) + c4 * K1, is >= floor(p * 65536);
row = b * (L + 6) + 3 + l in the block's CBL4 row space, c4 = channel / 4 of the (groups * 128)-channel tensor,
seed = drop_seed * 16 + block index (hash32 = the two-multiply "lowbias32" finaliser).
"""
import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
HALO = 3
BLOCKS = ("W_encoder.layer1.0", "W_encoder.layer1.1", "W_encoder.layer1.2", "w_conv.0", "z1_conv.0", "z2_conv1.0",
          "z2_conv2.0", "z2_conv2.2")   # seed index = position (nef_plan.cu: seed + i)


def _hash32(x):
    x = x.astype(np.uint32) if isinstance(x, np.ndarray) else np.uint32(x)
    x = (x ^ (x >> np.uint32(16))) * np.uint32(0x7FEB352D)
    x = (x ^ (x >> np.uint32(15))) * np.uint32(0x846CA68B)
    return x ^ (x >> np.uint32(16))


def keep_mask(seed, B, C, L, p):
    """(B, C, L) bool keep-mask of a CBL4 tensor with C channels, B segments, L samples."""
    seed &= 0xFFFFFFFFFFFFFFFF
    with np.errstate(over="ignore"):
        sk = _hash32(np.uint32(seed & 0xFFFFFFFF) ^ _hash32(np.uint32((seed >> 32) & 0xFFFFFFFF) + np.uint32(0x9E3779B9)))
        rows = (np.arange(B, dtype=np.uint64)[:, None] * np.uint64(L + 2 * HALO) + np.uint64(HALO)
                + np.arange(L, dtype=np.uint64)[None, :]).reshape(-1)                       # (B*L,)
        rk = _hash32((rows & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ sk)                   # (B*L,)
        c4 = np.arange(C // 4, dtype=np.uint32)
        k = rk[None, :] + c4[:, None] * np.uint32(0x9E3779B1)                                 # (C/4, B*L)
        lo, hi = _hash32(k), _hash32(k ^ np.uint32(0x85EBCA6B))
    thr = np.uint32(int(np.float32(p) * np.float32(65536.0)))
    fields = [lo & np.uint32(0xFFFF), lo >> np.uint32(16), hi & np.uint32(0xFFFF), hi >> np.uint32(16)]
    lanes = np.stack([f >= thr for f in fields], axis=1)                                      # (C/4, 4, B*L)
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
