"""CPU: the address arithmetic of the experimental fp16 weight-gradient kernel (csrc/nef_wgrad_f16.cu), emulated with numpy.

The kernel feeds tcgen05.mma MN-major no-swizzle operands exactly as the bulk copies land them.  This test replays its staging
(chunk pitches, tap overhang), its descriptor start addresses (K step = 16 units, tap = 1 unit) and LBO / SBO choices through
the canonical layout CUTLASS documents for such operands -- element (mn, k) at
    start + (mn / 8) * SBO + (mn % 8) * 2 + (k / 8) * LBO + (k % 8) * 16      [bytes, fp16]
-- and checks that the accumulated tiles equal the weight gradient.  It pins the index math, not the hardware: whether the
part honours that layout (incl. 16-byte start shifts) is what tools/probe_umma16.cu and tests/test_gpu_experimental.py ask."""
import os
import re

import numpy as np
import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "electrocardio-panorama_b200", "csrc",
                   "nef_wgrad_f16.cu")


def _const(name):
    m = re.search(r"constexpr int %s = ([^;]+);" % name, open(SRC).read())
    assert m, name
    return m.group(1).split("//")[0].strip()


def _mn_major(smem, start, sbo, lbo, mn_n, k_n):
    h = smem.view(np.float16)
    mn, k = np.meshgrid(np.arange(mn_n), np.arange(k_n), indexing="ij")
    byte = start + (mn // 8) * sbo + (mn % 8) * 2 + (k // 8) * lbo + (k % 8) * 16
    return h[byte // 2].astype(np.float64)


@pytest.mark.parametrize("ntap,dcols", [(1, 64), (3, 128), (7, 64)])
def test_staging_and_descriptors_reproduce_the_weight_gradient(ntap, dcols):
    ROWS, YCH = int(_const("ROWS")), int(_const("YCH"))
    assert _const("YP") == "ROWS * 16" and _const("XP") == "(ROWS + 8) * 16" and ntap - 1 <= 8
    YP, XP = ROWS * 16, (ROWS + 8) * 16
    cout, rows_total, guard = 8 * YCH, 2 * ROWS, 8
    tap_off = -(ntap // 2)
    rng = np.random.default_rng(ntap)
    dy = rng.standard_normal((YCH, rows_total, 8)).astype(np.float16)                       # half8 [C/8][rows]
    x = rng.standard_normal((dcols // 8, rows_total + 2 * guard, 8)).astype(np.float16)     # with readable guard rows
    ref = np.zeros((ntap, cout, dcols))
    dyf = dy.transpose(1, 0, 2).reshape(rows_total, cout).astype(np.float64)
    xf = x.transpose(1, 0, 2).reshape(rows_total + 2 * guard, dcols).astype(np.float64)
    for t in range(ntap):
        ref[t] = dyf.T @ xf[guard + tap_off + t: guard + tap_off + t + rows_total]
    D = np.zeros_like(ref)
    for it in range(rows_total // ROWS):
        r0 = it * ROWS
        smem = np.zeros(YCH * YP + 16 * XP, dtype=np.uint8)
        h = smem.view(np.float16)
        for c in range(YCH):                        # producer: YP bytes of chunk c from row r0
            h[c * YP // 2: c * YP // 2 + ROWS * 8] = dy[c, r0:r0 + ROWS].reshape(-1)
        xrows = ROWS + ntap - 1                     # producer: (ROWS + taps - 1) rows of chunk c from row r0 + tap_off
        for c in range(dcols // 8):
            o = (YCH * YP + c * XP) // 2
            h[o:o + xrows * 8] = x[c, guard + r0 + tap_off: guard + r0 + tap_off + xrows].reshape(-1)
        for ks in range(ROWS // 16):                # issuer: yd = ylo + ks * 16 units, xd = xlo + ks * 16 + tap units
            A = _mn_major(smem, ks * 256, YP, 128, cout, 16)
            for tp in range(ntap):
                Bm = _mn_major(smem, YCH * YP + ks * 256 + tp * 16, XP, 128, dcols, 16)
                D[tp] += A @ Bm.T
    assert np.abs(D - ref).max() <= 1e-9 * np.abs(ref).max()
