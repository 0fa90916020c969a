"""GPU: nef_prepare_segments / nef_psnr through the C ABI (dataset.tianchi, utils.mertic) against the reference's golden
vectors and against the numpy oracle on seeded inputs.  Byte / index outputs and the normalised samples are bit-exact
(double arithmetic, one cast); PSNR is a double reduction compared with numpy's float32 pairwise mean at 1e-5 dB."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as D

pytestmark = pytest.mark.gpu
PSNR_ATOL_DB = 1e-5


def test_prepare_segments_golden(golden_dir):
    from dataset import tianchi as T
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "data_tianchi.npz"))
    n = int(g["n"])
    recs = [g["record0"], g["record1"]]
    raw, off, lens = T.pack_records(recs, dev)
    which = [int(g["s%d/record" % i]) for i in range(n)]
    marks = torch.tensor(np.stack([g["s%d/marks" % i] for i in range(n)]))
    tidx = [int(g["s%d/target_index" % i]) for i in range(n)]
    out = T.prepare_segments(raw, off[which].contiguous(), lens[which].contiguous(), marks, L=512,
                             select_index=list(range(12)), target_index=tidx)
    for i in range(n):
        assert np.array_equal(out["ori_data"][i].cpu().numpy(), g["s%d/ori_data" % i]), i
        assert np.array_equal(out["data"][i].cpu().numpy(), g["s%d/data" % i]), i
        assert np.array_equal(out["target_view"][i, 0].cpu().numpy(), g["s%d/target_view" % i]), i
        assert np.array_equal(out["rois"][i].cpu().numpy(), g["s%d/rois" % i]), i


@pytest.mark.parametrize("L,G", [(512, 3), (5000, 12), (64, 1)])
def test_prepare_segments_vs_oracle(L, G):
    """Ragged record lengths, crops longer and shorter than L, arbitrary lead selections."""
    from dataset import tianchi as T
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(L + G)
    B = 9
    recs = [rng.integers(-400, 600, size=(8, int(t))).astype(np.int64) for t in rng.integers(L + 50, 3 * L + 200, size=B)]
    marks, sel, tgt = [], [], []
    for r in recs:
        T_ = r.shape[1]
        n = int(rng.integers(8, min(T_, 2 * L)))       # crop length: some < L (padding), some > L (truncation)
        p_on = int(rng.integers(0, T_ - n + 1))
        cuts = np.sort(rng.integers(p_on, p_on + n + 1, size=5))
        marks.append([p_on, *cuts.tolist(), p_on + n])
        sel.append(rng.permutation(12)[:G])
        tgt.append(int(rng.integers(0, 12)))
    raw, off, lens = T.pack_records(recs, dev)
    out = T.prepare_segments(raw, off, lens, torch.tensor(marks), L=L, select_index=np.stack(sel), target_index=tgt)
    for b in range(B):
        o = D.prepare_segment(recs[b], marks[b], L=L, select_index=sel[b], target_index=tgt[b])
        assert np.array_equal(out["ori_data"][b].cpu().numpy(), o["ori_data"]), b
        assert np.array_equal(out["data"][b].cpu().numpy(), o["data"]), b
        assert np.array_equal(out["target_view"][b, 0].cpu().numpy(), o["target_view"]), b
        assert np.array_equal(out["rois"][b].cpu().numpy(), o["rois"]), b
    # the prepared batch is what the model consumes: values in [0, 1], every segment touches both ends
    d = out["ori_data"]
    assert float(d.min()) == 0.0 and float(d.max()) == 1.0


def test_psnr_golden(golden_dir):
    from utils import mertic as M
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "data_psnr.npz"))
    pred, gt, rois = (torch.from_numpy(g[k]).to(dev) for k in ("pred", "gt", "rois"))
    assert abs(M.PSNR(pred, gt, rois) - float(g["psnr_rois"])) < PSNR_ATOL_DB
    assert abs(M.PSNR(pred, gt) - float(g["psnr_full"])) < PSNR_ATOL_DB
    acc = M.PsnrAccumulator(dev)
    acc.update(pred, gt, rois)
    rows = acc.rows(pred.shape[0] * pred.shape[1]).cpu().numpy()
    np.testing.assert_allclose(rows, D.psnr_rows(g["pred"], g["gt"], g["rois"]), rtol=0, atol=PSNR_ATOL_DB)
    assert rows[2 * 4 + 1] == 100.0


def test_psnr_accumulates_without_sync():
    """Two batches through one accumulator = the reference's mean over the concatenated rows (size-independent:
    a mean of means weighted by row count)."""
    from utils import mertic as M
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    chunks = []
    acc = M.PsnrAccumulator(dev)
    for B in (3, 7):
        gt = rng.random((B, 24, 5000), dtype=np.float32)
        pred = (gt + 0.02 * rng.standard_normal(gt.shape).astype(np.float32)).astype(np.float32)
        acc.update(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
        chunks.append(D.psnr_rows(pred, gt))
    assert abs(acc.value() - float(np.mean(np.concatenate(chunks)))) < PSNR_ATOL_DB
