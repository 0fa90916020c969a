"""Device-side PSNR, drop-in for ``utils/mertic.py:7-21`` (the module name keeps the reference's spelling).

The reference copies every synthesized view to the host and loops over (segment, view) rows in Python
(solver.py:179-240 calls it once per validation step).  Here the rows are reduced by ``nef_psnr`` on the device and
accumulated in device memory, so a validation epoch needs ONE device-to-host read (``PsnrAccumulator.value()``)."""
import ctypes as C

import torch

from network import _native as N


class PsnrAccumulator:
    """Running mean of the per-row PSNR values over any number of ``update`` calls; nothing synchronises until
    ``value()``."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.result = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._rows = None

    def reset(self):
        self.acc.zero_()
        self.result.zero_()

    def update(self, pred, gt, rois=None):
        """pred, gt (B, V, L) fp32 CUDA tensors; rois (B, 7, 2) int64 or None.  Returns the 0-dim running mean (device)."""
        if not pred.is_cuda:
            raise RuntimeError("PsnrAccumulator: CUDA tensors required (there is no CPU path)")
        lib = N.init(N.device_index(pred))
        pred = pred.detach().to(torch.float32).contiguous()
        gt = gt.detach().to(device=pred.device, dtype=torch.float32).contiguous()
        if pred.shape != gt.shape or pred.dim() != 3:
            raise ValueError("PSNR: pred and gt must both be (B, V, L)")
        B, V, L = pred.shape
        if rois is not None:
            rois = rois.detach().to(device=pred.device, dtype=torch.int64).contiguous()
        if self._rows is None or self._rows.numel() < B * V:
            self._rows = torch.empty(B * V, dtype=torch.float64, device=pred.device)
        with N.guard(pred):
            N.check(lib.nef_psnr(N.ptr(pred), N.ptr(gt), N.ptr(rois), B, V, L, N.ptr(self._rows), N.ptr(self.acc),
                                 N.ptr(self.result), N.stream_ptr()), "nef_psnr")
        return self.result[0]

    def rows(self, n):
        """The per-row values of the last update (device tensor view)."""
        return self._rows[:n]

    def value(self):
        """Mean PSNR so far as a Python float (the one synchronising read)."""
        return float(self.result[0])


def PSNR(pred, gt, rois=None, shave_border=0):
    """mertic.py:7-21 signature.  Accepts CUDA tensors (or numpy arrays, which are moved to the current device) and
    returns a Python float like the reference; use PsnrAccumulator in a loop to avoid the per-call synchronisation."""
    if not torch.is_tensor(pred):
        dev = torch.device("cuda", torch.cuda.current_device())
        pred = torch.as_tensor(pred, dtype=torch.float32).to(dev)
        gt = torch.as_tensor(gt, dtype=torch.float32).to(dev)
        if rois is not None:
            rois = torch.as_tensor(rois).to(dev)
    acc = PsnrAccumulator(pred.device)
    acc.update(pred, gt, rois)
    return acc.value()
