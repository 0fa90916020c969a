"""Standin-Learning loss of the hot path, computed on the device (nef_loss_fwd / nef_loss_bwd through the C ABI).
Only the wrapper the Solver calls is part of the path; the reference's unused criteria are not mirrored."""
from .losses import losswrapper, pair_loss  # noqa: F401

__all__ = ["losswrapper", "pair_loss"]
