"""Fused optimiser step and data-parallel gradient exchange over the module's flat buffers.

Replaces ``torch.optim.SGD(params, lr, momentum=0.9).step()`` (reference: solver/optim_scheduler.py:10,
solver/solver.py:234-235) with one kernel over all parameters, and ``nn.DataParallel``'s gradient
reduction (solver/solver.py:32-34) with ONE NCCL all-reduce of the flat gradient buffer per step
(one process per GPU; the 1/world_size scale is folded into the optimiser kernel)."""
import torch

from . import _native as N


class FlatSGD:
    def __init__(self, model, lr=0.1, momentum=0.9):
        self.model = model
        self.lr = float(lr)
        self.momentum = float(momentum)
        self._mom = None

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None

    def step(self, world_size=1):
        m = self.model
        flat, grad = m.flat_params, m.flat_grads
        if flat is None:
            raise RuntimeError("FlatSGD.step() before the first forward/backward")
        if self._mom is None or self._mom.data_ptr() == 0 or self._mom.numel() != flat.numel():
            self._mom = torch.zeros_like(flat)
        lib = N.load()
        N.check(lib.nef_sgd_step(N.ptr(flat), N.ptr(grad), N.ptr(self._mom), flat.numel(), self.lr, self.momentum,
                                 1.0 / float(world_size), N.stream_ptr()), "nef_sgd_step")


def allreduce_gradients(model, group=None):
    """One all-reduce (sum) of the flat gradient buffer; the unused parameters have zero slots."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(model.flat_grads, op=dist.ReduceOp.SUM, group=group)
