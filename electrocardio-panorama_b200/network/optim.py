"""Fused optimiser step and data-parallel gradient exchange over the module's flat buffers.

Replaces ``torch.optim.SGD(params, lr, momentum=0.9).step()`` (reference: solver/optim_scheduler.py:10,
solver/solver.py:234-235) with one kernel over all parameters, and ``nn.DataParallel``'s gradient
reduction (solver/solver.py:32-34) with ONE NCCL all-reduce of the flat gradient buffer per step
(one process per GPU; the 1/world_size scale is folded into the optimiser kernel)."""
import torch

from . import _native as N


class FlatSGD(torch.optim.Optimizer):
    """SGD with momentum over the module's flat parameter / gradient buffers.  It is a ``torch.optim.Optimizer`` (one
    param group), so the reference's schedulers (``StepLR(optim, 50, gamma=0.1)`` / ``MultiStepLR``,
    optim_scheduler.py:13-18) and ``optim.zero_grad()`` / ``optim.step()`` (solver.py:232-235) drive it unchanged; the
    learning rate is read from ``param_groups[0]['lr']`` at every step.

    The step reads ``model.flat_grads``, the buffer ``backward`` writes; every ``p.grad`` is a VIEW of it, so editing
    gradients between ``backward`` and ``step`` (clipping, noise; the reference does neither) through ``p.grad`` or through
    ``flat_grads`` is the same thing.  As with ``torch.optim.SGD``: a ``step()`` without fresh gradients (none computed since
    ``zero_grad()``) changes nothing; parameters with ``requires_grad=False`` are not updated; ``step(closure)`` evaluates
    the closure first.  In a data-parallel run the module's ``backward`` has already averaged the gradients over the ranks
    (``Model_nefnet.ddp_allreduce``); ``world_size`` is only for callers that all-reduce SUMS themselves
    (``allreduce_gradients(model)`` + ``step(world_size=n)`` folds the 1 / n into the update kernel)."""

    def __init__(self, model, lr=0.1, momentum=0.9):
        self.model = model
        self._mom = None
        super().__init__(list(model.parameters()), dict(lr=float(lr), momentum=float(momentum)))

    @property
    def lr(self):
        return float(self.param_groups[0]["lr"])

    @property
    def momentum(self):
        return float(self.param_groups[0]["momentum"])

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None
        self.model._grads_valid = False   # the flat buffer is re-zeroed by the next backward

    def step(self, closure=None, *, world_size=1):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        m = self.model
        flat, grad = m.flat_params, m.flat_grads
        if flat is None:
            raise RuntimeError("FlatSGD.step() before the first forward/backward")
        if not getattr(m, "_grads_valid", True):
            return loss                   # no gradient since zero_grad(): torch.optim.SGD skips parameters without one
        with torch.no_grad():
            if self._mom is None or self._mom.data_ptr() == 0 or self._mom.numel() != flat.numel():
                self._mom = torch.zeros_like(flat)
            for n, p in m.named_parameters():   # frozen parameters: zero gradient and momentum slots leave them untouched
                if not p.requires_grad:
                    o = m._offsets[n]
                    grad[o:o + p.numel()].zero_()
                    self._mom[o:o + p.numel()].zero_()
            lib = N.load()
            with N.guard(flat):
                N.check(lib.nef_sgd_step(N.ptr(flat), N.ptr(grad), N.ptr(self._mom), flat.numel(), self.lr, self.momentum,
                                         1.0 / float(world_size), N.stream_ptr()), "nef_sgd_step")
        return loss

    # checkpoints (utils/checkpointer.py:28-31 saves optimizer.state_dict()): the momentum lives in one flat buffer
    def state_dict(self):
        return {"param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}], "state": {},
                "flat_momentum": None if self._mom is None else self._mom.detach().cpu()}

    def load_state_dict(self, sd):
        """Accepts its own format and ``torch.optim.SGD``'s (a checkpoint the reference's solver wrote with the stock optimiser,
        utils/checkpointer.py:28-31: ``state[i]['momentum_buffer']`` per parameter index)."""
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        m = self.model
        dev = m.flat_params.device if m.flat_params is not None else "cuda"
        mom = sd.get("flat_momentum")
        if mom is not None:
            self._mom = mom.to(dev)
            return
        self._mom = None
        state = sd.get("state") or {}
        if state:
            if m.flat_params is None:
                raise RuntimeError("FlatSGD.load_state_dict: move the model to its device and run one forward before loading a "
                                   "torch.optim.SGD state (the flat layout does not exist yet)")
            self._mom = torch.zeros_like(m.flat_params)
            for i, (n, p) in enumerate(m.named_parameters()):
                buf = state.get(i, state.get(str(i), {})).get("momentum_buffer")
                if buf is not None:
                    o = m._offsets[n]
                    self._mom[o:o + p.numel()].copy_(buf.reshape(-1))


def get_optimizer(cfg, model):
    """optim_scheduler.py:5-10 with the model instead of its parameter list: 'sgd' -> FlatSGD(momentum 0.9); 'adam' is
    not on the hot path and stays torch's."""
    if cfg.SOLVER.optim == "sgd":
        return FlatSGD(model, lr=cfg.SOLVER.lr, momentum=0.9)
    if cfg.SOLVER.optim == "adam":
        return torch.optim.Adam(model.parameters(), lr=cfg.SOLVER.lr)
    raise ValueError("get_optimizer: unknown optimiser %r" % (cfg.SOLVER.optim,))


def get_lr_scheduler(cfg, optim=None):
    """optim_scheduler.py:13-18, unchanged semantics."""
    from torch.optim.lr_scheduler import MultiStepLR, StepLR
    if cfg.SOLVER.scheduler == "steplr":
        return StepLR(optim, 50, gamma=0.1)
    if cfg.SOLVER.scheduler == "MultiStep":
        return MultiStepLR(optim, cfg.SOLVER.lr_step, gamma=0.1)
    return None


def allreduce_gradients(model, group=None, average=False):
    """One all-reduce (sum) of the flat gradient buffer; the unused parameters have zero slots.  Manual route for
    callers that switched the module's built-in exchange off (``model.ddp_allreduce = False``): ``FlatSGD`` folds
    1 / world_size into its kernel (``opt.step(world_size=n)``); for a stock torch optimiser pass ``average=True``."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(model.flat_grads, op=dist.ReduceOp.SUM, group=group)
        if average:
            model.flat_grads.mul_(1.0 / dist.get_world_size(group))
            model.sync_param_grads()
