"""Nef-Net on B200: the nn.Module surface of the reference's ``network.model_nefnet.Model_nefnet``
(/root/reference/codes/network/model_nefnet.py:63-218) over the hand-written sm_100a kernels in
``csrc/`` (through the C ABI in include/nefnet_b200.h).

Same constructor, same ``forward(x, input_thetas, query_theta, rois, rest_theta=None, phase='train')``
and ``gen_ecg(z1, z2, query_theta, rois)`` signatures and return tuples, same ``state_dict`` keys,
shapes and default initialisation, same consumption of Python's ``random`` (two ``randint`` draws per
train/val/test forward, z1 first).  There is no CPU or eager-PyTorch path: calling the module on a
non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import random

import torch
import torch.nn as nn

from . import _native as N

_UNUSED = ("w_feature_extractor.0.weight", "w_feature_extractor.0.bias", "w_conv.0.residual_conv.weight",
           "w_conv.0.residual_conv.bias", "z2_conv2.0.residual_conv.weight", "z2_conv2.0.residual_conv.bias")


def _param_specs(G):
    """(name, shape, kind) in the reference's registration order (model_nefnet.py:67-107)."""
    specs = [("W_encoder.conv1.weight", (128 * G, 1, 15), "enc")]
    for i in range(3):
        specs.append((f"W_encoder.layer1.{i}.conv1.weight", (128 * G, 128, 7), "enc"))
        specs.append((f"W_encoder.layer1.{i}.conv2.weight", (128 * G, 128, 7), "enc"))
    specs += [("mlp1.weight", (128, 12), "w"), ("mlp1.bias", (128,), "b"), ("mlp2.weight", (256, 12), "w"),
              ("mlp2.bias", (256,), "b"), ("w_feature_extractor.0.weight", (128, 128, 3), "w"),
              ("w_feature_extractor.0.bias", (128,), "b")]

    def block(prefix, cin_g, groups):
        return [(prefix + ".conv1.weight", (128 * groups, cin_g, 3), "w"),
                (prefix + ".conv2.weight", (128 * groups, 128, 3), "w"),
                (prefix + ".residual_conv.weight", (128 * groups, cin_g, 1), "w"),
                (prefix + ".residual_conv.bias", (128 * groups,), "b")]

    specs += block("w_conv.0", 128, G) + block("z1_conv.0", 64, G) + block("z2_conv1.0", 64, G)
    specs += block("z2_conv2.0", 128, 7 * G)
    specs += [("z2_conv2.1.weight", (896 * G, 64, 2), "wt"), ("z2_conv2.1.bias", (448 * G,), "bt")]
    specs += block("z2_conv2.2", 64, 7 * G)
    for stage, cin, cout in (("decoder.1", 256, 128), ("decoder.3", 128, 64)):
        p = stage + ".double_conv."
        specs += [(p + "0.weight", (cout, cin, 3), "w"), (p + "0.bias", (cout,), "b"),
                  (p + "1.weight", (cout,), "one"), (p + "1.bias", (cout,), "zero"),
                  (p + "1.running_mean", (cout,), "buf0"), (p + "1.running_var", (cout,), "buf1"),
                  (p + "1.num_batches_tracked", (), "nbt"),
                  (p + "3.weight", (cout, cout, 3), "w"), (p + "3.bias", (cout,), "b"),
                  (p + "4.weight", (cout,), "one"), (p + "4.bias", (cout,), "zero"),
                  (p + "4.running_mean", (cout,), "buf0"), (p + "4.running_var", (cout,), "buf1"),
                  (p + "4.num_batches_tracked", (), "nbt")]
    specs += [("decoder.4.weight", (1, 64, 3), "w"), ("decoder.4.bias", (1,), "b")]
    return specs


class _Node(nn.Module):
    """Bare container so that parameters registered under dotted names give the reference's keys."""


class _NefFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, input_thetas, query_theta, rois, c1, c2, *live_params):
        outs = model._run_forward(x, input_thetas, query_theta, rois, None, N.PHASE_TRAIN, c1, c2, save=True)
        ctx.model = model
        ctx.token = model._fwd_token
        ctx.n_live = len(live_params)
        return outs

    @staticmethod
    def backward(ctx, dout, dout_p, dout_l):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise RuntimeError("Model_nefnet (B200): backward() after a newer forward(); only the latest training "
                               "forward is retained")
        # The parameter gradients are written straight into the flat gradient buffer and every p.grad is pointed at its view
        # of it (no per-step copies; a stock torch optimiser, FlatSGD and the data-parallel all-reduce all see one buffer),
        # so autograd itself receives no parameter gradients.
        with N.guard(model._flat):
            model._run_backward(dout, dout_p, dout_l)
        return (None,) * (7 + ctx.n_live)


class Model_nefnet(nn.Module):
    _variant = 1      # NefPlan variant (include/nefnet_b200.h: nef_plan_create_v)

    def _make_specs(self):
        specs = _param_specs(self.lead_num)
        return specs, [s[0] for s in specs]

    def __init__(self, theta_encoder_len=1, lead_num=1):
        super().__init__()
        if theta_encoder_len != 1:
            # theta_encoder.py:13-29 ignores encoder_len (one frequency); mlp1/mlp2 take 12 features only
            raise ValueError("theta_encoder_len must be 1 (the reference's ThetaEncoder emits 12 features)")
        self.theta_encoder_len = theta_encoder_len
        self.lead_num = int(lead_num)
        self.dropout_p = 0.2          # nn.Dropout(0.2) of every residual block; active in train() mode
        # _specs: state_dict entries in the reference's registration order; _names: the order of the C ABI's params / grads arrays
        self._specs, self._names = self._make_specs()
        self._plans = {}
        self._fwd_token = 0
        self._step = 0
        self._flat = None
        self._flat_grad = None
        # data parallel (one process per GPU): when a torch.distributed process group with more than one rank is initialised,
        # backward() itself averages the gradients over the ranks -- one all-reduce of the flat gradient buffer (or, with
        # NEF_DDP_OVERLAP=1, two buckets, the first overlapped with the encoder's backward) -- so the reference's unmodified
        # solver.py:232-235 (backward(); optim.step()) trains data
        # parallel under torchrun.  Set False to exchange gradients yourself (network.optim.allreduce_gradients).
        self.ddp_allreduce = True
        # Mis-tiled ROIs: the reference raises (torch.stack / torch.cat size mismatch, roi_pooling_1d.py:96-98) when the
        # truncated ROI lengths of a segment do not sum to L / 4.  "deferred" (default): a three-int flag written by one small
        # launch of every forward is read back without stalling the launch queue and the RuntimeError is raised at the
        # module's next entry point (the same step's backward() in a training loop, else the next forward / gen_ecg /
        # check_rois()); "sync": raised inside forward (one device synchronisation per call); "off": no check.
        self.roi_check = os.environ.get("NEF_ROI_CHECK", "deferred")
        self._roi_pending = None
        self._ddp_ready = False
        self._grads_valid = False
        self._build_parameters()

    # ------------------------------------------------------------------ parameters
    def _register(self, dotted, tensor, is_buffer):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        if is_buffer:
            node.register_buffer(parts[-1], tensor)
        else:
            node.register_parameter(parts[-1], nn.Parameter(tensor))

    def _build_parameters(self):
        """Default initialisation of the reference: N(0, sqrt(2 / (k*k*Cout))) for the resnet convs
        (resnet_1d.py:114-120), PyTorch's Conv/Linear defaults elsewhere, BN gamma 1 / beta 0."""
        for name, shape, kind in self._specs:
            if kind == "nbt":
                t = torch.zeros((), dtype=torch.long)
            elif kind in ("buf0", "zero"):
                t = torch.zeros(shape)
            elif kind in ("buf1", "one"):
                t = torch.ones(shape)
            elif kind == "enc":
                k = shape[2]
                t = torch.randn(shape) * math.sqrt(2.0 / (k * k * shape[0]))
            else:
                if kind in ("w", "b"):
                    wshape = shape if kind == "w" else dict((n, s) for n, s, _ in self._specs)[name[:-4] + "weight"]
                    fan_in = 1
                    for d in wshape[1:]:
                        fan_in *= d
                else:  # ConvTranspose1d: fan_in is computed from weight.size(1) * k
                    wshape = (896 * self.lead_num, 64, 2)
                    fan_in = wshape[1] * wshape[2]
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape) * 2.0 - 1.0) * bound
            self._register(name, t, is_buffer=kind in ("buf0", "buf1", "nbt"))

    def _tensors(self):
        """name -> tensor for every state_dict entry, in order."""
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return [sd[n] for n in self._names]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._flat = None  # parameter storage was replaced; re-flatten lazily on the next forward
        self._flat_grad = None
        for plan in self._plans.values():
            plan.close()
        self._plans = {}
        return out

    def _flatten(self, device):
        """Moves all float parameters into one contiguous fp32 buffer (16-byte aligned slots) and points
        the nn.Parameters at views of it; the gradient buffer has the same layout.  The optimiser step
        and the data-parallel all-reduce then run over flat memory."""
        live = [(n, p) for n, p in self.named_parameters()]
        offs, total = {}, 0
        for n, p in live:
            offs[n] = total
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        for n, p in live:
            view = flat[offs[n]:offs[n] + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat = flat
        self._flat_grad = torch.zeros_like(flat)
        self._offsets = offs
        self._grad_views = {n: self._flat_grad[offs[n]:offs[n] + p.numel()].view(p.shape) for n, p in live}

    @property
    def flat_params(self):
        return self._flat

    @property
    def flat_grads(self):
        return self._flat_grad

    @torch.no_grad()
    def sync_param_grads(self):
        """Writes the flat gradient buffer back into every ``p.grad`` that does not alias it.  Autograd stores COPIES
        of the views ``backward`` returns, so after ``allreduce_gradients`` changed the flat buffer a stock torch
        optimiser (which reads ``p.grad``) needs this; ``FlatSGD`` reads the flat buffer and does not."""
        if self._flat_grad is None:
            raise RuntimeError("Model_nefnet.sync_param_grads() before the first forward/backward")
        for n, p in self.named_parameters():
            gv = self._grad_views[n]
            if p.grad is not None and p.grad.data_ptr() != gv.data_ptr():
                p.grad.copy_(gv)

    def _ensure_ready(self, device):
        if getattr(self, "_is_replica", False):
            # solver.py:32-34 wraps the module in nn.DataParallel when several GPUs are visible; its per-call replicas
            # share this object's plans and flat buffers across devices.  The B200 path is one process per GPU.
            raise RuntimeError("Model_nefnet (B200) does not run under nn.DataParallel: launch one process per GPU "
                               "(CUDA_VISIBLE_DEVICES=<rank>) and use network.optim.allreduce_gradients")
        if device.type != "cuda":
            raise RuntimeError("Model_nefnet (B200) runs on an sm_100 CUDA device only; got tensors on %s. "
                               "There is no CPU path." % device)
        N.init(N.device_index(device))
        p0 = next(self.parameters())
        if p0.device != device:
            raise RuntimeError("Model_nefnet: parameters are on %s but inputs on %s" % (p0.device, device))
        ok = self._flat is not None and self._flat.device == device
        if ok:
            base = self._flat.data_ptr()
            end = base + self._flat.numel() * 4
            for n, p in self.named_parameters():
                if not (base <= p.data_ptr() < end) or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            for p in self.parameters():
                if p.dtype != torch.float32:
                    raise RuntimeError("Model_nefnet (B200): parameters must be float32")
            self._flatten(device)
            self._ddp_ready = False
        if not self._ddp_ready and self._ddp_world() > 1 and self.training:
            self._ddp_setup(device)

    def _param_ptr_array(self):
        arr = (C.c_void_p * len(self._names))()
        for i, t in enumerate(self._tensors()):
            arr[i] = t.data_ptr()
        return arr

    # ------------------------------------------------------------------ plans / workspace
    class _Plan:
        def __init__(self, B, G, L, V, device, variant=1):
            lib = N.load()
            h = C.c_void_p()
            N.check(lib.nef_plan_create_v(B, G, L, V, variant, C.byref(h)), "nef_plan_create_v")
            self.handle = h
            self.B = B
            self.bytes = lib.nef_plan_workspace_bytes(h)
            self.ws = torch.empty(self.bytes, dtype=torch.uint8, device=device)
            N.check(lib.nef_plan_bind(h, C.c_void_p(self.ws.data_ptr()), self.bytes, N.stream_ptr()), "nef_plan_bind")

        def close(self):
            if self.handle is not None:
                N.load().nef_plan_destroy(self.handle)
                self.handle = None
                self.ws = None

    def _plan(self, B, L, V, device):
        key = (B, L, V, device.index)
        plan = self._plans.get(key)
        if plan is None:
            for old in self._plans.values():  # one live workspace at a time (tens of GB at batch 256)
                old.close()
            self._plans = {}
            plan = Model_nefnet._Plan(B, self.lead_num, L, V, device, self._variant)
            self._plans[key] = plan
        return plan

    @staticmethod
    def _f32(t, device):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ execution
    def _run_forward(self, x, input_thetas, query_theta, rois, rest_theta, phase, c1, c2, save):
        device = x.device
        self._ensure_ready(device)
        lib = N.load()
        B, G, L = x.shape
        if G != self.lead_num:
            raise ValueError("expected %d leads, got %d" % (self.lead_num, G))
        if L % 4 != 0 or L < 16:
            raise ValueError("segment length must be a multiple of 4 and >= 16")
        V = int(rest_theta.shape[1]) if (rest_theta is not None and phase == N.PHASE_TEST) else 0
        plan = self._plan(B, L, V, device)
        x = self._f32(x, device)
        input_thetas = self._f32(input_thetas, device)
        query_theta = self._f32(query_theta, device)
        rois = rois.detach().to(device=device, dtype=torch.int64).contiguous()
        rest = self._f32(rest_theta, device) if V > 0 else None
        self._roi_launch_check(rois, B, L, device)
        a = N.NefForwardArgs()
        self._params_arr = self._param_ptr_array()
        a.params = C.cast(self._params_arr, C.POINTER(C.c_void_p))
        a.x, a.input_thetas, a.query_theta, a.rois = x.data_ptr(), input_thetas.data_ptr(), query_theta.data_ptr(), rois.data_ptr()
        a.rest_theta = rest.data_ptr() if rest is not None else None
        a.phase = phase
        a.bn_training = 1 if self.training else 0
        a.lead_choice_z1, a.lead_choice_z2 = int(c1), int(c2)
        a.drop_p = float(self.dropout_p) if self.training else 0.0
        a.save_for_backward = 1 if save else 0
        a.drop_seed = (((torch.initial_seed() * 6364136223846793005) + self._step) ^ (self._rank() * 0x9E3779B97F4A7C15)) \
            & 0x0FFFFFFFFFFFFFFF   # per step, and per rank: data-parallel replicas must not share dropout masks
        self._step += 1
        self._last_drop_seed = int(a.drop_seed)   # tests rebuild the keep masks from it (tests/_dropmask.py)
        if phase == N.PHASE_GEN:
            z1 = torch.empty((B, 128 * G, L // 4), dtype=torch.float32, device=device)
            z2 = torch.empty((B, 128 * G, 7, 32), dtype=torch.float32, device=device)
            a.out, a.out_p = z1.data_ptr(), z2.data_ptr()
            outs = (z1, z2)
        else:
            o = [torch.empty((B, 1, L), dtype=torch.float32, device=device) for _ in range(3)]
            a.out, a.out_p, a.out_l = o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()
            outs = tuple(o)
            if V > 0:
                ro = torch.empty((B, V, L), dtype=torch.float32, device=device)
                a.rest_out = ro.data_ptr()
                outs = outs + (ro,)
        N.check(lib.nef_forward(plan.handle, C.byref(a), N.stream_ptr()), "nef_forward")
        self._fwd_token += 1
        # keep the inputs alive until backward (the plan holds raw pointers to them)
        self._saved = (x, input_thetas, query_theta, rois, plan) if save else None
        return outs

    # ------------------------------------------------------------------ ROI tiling diagnosis
    def _roi_launch_check(self, rois, B, L, device):
        self.check_rois()                      # a pending verdict of an earlier call is due now
        if self.roi_check == "off":
            return
        if getattr(self, "_roi_flag_dev", None) is None or self._roi_flag_dev.device != device:
            self._roi_flag_dev = torch.zeros(3, dtype=torch.int32, device=device)
            self._roi_flag_host = torch.zeros(3, dtype=torch.int32).pin_memory()
        N.check(N.load().nef_roi_check(N.ptr(rois), B, L, N.ptr(self._roi_flag_dev), N.stream_ptr()), "nef_roi_check")
        self._roi_flag_host.copy_(self._roi_flag_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._roi_pending = (ev, L // 4)
        if self.roi_check == "sync":
            self.check_rois()

    def check_rois(self):
        """Raises the reference's RuntimeError if the ROIs of the last forward / gen_ecg did not tile their segments
        (waits for that call's first launch only; a no-op when nothing is pending)."""
        pend, self._roi_pending = self._roi_pending, None
        if pend is None:
            return
        ev, L4 = pend
        ev.synchronize()
        bad, first, total = (int(v) for v in self._roi_flag_host)
        if bad:
            self._saved = None
            raise RuntimeError("roi_pooling_reverse: the truncated ROI lengths of %d segment(s) do not tile L/4 = %d (first: segment "
                               "%d, lengths sum to %d); the reference fails here with a torch.stack / torch.cat size mismatch "
                               "(network/utils/roi_pooling_1d.py:96-98)" % (bad, L4, first, total))

    @staticmethod
    def _rank():
        import torch.distributed as dist
        return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0

    def _ddp_world(self):
        import torch.distributed as dist
        if self.ddp_allreduce and dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
        return 1

    def _ddp_setup(self, device):
        """Once per process group: every rank starts from rank 0's parameters and BatchNorm buffers (what DataParallel's
        replicate does every step, solver.py:32-34), and the side stream / event of the overlapped all-reduce exist."""
        import torch.distributed as dist
        with torch.no_grad():
            dist.broadcast(self._flat, src=0)
            for b in self.buffers():
                dist.broadcast(b, src=0)
        self._ddp_stream = torch.cuda.Stream(device=device)
        self._ddp_event = torch.cuda.Event()
        self._ddp_event.record()          # creates the underlying cudaEvent_t; nef_backward re-records it
        self._ddp_split = self._offsets["z1_conv.0.conv1.weight"]
        self._ddp_ready = True

    def _run_backward(self, dout, dout_p, dout_l):
        import torch.distributed as dist
        lib = N.load()
        self.check_rois()
        x, input_thetas, query_theta, rois, plan = self._saved
        device = x.device
        world = self._ddp_world()
        params = dict(self.named_parameters())
        # nef_backward accumulates (+=) like autograd: start from zero unless the caller kept the previous gradients (p.grad
        # still set).  With the built-in all-reduce that stays exact: what was accumulated before is already the same mean on
        # every rank, and averaging (mean + local) over the ranks gives mean + mean(local).
        accumulate = any(p.grad is not None and p.grad.data_ptr() == self._grad_views[n].data_ptr() for n, p in params.items())
        if not accumulate:
            self._flat_grad.zero_()
        names = self._names
        garr = (C.c_void_p * len(names))()
        for i, n in enumerate(names):
            gv = self._grad_views.get(n)
            garr[i] = gv.data_ptr() if (gv is not None and n not in _UNUSED) else None
        b = N.NefBackwardArgs()
        self._params_arr = self._param_ptr_array()
        b.params = C.cast(self._params_arr, C.POINTER(C.c_void_p))
        b.grads = C.cast(garr, C.POINTER(C.c_void_p))
        keep = []
        for field, g in (("dout", dout), ("dout_p", dout_p), ("dout_l", dout_l)):
            if g is not None:
                g = self._f32(g, device)
                keep.append(g)
                setattr(b, field, g.data_ptr())
        if world > 1:
            if not self._ddp_ready:
                self._ddp_setup(device)
            b.ev_late_params_done = self._ddp_event.cuda_event
        N.check(lib.nef_backward(plan.handle, C.byref(b), N.stream_ptr()), "nef_backward")
        self._saved = None
        if world > 1:
            # bucket 1 (z1_conv .. decoder, 2/3 of the bytes) is final long before the encoder's backward ends: reduce it on
            # a side stream behind the event nef_backward recorded; bucket 0 (stem, encoder, mlp, w_conv) follows on the
            # main stream.  ReduceOp.AVG: the mean over ranks, as DataParallel's gradient of the batch-mean loss.
            # Default: ONE all-reduce behind the whole backward; NEF_DDP_OVERLAP=1 selects the bucketed, overlapped form.
            # Measured on one box each: at 2 ranks the two forms tie (35.35 vs 35.47 ms; 35.30 vs 35.22 / 35.34 ms), at 8 the
            # overlapped one loses (35.01 vs 34.79 ms) -- the NCCL kernel's CTAs keep a few SMs from taking their CTA of every
            # persistent conv launch it runs beside (static tile walk), and at 8 ranks it runs longer.
            split = self._ddp_split
            overlap = os.environ.get("NEF_DDP_OVERLAP", "0") != "0"
            if not overlap:     # one all-reduce behind the whole backward
                dist.all_reduce(self._flat_grad, op=dist.ReduceOp.AVG)
                split = None
        if world > 1 and split is not None:
            self._ddp_stream.wait_event(self._ddp_event)
            with torch.cuda.stream(self._ddp_stream):
                w1 = dist.all_reduce(self._flat_grad[split:], op=dist.ReduceOp.AVG, async_op=True)
            w0 = dist.all_reduce(self._flat_grad[:split], op=dist.ReduceOp.AVG, async_op=True)
            w1.wait()
            w0.wait()
        for n, p in params.items():
            if n not in _UNUSED and p.requires_grad:
                p.grad = self._grad_views[n]
        self._grads_valid = True

    def export_activation(self, name):
        """Test hook (nef_plan_export): a named internal activation of the last forward as a (B, C, L) tensor."""
        lib = N.load()
        if not self._plans:
            raise RuntimeError("Model_nefnet.export_activation() before the first forward")
        plan = next(iter(self._plans.values()))
        c, l = C.c_int(), C.c_int()
        N.check(lib.nef_plan_tensor_info(plan.handle, name.encode(), C.byref(c), C.byref(l)), "nef_plan_tensor_info")
        B = plan.B
        shape = (B, c.value, l.value) if l.value > 0 else (c.value,)
        out = torch.empty(shape, dtype=torch.float32, device=self._flat.device)
        N.check(lib.nef_plan_export(plan.handle, name.encode(), N.ptr(out), N.stream_ptr()), "nef_plan_export")
        return out

    def poison_activation(self, name):
        """Test hook (nef_plan_poison): NaN-fill the fp32 storage of a named activation of the last forward."""
        lib = N.load()
        plan = next(iter(self._plans.values()))
        c, l = C.c_int(), C.c_int()
        N.check(lib.nef_plan_tensor_info(plan.handle, name.encode(), C.byref(c), C.byref(l)), "nef_plan_tensor_info")
        scratch = torch.empty((plan.B, c.value, l.value), dtype=torch.float32, device=self._flat.device)
        N.check(lib.nef_plan_poison(plan.handle, name.encode(), N.ptr(scratch), N.stream_ptr()), "nef_plan_poison")

    def _live_names(self):
        return [n for n, _ in self.named_parameters() if n not in _UNUSED]

    def _live_params(self):
        d = dict(self.named_parameters())
        return [d[n] for n in self._live_names()]

    # ------------------------------------------------------------------ public surface
    def forward(self, x, input_thetas, query_theta, rois, rest_theta=None, phase="train"):
        """model_nefnet.py:109-194.  x (B, lead_num, L) fp32, input_thetas (B, lead_num, 2), query_theta
        (B, 2), rois (B, 7, 2) int64, rest_theta (B, V, 2)."""
        if torch.is_tensor(x) and x.is_cuda:
            with N.guard(x):      # the native calls run on x's device and its current stream, whatever device is current
                return self._forward(x, input_thetas, query_theta, rois, rest_theta, phase)
        return self._forward(x, input_thetas, query_theta, rois, rest_theta, phase)

    def _forward(self, x, input_thetas, query_theta, rois, rest_theta=None, phase="train"):
        if phase == "gen":  # latents before roi reverse (:140-141); no random draws
            with torch.no_grad():
                return self._run_forward(x, input_thetas, query_theta, rois, None, N.PHASE_GEN, 0, 0, save=False)
        if phase not in ("train", "val", "test"):
            raise KeyError("please type correct phase")  # :194
        c1 = random.randint(0, self.lead_num - 1)  # :154
        c2 = random.randint(0, self.lead_num - 1)  # :156
        if phase == "train":
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                self._ensure_ready(x.device)
                return _NefFunction.apply(self, x, input_thetas, query_theta, rois, c1, c2, *self._live_params())
            return self._run_forward(x, input_thetas, query_theta, rois, None, N.PHASE_TRAIN, c1, c2, save=False)
        with torch.no_grad():  # the reference runs val/test under no_grad (solver.py:121)
            return self._run_forward(x, input_thetas, query_theta, rois, rest_theta, N.PHASE_TEST, c1, c2, save=False)

    def gen_ecg(self, z1, z2, query_theta, rois):
        """model_nefnet.py:196-218: decode V views from latents; flips the module to eval (:197)."""
        self.eval()
        device = z1.device
        if device.type == "cuda" and torch.cuda.current_device() != N.device_index(device):
            with N.guard(device):
                return self.gen_ecg(z1, z2, query_theta, rois)
        self._ensure_ready(device)
        lib = N.load()
        B = z1.shape[0]
        L = z1.shape[2] * 4
        V = int(query_theta.shape[1])
        plan = self._plan(B, L, V, device)
        z1 = self._f32(z1, device)
        z2 = self._f32(z2, device)
        q = self._f32(query_theta, device)
        rois = rois.detach().to(device=device, dtype=torch.int64).contiguous()
        self._roi_launch_check(rois, B, L, device)
        out = torch.empty((B, V, L), dtype=torch.float32, device=device)
        arr = self._param_ptr_array()
        N.check(lib.nef_gen_ecg(plan.handle, C.cast(arr, C.POINTER(C.c_void_p)), N.ptr(z1), N.ptr(z2), N.ptr(q),
                                N.ptr(rois), V, N.ptr(out), N.stream_ptr()), "nef_gen_ecg")
        return out
