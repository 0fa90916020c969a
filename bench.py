#!/usr/bin/env python
"""Benchmark of the Nef-Net hot path on B200 (BASELINE.json: ECG segments/sec, B x 12 x 5000 train step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|4|5] [--batch B] [--length L]

A step is one Nef-Net training step on one synthetic batch: forward (three decoder passes, BatchNorm batch
statistics, dropout 0.2) + Standin-Learning loss + hand-written backward + (N > 1: one NCCL all-reduce of
the flat gradient buffer) + fused SGD-momentum update.  At N = 1 the workload is BASELINE.json configs[1]
(batch 256 x 12 leads x 5000 samples, fp32 storage / TF32 tensor-core multiply); N > 1 is weak scaling with
256 segments per GPU (configs[2]).  --config 4 times the same step at 64 x 12 x 20000 (configs[3]); --config 5 the
24-view panorama sweep in eval mode (configs[4]: encode once, 24 angular queries decoded, 64 segments per GPU, views/s).
One JSON line is printed by rank 0.

--impl reference times the reference algorithm's CPU restatement (oracle/, kind "port": the reference is
pure PyTorch and /root/reference does not exist on the GPU box) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "electrocardio-panorama_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "ECG segments/sec (Bx12x5000) Nef-Net train step"
UNIT = "segments/s"
METRIC_SWEEP = "panorama views/sec (24-view sweep, Bx12x5000) Nef-Net eval"


class Cfg:
    class SOLVER:
        reg_loss = "l1_loss"
        loss_using = [1, 2, 3]
        loss_factor = [0.5, 0.5, 1]


def workload_name(world, B, L, phase="train"):
    if phase == "sweep":
        return ("Nef-Net 24-view panorama sweep, eval mode (encode once, query view + %d angular queries decoded, BatchNorm folded), "
                "batch %d/GPU x 12 leads x %d samples" % (N_VIEWS, B, L))
    return ("Nef-Net train step (fwd x3 decoder passes + Standin L1 loss + bwd + %sSGD-momentum), batch %d/GPU x 12 leads x %d "
            "samples, dropout 0.2, BN batch stats" % ("NCCL grad all-reduce + " if world > 1 else "", B, L))


def measured_traffic(B, G, L):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        if (t["B"], t["G"], t["L"]) == (B, G, L):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


def algorithmic_bytes_per_segment(G, L, n_dec=3, live=True):
    """forward floats per segment (fp32 convention) with every fused block reading its inputs once and
    writing its output once.  Term by term:
        G L                 the input segment
        A                   stem output                      (A = 128 G L/4)
        15 A                three encoder blocks (2 A + 3 A each)
        5 A, 4 A            w_conv, z1_conv
        4 A                 z2_conv1 -- NOT in the live count: only its centre window reaches an output
        256 G + a, 15 a     roi_algin and the z2_conv2 chain (a = 128 G 7 16)
        2 a + A             roi_pooling_reverse
        A + A / G           lead means
        n_dec (9 d + L)     decoders (d = 256 L/4)
    'live' here removes only the 4 A of z2_conv1.  SURVEY's own live figure (241.8 MB at 12 x 5000) is a further 1.755 A
    lower (it also bills nothing for the reversed z2 tensor and the lead-mean pass, which latent_fwd fuses); it is reported
    beside this one as `survey_live` so that either convention can be read off the line."""
    A = 128 * G * (L // 4)
    a = 128 * G * 7 * 16
    d = 256 * (L // 4)
    fl = G * L + A + 15 * A + 5 * A + 4 * A + (0 if live else 4 * A) + 256 * G + a + 15 * a + 2 * a + A + A + A / G
    fl += n_dec * (9 * d + L)
    return 4.0 * fl


def survey_live_bytes_per_segment(G, L, n_dec=3):
    """SURVEY 8(d) 'live' variant: nominal minus 5.755 A (241.8 MB at 12 x 5000 x 3 decoders, 930.1 MB at 12 x 20000)."""
    A = 128 * G * (L // 4)
    return algorithmic_bytes_per_segment(G, L, n_dec, live=False) - 4.0 * 5.755 * A


def forward_report(ms_fwd, world, B, G, L, hbm_peak, n_dec=3):
    """Forward alone against the HBM roofline, with the byte conventions of SURVEY 8(d): 'nominal' (every reference
    block, 286.0 MB/segment at 12 x 5000), 'live' (z2_conv1 on its live window) and SURVEY's own live figure."""
    nominal = algorithmic_bytes_per_segment(G, L, n_dec, live=False) * B
    live = algorithmic_bytes_per_segment(G, L, n_dec, live=True) * B
    slive = survey_live_bytes_per_segment(G, L, n_dec) * B
    sec = ms_fwd / 1000.0
    return {"ms": ms_fwd, "segments_per_s": world * B / sec,
            "algorithmic_gb_nominal": nominal / 1e9, "hbm_frac_nominal": nominal / sec / 1e9 / hbm_peak,
            "algorithmic_gb_live": live / 1e9, "hbm_frac_live": live / sec / 1e9 / hbm_peak,
            "algorithmic_gb_survey_live": slive / 1e9, "hbm_frac_survey_live": slive / sec / 1e9 / hbm_peak,
            "what": "max over ranks of the forward (training mode: dropout, BN batch statistics, 3 decoder passes, "
                    "activations saved for backward) per GPU"}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm, self.smmax, self.reasons = [], [], set()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.sm.append(float(f[0]))
                self.smmax.append(float(f[1]))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.smmax), "reasons": sorted(self.reasons)}


def cpu_baseline(G, L, steps=2, warmup=1, B=None, device="cpu", reduce="min", budget_s=None, phase="train"):
    """The reference algorithm (oracle port) on the host: forward + loss + backward + SGD, dropout off is
    NOT used here -- the reference trains with dropout, so keep-masks are drawn on the host as it does.
    phase "sweep": the eval-mode 24-view forward (views/s)."""
    import torch
    from oracle import nefnet_oracle as O
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count() or 1)
    sweep = phase == "sweep"
    if B is None:
        B = 4 if sweep else (8 if L <= 5000 else 2)     # bounded samples: ~10-30 s of host work
    P = O.make_params(G, 0)
    inp = O.make_inputs(B, G, L, 0, V=N_VIEWS if sweep else 0)
    on_gpu = device != "cpu"   # opt-in context number (--ref-device cuda): the same port in PyTorch eager on the GPU, library kernels
    if on_gpu:
        P = {k: v.to(device) for k, v in P.items()}
        inp = {k: v.to(device) for k, v in inp.items()}
    mom = {}
    gen = torch.Generator().manual_seed(0)
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)

    def keeps():
        k = {}
        for name, ch, ln in ([(f"W_encoder.layer1.{i}", 128 * G, L // 4) for i in range(3)] +
                             [("w_conv.0", 128 * G, L // 4), ("z1_conv.0", 128 * G, L // 4),
                              ("z2_conv1.0", 128 * G, L // 4), ("z2_conv2.0", 896 * G, 16), ("z2_conv2.2", 896 * G, 32)]):
            k[name] = (torch.rand(B, ch, ln, generator=gen) >= 0.2).to(device)
        return k
    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        kk = None if sweep else keeps()
        sync()
        t0 = time.perf_counter()
        if sweep:
            with torch.no_grad():
                O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"],
                          phase="test", lead_choice=(0, 1 % G), bn_training=False)
        else:
            O.train_step(P, inp, lead_choice=(it % G, (it + 1) % G), momentum_buf=mom, keeps=kk)
        sync()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_begin > budget_s:
            break   # the run must end within minutes on any host: report the steps that were timed
    t = min(times) if reduce == "min" else sum(times) / len(times)
    kind = "port" if not on_gpu else "port on the GPU (PyTorch eager, cuDNN / library kernels; context only)"
    if sweep:
        return {"value": B * N_VIEWS / t, "unit": "views/s", "cores": torch.get_num_threads(), "kind": kind,
                "sample": "oracle eval forward with %d panorama views at B=%d x %d x %d fp32, %d warm-up + %s of %d"
                          % (N_VIEWS, B, G, L, warmup, "best" if reduce == "min" else "mean", len(times)),
                "timed_steps": len(times), "host_cpus": os.cpu_count(), "sample_batch": B}
    # the forward alone (training mode: dropout masks, BN batch statistics, 3 decoder passes), next to bench.py's `forward`
    ftimes = []
    for it in range(2):
        k = keeps()
        sync()
        t0 = time.perf_counter()
        with torch.no_grad():
            O.forward(P, inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train", lead_choice=(0, 1 % G),
                      keeps=k)
        sync()
        ftimes.append(time.perf_counter() - t0)
    return {"value": B / t, "forward_value": B / min(ftimes), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": kind,
            "sample": "oracle train step (fwd + Standin loss + bwd + SGD) at B=%d x %d x %d fp32, %d warm-up + %s of %d; "
                      "forward_value = the training-mode forward alone, best of 2"
                      % (B, G, L, warmup, "best" if reduce == "min" else "mean", len(times)),
            "timed_steps": len(times), "host_cpus": os.cpu_count(), "sample_batch": B}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all the host threads it can (torch is not imported yet)
    ncpu = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(ncpu)
    G, L = 12, args.length
    sweep = args.phase == "sweep"
    # exactly K timed steps after W warm-ups (mean), unless the host is so slow that the run would not end within minutes
    cb = cpu_baseline(G, L, steps=max(1, args.steps), warmup=max(0, args.warmup), B=args.ref_batch or None, device=args.ref_device,
                      reduce="mean", budget_s=240.0, phase=args.phase)
    sb = cb["sample_batch"]
    unit = "views/s" if sweep else UNIT
    per_step_units = sb * (N_VIEWS if sweep else 1)
    line = {"impl": "reference", "metric": METRIC_SWEEP if sweep else METRIC, "value": cb["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": cb["timed_steps"], "warmup": args.warmup, "ms_per_step": 1000.0 * per_step_units / cb["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus, args.batch, L, args.phase), "baseline_config": args.config,
                       "batch_per_gpu": args.batch,
                       "global_batch": args.batch * args.gpus, "leads": 12, "length": L,
                       "sample": "each step = the same workload on a bounded sample of %d segments on the host cores " % sb +
                                 "(reference algorithm, oracle port; the reference is pure PyTorch and cannot travel to the box)",
                       "scope": "ONE host: the value does not grow with --gpus (rank 0 alone runs it), so at N GPUs compare it "
                                "with the product arm's value / N (per-GPU ratio)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _event_ms(torch, fn, reps):
    for _ in range(2):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs)


def time_production_kernels(torch, dev, B, G, L, reps=6):
    """Average CUDA-event duration of the three heaviest launches of the train step, built through the public op layer
    EXACTLY as csrc/nef_plan.cu builds them for the encoder's k7 layers (128 G -> 128 G channels at L/4, operands > L2):
        wgrad_f16   nef_gconv_wgrad_f16 on fp16 copies of dY (loss-scaled) and X              (6 launches per step)
        dgrad       masked data gradient: fp16 operands, bit-plane mask, loss scale, fp16-only store (EPI 14368 | NOY)
        fwd         first convolution of a block: fp16 operands, ReLU + dropout 0.2 + fp16 copy + bit plane (EPI 5164 | NOY)
    All three are bound by the tensor pipe (2 x 128 x 7 MACs per operand byte pair), so they are reported against the
    measured bf16 tensor peak; their algorithmic bytes (read X and dY / write Y once, fp16) are given beside it."""
    from network import ops, _native as N
    C1, L4 = 128 * G, L // 4
    flops = 2.0 * B * L4 * C1 * 128 * 7
    abytes = 2.0 * C1 * L4 * B * 2
    out = {}

    def bits_plane():
        return torch.zeros((C1 // 32) * B * (L4 + 2 * N.HALO) + 2 * N.GUARD_ROWS, dtype=torch.int32, device=dev)

    # ---- fp16 weight gradient
    dy16, x16 = ops.H8(C1, B, L4, dev), ops.H8(C1, B, L4, dev)
    dy16.data.normal_()
    x16.data.normal_()
    dw = torch.zeros(C1, 128, 7, device=dev)
    inv = torch.tensor([1.0 / 64.0], device=dev)
    out["wgrad_f16"] = _event_ms(torch, lambda: ops.gconv_wgrad_f16(dy16, x16, dw, G, 128, 128, 7, out_scale=inv), reps)
    # ---- masked fp16 data gradient
    y, y16 = ops.Cbl4(C1, B, L4, dev), ops.H8(C1, B, L4, dev)
    w = torch.randn(C1, 128, 7, device=dev) * 0.03
    wd = ops.pack_conv_weight(w, G, dgrad=True, f16=True)
    d = ops.conv_desc(y, wd, y, G, 128, 128, 7, round_tf32=True)
    ops.use_f16_operand(d, dy16, wd, 128)
    bp = bits_plane()
    bp.fill_(0x55555555)
    sc = torch.tensor([64.0, 1.0 / 64.0], device=dev)
    d.y16 = y16.ptr
    d.mask_bits = bp.data_ptr() + 4 * N.GUARD_ROWS
    d.mask_mode, d.mask_scale = 1, 1.25
    d.mask, d.mask_cstride, d.mask_c4_off, d.mask_c4_gstride = y.ptr, y.rows, 0, 32
    d.acc_scale, d.y16_scale = sc.data_ptr() + 4, sc.data_ptr()
    d.y = None
    out["dgrad"] = _event_ms(torch, lambda: ops.gconv_fwd(d), reps)
    # ---- forward with dropout
    wf = ops.pack_conv_weight(w, G, f16=True)
    f = ops.conv_desc(y, wf, y, G, 128, 128, 7, relu=True, round_tf32=True)
    ops.use_f16_operand(f, x16, wf, 128)
    ob = bits_plane()
    f.y16 = y16.ptr
    f.out_bits = ob.data_ptr() + 4 * N.GUARD_ROWS
    f.drop_p, f.drop_seed = 0.2, 1234
    f.y = None
    out["fwd"] = _event_ms(torch, lambda: ops.gconv_fwd(f), reps)
    del dy16, x16, y, y16
    return out, flops, abytes


CONFIGS = {   # BASELINE.json configs[] -> (default segments per GPU, length, phase)
    2: (256, 5000, "train"),    # configs[1] (and configs[2] at N > 1: 256 per GPU)
    4: (64, 20000, "train"),    # configs[3]: long sequence
    5: (64, 5000, "sweep"),     # configs[4]: 512 segments over 8 GPUs, 24 views, eval
}
N_VIEWS = 24


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: 2 = 256 x 12 x 5000 train step (the headline; configs[2] at N > 1), "
                         "4 = 64 x 12 x 20000 train step, 5 = 24-view panorama sweep (eval, 64 segments per GPU, views/s)")
    ap.add_argument("--batch", type=int, default=0, help="segments per GPU (default: the configuration's)")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference only: 'cuda' times the oracle port in PyTorch eager on "
                    "the GPU (library kernels) as a context number; the reference arm proper is the default, 'cpu'")
    ap.add_argument("--ref-batch", type=int, default=0, help="--impl reference only: segments per step of the bounded sample")
    args = ap.parse_args()
    cB, cL, phase = CONFIGS[args.config]
    args.batch = args.batch or cB
    args.length = args.length or cL
    args.phase = phase
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator is created; stdout carries exactly one
        # JSON line, so the file descriptor points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            try:
                import ctypes
                ctypes.CDLL(None).fflush(None)   # the banner goes through C stdio
            except Exception:
                pass
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    import network
    from network import _native as N
    from network.optim import FlatSGD
    from dataset.synthetic import make_inputs  # product-side synthetic batches; oracle/ is only used by cpu_baseline()

    G, L, B = 12, args.length, args.batch
    sweep = phase == "sweep"
    V = N_VIEWS if sweep else 0
    lib = N.init(local)
    torch.manual_seed(0)
    random.seed(0)
    model = network.Model_nefnet(theta_encoder_len=1, lead_num=G).to(dev)
    model = model.eval() if sweep else model.train()
    opt = None if sweep else FlatSGD(model, lr=0.1, momentum=0.9)
    loss_fn = network.build_loss(type("C", (), {"MODEL": type("M", (), {"loss": "v1"})}))

    host = make_inputs(min(B, 16), G, L, seed=rank, V=V)
    reps = (B + host["x"].shape[0] - 1) // host["x"].shape[0]
    host = {k: v.repeat(*([reps] + [1] * (v.dim() - 1)))[:B].contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    in_keys = ("x", "input_thetas", "query_theta", "rois") + (("rest_theta",) if sweep else ("target",))
    h2d = sum(host[k].numel() * host[k].element_size() for k in in_keys)

    if sweep:
        def step(inp):     # encode once, decode the query view + 24 panorama views (model_nefnet.py:178-191)
            with torch.no_grad():
                return model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], rest_theta=inp["rest_theta"],
                             phase="test")[3]
    else:
        def step(inp):
            outs = model(inp["x"], inp["input_thetas"], inp["query_theta"], inp["rois"], phase="train")
            loss = loss_fn(outs[0], outs[1], outs[2], inp["target"], Cfg)[0]
            loss.backward()      # world > 1: backward() itself averages the gradients over the ranks (one NCCL all-reduce of
            opt.step()           # the flat buffer; NEF_DDP_OVERLAP=1: two buckets, the first beside the encoder's backward)
            opt.zero_grad()
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.nef_launch_count()
    ms = timed(lambda: step(resident), args.steps)
    launches = lib.nef_launch_count() - l0

    # forward only (the north-star's "fused encoder+decoder forward" target): the same training-mode forward with its
    # activations saved for backward (dropout, BN batch statistics, 3 decoder passes), no loss / backward / SGD
    ms_fwd = None
    if not sweep:
        def fwd_only():
            model(resident["x"], resident["input_thetas"], resident["query_theta"], resident["rois"], phase="train")
        fwd_only()
        ms_fwd = timed(fwd_only, args.steps) / args.steps

    # End to end through the public API with HOST inputs, as a training / serving loop feeds it: every step's inputs are
    # copied from pinned host memory (on a copy stream, one step ahead of the compute that consumes them -- what a
    # prefetching loader does) and every step's result (the loss; the B x 24 views of the sweep) is read back to the host,
    # one step late so that the read does not drain the launch queue.  All K copies in and all K reads complete inside the
    # timed region (the last read is waited for before the closing event).
    copy_stream = torch.cuda.Stream(device=dev)
    res_host = [torch.empty((B, V, L) if sweep else (), dtype=torch.float32).pin_memory() for _ in range(2)]
    state = {"next": None, "pending": None, "i": 0, "reads": 0}

    def prefetch():
        with torch.cuda.stream(copy_stream):
            inp = {k: host[k].to(dev, non_blocking=True) for k in in_keys}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return inp, ev

    def consume_pending():
        if state["pending"] is not None:
            ev, buf = state["pending"]
            ev.synchronize()
            _ = float(buf.flatten()[0])          # the value is on the host now
            state["reads"] += 1
            state["pending"] = None

    def e2e_step():
        if state["next"] is None:
            state["next"] = prefetch()
        inp, ev = state["next"]
        torch.cuda.current_stream().wait_event(ev)
        for t in inp.values():
            t.record_stream(torch.cuda.current_stream())
        state["next"] = prefetch()               # the next step's inputs travel while this step computes
        r = step(inp)
        buf = res_host[state["i"] & 1]
        buf.copy_(r.detach(), non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        consume_pending()                        # result of the previous step
        state["pending"] = (done, buf)
        state["i"] += 1

    def e2e_loop(k):
        for _ in range(k):
            e2e_step()
        consume_pending()
    e2e_loop(2)
    state["reads"] = 0
    ms_e2e = timed(lambda: e2e_loop(args.steps), 1)
    assert state["reads"] == args.steps, "every step's result must have been read back inside the timed region"
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # N > 1: what the step pays for the exchange.  (a) the gradient all-reduce of the flat buffer alone; (b) the same step
    # with the exchange switched off, per rank -- the ranks then run independently, so the spread between the fastest and the
    # slowest GPU of the box (power cap) shows, which a synchronised step hides behind its max.
    breakdown = None
    if world > 1 and not sweep:
        fg = model._flat_grad
        dist.all_reduce(fg, op=dist.ReduceOp.AVG)
        ms_ar = timed(lambda: dist.all_reduce(fg, op=dist.ReduceOp.AVG), args.steps) / args.steps
        model.ddp_allreduce = False
        step(resident)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        for _ in range(args.steps):
            step(resident)
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        model.ddp_allreduce = True
        per_rank = [round(float(t), 3) for t in allr]
        breakdown = {"allreduce_ms": ms_ar, "allreduce_bytes": int(fg.numel()) * 4,
                     "step_ms_without_exchange_per_rank": per_rank,
                     "step_ms_without_exchange_min": min(per_rank), "step_ms_without_exchange_max": max(per_rank),
                     "overlap": os.environ.get("NEF_DDP_OVERLAP", "0") != "0"}

    ms_step = ms / args.steps
    units = B * (V if sweep else 1)
    value = world * units / (ms_step / 1000.0)
    e2e_value = world * units / (ms_e2e / args.steps / 1000.0)
    d2h = B * V * L * 4 if sweep else 4
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops", 1590.0))          # burst figure: the kernels below are timed alone
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        kms, kflops, kbytes = time_production_kernels(torch, dev, B, G, L)
        n_dec = 1 + V if sweep else 3
        mult = 1.0 if sweep else 3.0    # a train step moves 3 x the forward bytes
        step_bytes = mult * algorithmic_bytes_per_segment(G, L, n_dec) * B
        step_bytes_survey = mult * survey_live_bytes_per_segment(G, L, n_dec) * B
        kern = {k: {"ms": v, "tflops": kflops / (v / 1000.0) / 1e12, "tensor_frac": kflops / (v / 1000.0) / 1e12 / tc_peak,
                    "algorithmic_gb_per_s": kbytes / (v / 1000.0) / 1e9, "hbm_frac": kbytes / (v / 1000.0) / 1e9 / hbm_peak}
                for k, v in kms.items()}
        dom = kern["wgrad_f16"]
        line = {"metric": METRIC_SWEEP if sweep else METRIC, "value": value, "unit": "views/s" if sweep else UNIT, "n_gpus": world,
                "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f16 operands (11-bit significand, = TF32) with fp32 accumulation in the 128-channel blocks, TF32 elsewhere; "
                         "fp32 master weights, statistics and reductions", "data": "synthetic",
                "config": {"workload": workload_name(world, B, L, phase), "baseline_config": args.config,
                           "batch_per_gpu": B, "global_batch": B * world, "leads": G, "length": L,
                           "conv_impl": "tcgen05 kind::f16 / kind::tf32" if lib.nef_get_conv_impl() == 1 else "cuda-core-fp32",
                           "l2": "inputs (>= 1 GB of activations per layer) far exceed the 126 MB L2; no flush needed"},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "views/s" if sweep else UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "pipelining": "inputs of step i+1 are copied (pinned host -> device, copy stream) while step i computes; "
                                      "the result of step i is read on the host during step i+1; K copies and K reads inside the "
                                      "timed region"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": dom["tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                             "frac": dom["tensor_frac"], "traffic": measured_traffic(B, G, L), "peak_source": peak_src,
                             "kernel": "wf16::wgrad_f16_kernel, the encoder's k7 weight gradient on fp16 operand copies "
                                       "(%d x %d x %d; 6 launches = the largest kernel share of the step)" % (B, 128 * G, L // 4),
                             "kernel_ms": kms["wgrad_f16"], "kernel_flops": kflops, "kernel_algorithmic_bytes": kbytes,
                             "kernel_algorithmic_gb_per_s": dom["algorithmic_gb_per_s"], "kernel_hbm_frac": dom["hbm_frac"],
                             "hbm_peak": hbm_peak,
                             "other_production_kernels": {k: v for k, v in kern.items() if k != "wgrad_f16"},
                             "step_algorithmic_gb": step_bytes / 1e9,
                             "step_hbm_frac": step_bytes / (ms_step / 1000.0) / 1e9 / hbm_peak,
                             "step_algorithmic_gb_survey_live": step_bytes_survey / 1e9,
                             "step_hbm_frac_survey_live": step_bytes_survey / (ms_step / 1000.0) / 1e9 / hbm_peak},
                }
        if ms_fwd is not None:
            line["forward"] = forward_report(ms_fwd, world, B, G, L, hbm_peak)
        if breakdown is not None:
            line["scaling_breakdown"] = breakdown
    if world > 1:
        dist.barrier()
    if rank == 0:
        del model
        torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(G, L, phase=phase)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
