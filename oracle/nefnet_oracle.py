"""CPU oracle for the Nef-Net hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain CPU PyTorch ops, of the algorithm the
reference implements in ``codes/network`` (Model_nefnet.forward / gen_ecg and the
Standin-Learning ``losswrapper``).  It exists so that the CUDA path can be checked for
parity; it is never imported by the product package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.

Pinning: the reference ships no golden vectors or tests of its own, so
this restatement is pinned against outputs of the reference itself, imported from
``/root/reference/codes`` inside the build container by ``oracle/make_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run (no reference needed at test time).

Every function cites the reference file:line it follows (paths relative to
``/root/reference/codes``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

N_ROI = 7  # rois per segment (dataset/tianchi.py:103-106)
ROI_SIZE = 16  # model_nefnet.py:136
SPATIAL_SCALE = 128.0 / 512.0  # model_nefnet.py:136,143


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
def param_shapes(G: int) -> "Dict[str, Tuple[int, ...]]":
    """state_dict key -> shape, in the reference's registration order
    (model_nefnet.py:67-107, encoder/encoder.py:16-26, encoder/resnet_1d.py:99-137)."""
    s: Dict[str, Tuple[int, ...]] = {}
    s["W_encoder.conv1.weight"] = (128 * G, 1, 15)
    for i in range(3):
        s[f"W_encoder.layer1.{i}.conv1.weight"] = (128 * G, 128, 7)
        s[f"W_encoder.layer1.{i}.conv2.weight"] = (128 * G, 128, 7)
    s["mlp1.weight"] = (128, 12)
    s["mlp1.bias"] = (128,)
    s["mlp2.weight"] = (256, 12)
    s["mlp2.bias"] = (256,)
    s["w_feature_extractor.0.weight"] = (128, 128, 3)
    s["w_feature_extractor.0.bias"] = (128,)

    def block(prefix, cin_g, groups):
        s[prefix + ".conv1.weight"] = (128 * groups, cin_g, 3)
        s[prefix + ".conv2.weight"] = (128 * groups, 128, 3)
        s[prefix + ".residual_conv.weight"] = (128 * groups, cin_g, 1)
        s[prefix + ".residual_conv.bias"] = (128 * groups,)

    block("w_conv.0", 128, G)
    block("z1_conv.0", 64, G)
    block("z2_conv1.0", 64, G)
    block("z2_conv2.0", 128, 7 * G)
    s["z2_conv2.1.weight"] = (896 * G, 64, 2)  # ConvTranspose1d layout (Cin, Cout/groups, k)
    s["z2_conv2.1.bias"] = (448 * G,)
    block("z2_conv2.2", 64, 7 * G)
    for stage, (cin, cout) in (("decoder.1", (256, 128)), ("decoder.3", (128, 64))):
        s[f"{stage}.double_conv.0.weight"] = (cout, cin, 3)
        s[f"{stage}.double_conv.0.bias"] = (cout,)
        for bn in ("1", "4"):
            if bn == "4":
                s[f"{stage}.double_conv.3.weight"] = (cout, cout, 3)
                s[f"{stage}.double_conv.3.bias"] = (cout,)
            s[f"{stage}.double_conv.{bn}.weight"] = (cout,)
            s[f"{stage}.double_conv.{bn}.bias"] = (cout,)
            s[f"{stage}.double_conv.{bn}.running_mean"] = (cout,)
            s[f"{stage}.double_conv.{bn}.running_var"] = (cout,)
            s[f"{stage}.double_conv.{bn}.num_batches_tracked"] = ()
    s["decoder.4.weight"] = (1, 64, 3)
    s["decoder.4.bias"] = (1,)
    return s


UNUSED_PARAMS = (  # never touched by forward: grads stay None in the reference
    "w_feature_extractor.0.weight",
    "w_feature_extractor.0.bias",
    "w_conv.0.residual_conv.weight",
    "w_conv.0.residual_conv.bias",
    "z2_conv2.0.residual_conv.weight",
    "z2_conv2.0.residual_conv.bias",
)


ZERO_GRAD_PARAMS = (  # conv biases feeding a train-mode BatchNorm: the batch mean removes them, so the
    # true gradient is exactly zero and what any implementation returns is rounding noise
    "decoder.1.double_conv.0.bias",
    "decoder.1.double_conv.3.bias",
    "decoder.3.double_conv.0.bias",
    "decoder.3.double_conv.3.bias",
)


def make_params(G: int, seed: int = 0, dtype=torch.float32) -> "Dict[str, torch.Tensor]":
    """Deterministic weights with the reference's key set, shapes and init *scales*
    (resnet_1d.py:114-120 normal init for the encoder convs, PyTorch default uniform bounds
    elsewhere, BN gamma/beta perturbed so that they matter).  The stream is this function's
    own (a CPU torch.Generator), so fixtures regenerate bit-identically without the reference."""
    gen = torch.Generator().manual_seed(1000003 * seed + G)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(G).items():
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            out[name] = 0.05 * torch.randn(shape, generator=gen)
        elif name.endswith("running_var"):
            out[name] = 0.5 + torch.rand(shape, generator=gen)
        elif name.startswith("W_encoder"):
            k = shape[2]
            out[name] = torch.randn(shape, generator=gen) * math.sqrt(2.0 / (k * k * shape[0]))
        elif ".double_conv.1." in name or ".double_conv.4." in name:
            if name.endswith("weight"):
                out[name] = 1.0 + 0.2 * (torch.rand(shape, generator=gen) - 0.5)
            else:
                out[name] = 0.2 * (torch.rand(shape, generator=gen) - 0.5)
        else:
            wname = name[: -len("bias")] + "weight" if name.endswith("bias") else name
            wshape = param_shapes(G)[wname]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            out[name] = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
        if out[name].is_floating_point():
            out[name] = out[name].to(dtype)
    return out


# ----------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------
TIANCHI_THETA = torch.tensor(  # dataset/tianchi.py:55-67
    [
        [math.pi / 2, math.pi / 2],
        [math.pi * 5 / 6, math.pi / 2],
        [math.pi / 2, -math.pi / 18],
        [math.pi / 2, math.pi / 18],
        [math.pi * (19 / 36), math.pi / 12],
        [math.pi * (11 / 20), math.pi / 6],
        [math.pi * (16 / 30), math.pi / 3],
        [math.pi * (16 / 30), math.pi / 2],
        [math.pi * (5 / 6), -math.pi / 2],
        [math.pi * (1 / 3), -math.pi / 2],
        [math.pi * (1 / 3), math.pi / 2],
        [math.pi * 1, math.pi / 2],
    ],
    dtype=torch.float32,
)


def make_inputs(B: int, G: int, L: int, seed: int = 0, V: int = 0, ragged_rois: bool = False):
    """x in [0,1] (tianchi.py:110-111), Tianchi lead angles + 2.5 deg jitter (tianchi.py:77-82,
    nef_net.yml:5), int64 rois tiling [0, L] on multiples of 4 (tianchi.py:103-106), U(0,1) target."""
    assert L % 4 == 0
    gen = torch.Generator().manual_seed(7919 * seed + 13 * B + G + L)
    x = 0.4 + 0.1 * torch.rand(B, G, L, generator=gen)
    n_spk = max(1, L // 400)
    for b in range(B):
        pos = torch.randint(0, L, (n_spk,), generator=gen)
        for p in pos.tolist():
            lo, hi = max(0, p - 6), min(L, p + 7)
            bump = 0.4 * torch.exp(-0.5 * ((torch.arange(lo, hi) - p) / 2.5) ** 2)
            x[b, :, lo:hi] += bump * (0.5 + 0.5 * torch.rand(G, 1, generator=gen))
    x.clamp_(0.0, 1.0)
    lead_idx = torch.arange(G) % 12
    jitter = 2.5 / 180.0 * math.pi
    input_thetas = TIANCHI_THETA[lead_idx][None].repeat(B, 1, 1) + jitter * torch.randn(B, G, 2, generator=gen)
    q_idx = torch.randint(0, 12, (B,), generator=gen)
    query_theta = TIANCHI_THETA[q_idx].clone()
    rois = torch.zeros(B, N_ROI, 2, dtype=torch.long)
    for b in range(B):
        if ragged_rois and b % 2 == 1:
            # zero-length rois and non multiple-of-4 cut points whose truncated lengths still sum to L/4
            cuts = torch.sort(torch.randint(1, L // 4, (N_ROI - 1,), generator=gen)).values * 4
            cuts[2] = cuts[1]  # an empty roi
            cuts = cuts + torch.tensor([1, 2, 2, 3, 0, 1])  # .long() truncation cases (roi_pooling_1d.py:85)
            cuts = torch.clamp(cuts, max=L)
        else:
            cuts = torch.sort(torch.randint(1, L // 4, (N_ROI - 1,), generator=gen)).values * 4
        edges = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.full((1,), L, dtype=torch.long)])
        rois[b, :, 0] = edges[:-1]
        rois[b, :, 1] = edges[1:]
    target = torch.rand(B, 1, L, generator=gen)
    out = dict(x=x, input_thetas=input_thetas, query_theta=query_theta, rois=rois, target=target)
    if V > 0:
        th = torch.tensor([math.pi / 6, math.pi / 3, math.pi / 2, 2 * math.pi / 3])
        ph = torch.tensor([-math.pi, -2 * math.pi / 3, -math.pi / 3, 0.0, math.pi / 3, 2 * math.pi / 3])
        grid = torch.stack(torch.meshgrid(th, ph, indexing="ij"), dim=-1).reshape(-1, 2)
        rest = grid[torch.arange(V) % grid.shape[0]]
        out["rest_theta"] = rest[None].repeat(B, 1, 1).contiguous()
        out["rest_view"] = torch.rand(B, V, L, generator=gen)
    return out


# ----------------------------------------------------------------------------------------
# pieces of the path
# ----------------------------------------------------------------------------------------
def theta_features(theta: torch.Tensor) -> torch.Tensor:
    """Angular-encoding features, utils/theta_encoder.py:13-29: for a in (theta, phi, theta+phi,
    theta-phi): [a, sin a, cos a] -> 12 values per view (omega = 1, a single frequency)."""
    a = torch.cat([theta, theta[..., 0:1] + theta[..., 1:2], theta[..., 0:1] - theta[..., 1:2]], dim=-1)
    feat = torch.stack([a, torch.sin(a), torch.cos(a)], dim=-1)
    return feat.reshape(*theta.shape[:-1], 12)


# ----------------------------------------------------------------------------------------
# optional arithmetic model ("prec"): None = plain fp32, the reference's arithmetic.  An object with
# conv / store / pre hooks (oracle/b200_precision.py) restates WHERE the B200 path rounds (operands of the
# tensor-core convolutions, stored activations, stored gradients), so that tests can separate precision
# effects (ReLU masks that flip within rounding distance of zero) from logic errors.
# ----------------------------------------------------------------------------------------
def _conv(prec, kind, x, w, b=None, **kw):
    """F.conv1d as the reference evaluates it (fp32), or as `prec` models the device's operand rounding.
    kind: 'fp16' (encoder k7 convs), 'dec' (decoder convolutions 2-4: fp16 operand copies in the fp16 decoder dataflow, else
    TF32), 'z2' (the z2_conv2 chain: the same under z2_f16), 'tf32' (every other grouped conv), 'fp32' (CUDA-core / split-precision)."""
    if prec is None:
        return F.conv1d(x, w, b, **kw)
    return prec.conv(kind, x, w, b, **kw)


def _st(prec, x):
    """a stored activation (the device rounds most of them to TF32 when it writes them)"""
    return x if prec is None else prec.store(x)


def _pre(prec, x):
    """a pre-activation whose GRADIENT the device stores rounded (identity in the forward direction)"""
    return x if prec is None else prec.pre(x)


def _relu(prec, name, x):
    """F.relu, or -- when `prec` carries activation patterns observed on the device (prec.pattern(name)) -- x * pattern:
    the same function wherever the pattern equals (x > 0), and a way to evaluate the backward pass on EXACTLY the device's
    ReLU / dropout pattern, so that gradient parity is not blurred by masks that flip within rounding distance of zero."""
    if prec is not None:
        m = prec.pattern(name)
        if m is not None:
            return x * m.to(x.dtype)
    return F.relu(x)


def _obs(prec, name, x):
    """lets `prec` record a named intermediate tensor (layer-by-layer comparison with the device's workspace)"""
    if prec is not None:
        prec.observe(name, x)
    return x


def _maybe_drop(h: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """nn.Dropout(0.2) of the residual blocks (resnet_1d.py:37,45; model_nefnet.py:46,52).  The
    oracle takes the keep-mask explicitly (None = dropout disabled) so that parity can be
    checked with dropout on, using the mask the CUDA path generated."""
    if keep is None:
        return h
    return h * keep.to(h.dtype) / (1.0 - p)


def residual_block(x, w1, w2, groups, res_w=None, res_b=None, keep=None, p=0.2, prec=None, kind="tf32", scale=None,
                   name=""):
    """conv -> ReLU -> Dropout -> conv -> (+ 1x1 residual conv iff channel counts differ) -> add
    -> ReLU.  resnet_1d.py:39-53 (k=7, identity residual) and model_nefnet.py:48-60 (k=3).
    scale: optional per-(segment, channel) factor applied to the block output (the angular scaling of
    model_nefnet.py:120-123, which the device fuses into the last encoder block's epilogue)."""
    pad = w1.shape[2] // 2
    h = _relu(prec, name + ".h", _pre(prec, _conv(prec, kind, x, w1, padding=pad, groups=groups)))
    if prec is not None and prec.pattern(name + ".h") is not None and keep is not None:
        h = h / (1.0 - p)   # the observed pattern of h already contains the dropout keep-mask
    else:
        h = _maybe_drop(h, keep, p)
    h = _obs(prec, name + ".h", _st(prec, h))
    y = _conv(prec, kind, h, w2, padding=pad, groups=groups)
    if y.shape[1] != x.shape[1]:
        r = _conv(prec, "z2" if kind == "z2" else "tf32", x, res_w, res_b, groups=groups)
    else:
        r = x
    y = _relu(prec, name + ".y", _pre(prec, y + r))
    if scale is not None:
        y = y * scale
    return _obs(prec, name + ".y", _st(prec, y))


def encoder(P, x, G, keeps=None, prec=None, out_scale=None):
    """encoder/encoder.py:28-40 with resnet_1d.py:102-105: grouped stem k15 s2 p7 -> ReLU ->
    MaxPool(3,2,1) -> three k7 residual blocks.  out_scale: see residual_block (last block)."""
    h = _conv(prec, "fp32", x, P["W_encoder.conv1.weight"], stride=2, padding=7, groups=G)
    sel = None if prec is None else prec.pattern("stem.argmax")
    if sel is None:
        h = F.max_pool1d(F.relu(h), kernel_size=3, stride=2, padding=1)
    else:
        # the device's own max-pool selections (nef_plan_export "stem.argmax": 0..2 = conv position 2j-1 / 2j / 2j+1, 3 = clipped
        # by the ReLU): near-ties between two conv positions route the gradient like a flipped ReLU mask would
        sel = sel.to(torch.int64)
        idx = (2 * torch.arange(sel.shape[-1]) - 1)[None, None, :] + sel.clamp(max=2)
        h = torch.gather(h, 2, idx.clamp(min=0)) * (sel < 3).to(h.dtype)
    h = _obs(prec, "stem", _st(prec, h))
    for i in range(3):
        keep = None if keeps is None else keeps.get(f"W_encoder.layer1.{i}")
        h = residual_block(h, P[f"W_encoder.layer1.{i}.conv1.weight"], P[f"W_encoder.layer1.{i}.conv2.weight"], G,
                           keep=keep, prec=prec, kind="fp16", scale=out_scale if i == 2 else None,
                           name=f"W_encoder.layer1.{i}")
    return h


def roi_align_center(z: torch.Tensor, rois: torch.Tensor, size: int = ROI_SIZE, scale: float = SPATIAL_SCALE):
    """What utils/roi_pooling_1d.py:38-69 (``roi_algin``) actually computes.  It feeds the roi
    coordinate as grid-x over a width-1 axis and 0 as grid-y over the length axis of
    ``F.grid_sample(bilinear, zeros, align_corners=False)``; hence every output sample is the
    bilinear read at the *centre* of the sequence, (len-1)/2, times the x-tent weight
    max(0, 1 - |gx|/2), gx being the projected roi linspace.  z: (B,C,L4) ->
    (B,C,7,size)."""
    B, C, Lz = z.shape
    r = rois.to(torch.float32) * scale  # :50-52
    r = r * (2.0 / Lz) - 1.0  # :53
    start, end = r[..., 0:1], r[..., 1:2]  # (B,7,1)
    i = torch.arange(size, dtype=torch.float32)
    step = (end - start) / (size - 1)
    # torch.linspace evaluates symmetrically from both ends (:58)
    gx = torch.where(i < size // 2, start + step * i, end - step * (size - 1 - i))  # (B,7,size)
    wx = torch.clamp(1.0 - gx.abs() * 0.5, min=0.0)
    iy = (Lz - 1) * 0.5
    y0 = int(math.floor(iy))
    wy1 = iy - y0
    centre = z[..., y0] * (1.0 - wy1)
    if wy1 > 0.0 and y0 + 1 < Lz:
        centre = centre + z[..., y0 + 1] * wy1
    return centre[:, :, None, None] * wx[:, None, :, :].to(z.dtype)


def roi_reverse(z: torch.Tensor, rois: torch.Tensor, scale: float = SPATIAL_SCALE, out_len: Optional[int] = None):
    """utils/roi_pooling_1d.py:72-99: per (b, roi j) linear resample (align_corners=False) of the
    S samples to ``long(r1*scale) - long(r0*scale)`` points, concatenated over j.
    z: (B,C,7,S) -> (B,C,sum_len).  The reference needs every b to give the same total."""
    B, C, R, S = z.shape
    rl = (rois.to(torch.float32) * scale).long()  # :83-85 (truncation)
    outs = []
    for b in range(B):
        parts = []
        for j in range(R):
            n = int(rl[b, j, 1] - rl[b, j, 0])
            if n != 0:
                parts.append(F.interpolate(z[b:b + 1, :, j, :], n, mode="linear", align_corners=False))
        outs.append(torch.cat(parts, dim=-1))
    res = torch.cat(outs, dim=0)
    if out_len is not None:
        assert res.shape[-1] == out_len
    return res


def _bn(x, P, prefix, training, stats_out):
    """BatchNorm1d (eps 1e-5, momentum 0.1) of DoubleConv, model_nefnet.py:19,22."""
    rm, rv = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    if stats_out is not None:
        rm, rv = stats_out[prefix + ".running_mean"], stats_out[prefix + ".running_var"]
    y = F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], training, 0.1, 1e-5)
    if training and stats_out is not None:
        stats_out[prefix + ".num_batches_tracked"] += 1
    return y


def decoder(P, lat, training, stats_out=None, prec=None, name="dec"):
    """model_nefnet.py:101-107: Upsample x2 -> DoubleConv(256,128) -> Upsample x2 ->
    DoubleConv(128,64) -> Conv1d(64,1,3); then sigmoid(x/3) (:168)."""
    h = F.interpolate(lat, scale_factor=2, mode="linear", align_corners=False)
    first = True
    for stage in ("decoder.1", "decoder.3"):
        for conv, bn in (("0", "1"), ("3", "4")):
            pre = f"{stage}.double_conv."
            # the device evaluates the first convolution split-precision (x_hi w_hi + x_lo w_hi + x_hi w_lo): fp32-like
            h = _pre(prec, _conv(prec, "fp32" if first else "dec", h, P[pre + conv + ".weight"], P[pre + conv + ".bias"],
                                 padding=1))
            _obs(prec, f"{name}.{stage}.{conv}", h)
            first = False
            h = _relu(prec, f"{name}.{stage}.{bn}", _bn(h, P, pre + bn, training, stats_out))
            if conv == "0":          # stored (rounded) as the next convolution's operand; after the second conv of a stage the
                h = _st(prec, h)     # device stores the UPSAMPLED tensor (decoder.1) or feeds the fp32 output kernel (decoder.3)
        if stage == "decoder.1":
            h = _st(prec, F.interpolate(h, scale_factor=2, mode="linear", align_corners=False))
    h = _conv(prec, "fp32", h, P["decoder.4.weight"], P["decoder.4.bias"], padding=1)
    return torch.sigmoid(h / 3)


def latents(P, x, input_thetas, rois, G, keeps=None, stop_before_reverse=False, prec=None):
    """model_nefnet.py:117-143: encoder, input-view angular scaling, w_conv, z1/z2 split,
    z1_conv, z2_conv1, roi_algin, z2_conv2 chain, roi_pooling_reverse."""
    kp = (lambda n: None) if keeps is None else keeps.get
    B = x.shape[0]
    enc = F.linear(theta_features(input_thetas), P["mlp1.weight"], P["mlp1.bias"])  # (B,G,128) :118,121
    if prec is None:
        w = encoder(P, x, G, keeps)  # (B,128G,L4)
        L4 = w.shape[-1]
        w = (w.view(B, G, 128, L4) * enc[..., None]).view(B, 128 * G, L4)  # :122-123
    else:  # the device applies the scale in the last encoder block's epilogue, before the stored value is rounded
        w = encoder(P, x, G, keeps, prec=prec, out_scale=enc.reshape(B, 128 * G, 1))
        L4 = w.shape[-1]
    w = residual_block(w, P["w_conv.0.conv1.weight"], P["w_conv.0.conv2.weight"], G, keep=kp("w_conv.0"), prec=prec,
                       name="w_conv.0")  # :124
    w = w.view(B, G, 2, 64, L4)  # :125-131 each lead's 128 ch -> (z1 half, z2 half)
    z1 = w[:, :, 0].reshape(B, 64 * G, L4)
    z2 = w[:, :, 1].reshape(B, 64 * G, L4)
    z1 = residual_block(z1, P["z1_conv.0.conv1.weight"], P["z1_conv.0.conv2.weight"], G,
                        P["z1_conv.0.residual_conv.weight"], P["z1_conv.0.residual_conv.bias"], keep=kp("z1_conv.0"),
                        prec=prec, name="z1_conv.0")
    z2 = residual_block(z2, P["z2_conv1.0.conv1.weight"], P["z2_conv1.0.conv2.weight"], G,
                        P["z2_conv1.0.residual_conv.weight"], P["z2_conv1.0.residual_conv.bias"],
                        keep=kp("z2_conv1.0"), prec=prec, name="z2_conv1.0")
    z2 = _obs(prec, "roi_align", _st(prec, roi_align_center(z2, rois)))  # (B,128G,7,16) :136
    z2 = z2.reshape(B, 128 * G * N_ROI, ROI_SIZE)  # :137
    z2 = residual_block(z2, P["z2_conv2.0.conv1.weight"], P["z2_conv2.0.conv2.weight"], 7 * G, keep=kp("z2_conv2.0"),
                        prec=prec, kind="z2", name="z2_conv2.0")
    if prec is None:
        z2 = F.conv_transpose1d(z2, P["z2_conv2.1.weight"], P["z2_conv2.1.bias"], stride=2, groups=7 * G)
    else:
        z2 = _st(prec, _pre(prec, prec.conv_transpose(z2, P["z2_conv2.1.weight"], P["z2_conv2.1.bias"], stride=2,
                                                      groups=7 * G)))
    z2 = residual_block(z2, P["z2_conv2.2.conv1.weight"], P["z2_conv2.2.conv2.weight"], 7 * G,
                        P["z2_conv2.2.residual_conv.weight"], P["z2_conv2.2.residual_conv.bias"],
                        keep=kp("z2_conv2.2"), prec=prec, kind="z2", name="z2_conv2.2")
    z2 = z2.view(B, 128 * G, N_ROI, 2 * ROI_SIZE)  # :138
    if stop_before_reverse:
        return z1, z2
    z2 = roi_reverse(z2, rois, out_len=L4)  # :143
    return z1, z2


def forward(P, x, input_thetas, query_theta, rois, rest_theta=None, phase="train", lead_choice=(0, 0),
            bn_training=True, keeps=None, stats_out=None, prec=None):
    """Model_nefnet.forward, model_nefnet.py:109-194.  ``lead_choice`` are the two
    ``random.randint(0, G-1)`` draws (:154,156; z1 first).  ``stats_out``: dict of BN buffers
    updated in place (three sequential updates per train forward, in call order out, p, l)."""
    G = x.shape[1]
    B = x.shape[0]
    if phase == "gen":
        return latents(P, x, input_thetas, rois, G, keeps, stop_before_reverse=True, prec=prec)  # :140-141
    z1, z2 = latents(P, x, input_thetas, rois, G, keeps, prec=prec)
    L4 = z1.shape[-1]
    z1g, z2g = z1.view(B, G, 128, L4), z2.view(B, G, 128, L4)
    z1_mean, z2_mean = z1g.mean(dim=1), z2g.mean(dim=1)  # :146-149
    c1, c2 = lead_choice
    lat_all = torch.cat([z1_mean, z2_mean], dim=1)  # :151
    lat_p = torch.cat([z1g[:, c1], z2_mean], dim=1)  # :159
    lat_l = torch.cat([z1_mean, z2g[:, c2]], dim=1)  # :160
    q = F.linear(theta_features(query_theta).view(B, -1), P["mlp2.weight"], P["mlp2.bias"])  # :163-164
    outs = [decoder(P, q[:, :, None] * lat, bn_training, stats_out, prec, name=f"dec{k}")
            for k, lat in enumerate((lat_all, lat_p, lat_l))]  # :166-176
    if phase == "train":
        return tuple(outs)
    if phase in ("val", "test"):
        rq = F.linear(theta_features(rest_theta), P["mlp2.weight"], P["mlp2.bias"])  # (B,V,256) :182-183
        rest = [decoder(P, rq[:, v, :, None] * lat_all, bn_training, stats_out, prec) for v in range(rq.shape[1])]
        return tuple(outs) + (torch.cat(rest, dim=1),)  # :185-192
    raise KeyError("please type correct phase")  # :194


def gen_ecg(P, z1, z2, query_theta, rois, stats_out=None):
    """Model_nefnet.gen_ecg, model_nefnet.py:196-218 (always eval-mode BN, :197)."""
    B = z1.shape[0]
    G = z1.shape[1] // 128
    z2 = roi_reverse(z2, rois)
    L4 = z1.shape[-1]
    lat_all = torch.cat([z1.view(B, G, 128, L4).mean(1), z2.view(B, G, 128, L4).mean(1)], dim=1)
    q = F.linear(theta_features(query_theta), P["mlp2.weight"], P["mlp2.bias"])  # (B,V,256)
    rest = [decoder(P, q[:, v, :, None] * lat_all, False, stats_out) for v in range(q.shape[1])]
    return torch.cat(rest, dim=1)


def standin_loss(out, out_p, out_l, target, factor=(0.5, 0.5, 1.0), loss_using=(1, 2, 3), reg_loss="l1_loss",
                 rest_out=None, rest_view=None):
    """loss/losses.py:21-50: factor0*L1(out.detach(), out_p) + factor1*L1(out.detach(), out_l) +
    factor2*reg(out, target); reg = L1 or MSE; optional unsupervised term on val."""
    reg = F.l1_loss if reg_loss == "l1_loss" else F.mse_loss
    zero = out.new_zeros(())
    l1 = F.l1_loss(out.detach(), out_p) if 1 in loss_using else zero
    l2 = F.l1_loss(out.detach(), out_l) if 2 in loss_using else zero
    l3 = reg(out, target) if 3 in loss_using else zero
    total = l1 * factor[0] + l2 * factor[1] + l3 * factor[2]
    res = (total, l1 * factor[0], l2 * factor[1], l3 * factor[2])
    if rest_out is not None and rest_view is not None:
        res = res + (reg(rest_out, rest_view),)
    return res


# ----------------------------------------------------------------------------------------
# one full train step on CPU (used by the cpu_baseline / --impl reference legs of bench.py)
# ----------------------------------------------------------------------------------------
def live_param_names(G: int):
    return [n for n, sh in param_shapes(G).items()
            if n not in UNUSED_PARAMS and "running_" not in n and "num_batches" not in n]


def train_step(P, inputs, lead_choice=(0, 0), lr=0.1, momentum=0.9, momentum_buf=None, keeps=None):
    """forward + Standin loss + backward + SGD(momentum) update in place
    (solver/solver.py:171-235, solver/optim_scheduler.py:10)."""
    G = inputs["x"].shape[1]
    names = live_param_names(G)
    for n in names:
        P[n].requires_grad_(True)
        P[n].grad = None
    stats = {k: v for k, v in P.items() if "running_" in k or "num_batches" in k}
    out, out_p, out_l = forward(P, inputs["x"], inputs["input_thetas"], inputs["query_theta"], inputs["rois"],
                                phase="train", lead_choice=lead_choice, keeps=keeps, stats_out=stats)
    loss = standin_loss(out, out_p, out_l, inputs["target"])[0]
    loss.backward()
    with torch.no_grad():
        for n in names:
            g = P[n].grad
            if momentum_buf is not None:
                buf = momentum_buf.get(n)
                if buf is None:
                    buf = momentum_buf[n] = g.clone()
                else:
                    buf.mul_(momentum).add_(g)
                g = buf
            P[n].add_(g, alpha=-lr)
    return float(loss.detach()), (out.detach(), out_p.detach(), out_l.detach())
