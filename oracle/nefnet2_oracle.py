"""CPU oracle for the ``Model_nefnet2`` variant.  TEST INFRASTRUCTURE ONLY.

Functional restatement of ``codes/network/model_nefnet2.py:63-203``: ONE single-lead trunk (encoder,
angular scaling, w_conv, z1 / z2 branches) shared by all leads and applied lead by lead, two extra plain
k3 convolutions (``single_conv_z1`` after z1_conv, ``single_conv_z2`` after roi_pooling_reverse), then
the same lead mean / shuffle / query scaling / decoder as ``Model_nefnet``.  Because the trunk weights
are shared, "lead by lead" equals running the 1-lead trunk of ``oracle/nefnet_oracle.py`` on the
(B*lead_num) single-lead segments obtained by folding the leads into the batch; that is how it is
restated here, and it is the dataflow the CUDA path will use (the same kernels with G = 1 groups).

Pinning: ``oracle/make_golden_nefnet2.py`` runs the unmodified reference class on the seeded inputs
below and commits the vectors (tests/golden/nefnet2_*.npz); tests/test_oracle_golden.py re-checks them.

The reference never constructs this class (``network/__init__.py:7-12`` builds ``model_nefnet`` only) and
its ``gen_ecg`` (:205-227) cannot consume what its own ``phase='gen'`` returns (lead means after the
reverse, :159-160, fed to ``roi_pooling_reverse`` again), so ``gen_ecg`` is not restated.

The product package has NO CUDA path for this variant yet; nothing under
``electrocardio-panorama_b200/`` imports this file.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import nefnet_oracle as O

UNUSED_PARAMS = (  # never touched by forward: grads stay None in the reference
    "w_feature_extractor.0.weight", "w_feature_extractor.0.bias",
    "w_conv.0.residual_conv.weight", "w_conv.0.residual_conv.bias",
    "z2_conv2.0.residual_conv.weight", "z2_conv2.0.residual_conv.bias",
)


def param_shapes() -> "Dict[str, Tuple[int, ...]]":
    """state_dict key -> shape in registration order (model_nefnet2.py:68-115).  Independent of
    lead_num: the 1-lead key set of Model_nefnet with the two single_conv_* layers inserted before
    the decoder."""
    base = O.param_shapes(1)
    out: Dict[str, Tuple[int, ...]] = {}
    for k, v in base.items():
        if k == "decoder.1.double_conv.0.weight":
            out["single_conv_z1.0.weight"] = (128, 128, 3)  # :102-104
            out["single_conv_z1.0.bias"] = (128,)
            out["single_conv_z2.0.weight"] = (128, 128, 3)  # :105-107
            out["single_conv_z2.0.bias"] = (128,)
        out[k] = v
    return out


def make_params(seed: int = 0) -> "Dict[str, torch.Tensor]":
    """Deterministic weights: the G = 1 stream of nefnet_oracle.make_params plus the two extra convs
    drawn from their own generator (PyTorch default uniform bounds)."""
    base = O.make_params(1, seed)
    gen = torch.Generator().manual_seed(7000003 * seed + 2)
    out: Dict[str, torch.Tensor] = {}
    bound = 1.0 / math.sqrt(128 * 3)
    for name, shape in param_shapes().items():
        if name in base:
            out[name] = base[name]
        else:
            out[name] = (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound
    return out


def live_param_names():
    return [n for n in param_shapes()
            if n not in UNUSED_PARAMS and "running_" not in n and "num_batches" not in n]


def lead_latents(P, x, input_thetas, rois):
    """model_nefnet2.py:126-151 for all leads at once.  x (B,G,L) -> z1, z2 of shape (B,G,128,L/4)."""
    B, G, L = x.shape
    xs = x.reshape(B * G, 1, L)  # lead i of segment b -> row b*G+i  (:127 x[:, i:i+1])
    th = input_thetas.reshape(B * G, 1, 2)  # (:128)
    r = rois[:, None].expand(B, G, *rois.shape[1:]).reshape(B * G, *rois.shape[1:])  # same rois for every lead
    z1, z2 = O.latents(P, xs, th, r, 1)  # :130-147 (without :140) with the shared single-lead weights
    z1 = F.conv1d(z1, P["single_conv_z1.0.weight"], P["single_conv_z1.0.bias"], padding=1)  # :140
    z2 = F.conv1d(z2, P["single_conv_z2.0.weight"], P["single_conv_z2.0.bias"], padding=1)  # :148
    L4 = z1.shape[-1]
    return z1.view(B, G, 128, L4), z2.view(B, G, 128, L4)


def forward(P, x, input_thetas, query_theta, rois, rest_theta=None, phase="train", lead_choice=(0, 0),
            bn_training=True, stats_out=None):
    """Model_nefnet2.forward, model_nefnet2.py:117-203.  Dropout off (exact-parity convention of
    nefnet_oracle.forward); ``lead_choice`` = the two random.randint draws (:163,165; z1 first)."""
    B = x.shape[0]
    z1g, z2g = lead_latents(P, x, input_thetas, rois)
    z1_mean, z2_mean = z1g.mean(dim=1), z2g.mean(dim=1)  # :154-155
    if phase == "gen":
        return z1_mean, z2_mean  # :159-160
    c1, c2 = lead_choice
    lat_all = torch.cat([z1_mean, z2_mean], dim=1)  # :157
    lat_p = torch.cat([z1g[:, c1], z2_mean], dim=1)  # :168
    lat_l = torch.cat([z1_mean, z2g[:, c2]], dim=1)  # :169
    q = F.linear(O.theta_features(query_theta).view(B, -1), P["mlp2.weight"], P["mlp2.bias"])  # :172-173
    outs = [O.decoder(P, q[:, :, None] * lat, bn_training, stats_out) for lat in (lat_all, lat_p, lat_l)]
    if phase == "train":
        return tuple(outs)  # :187-188
    if phase in ("val", "test"):
        rq = F.linear(O.theta_features(rest_theta), P["mlp2.weight"], P["mlp2.bias"])  # :191-192
        rest = [O.decoder(P, rq[:, v, :, None] * lat_all, bn_training, stats_out) for v in range(rq.shape[1])]
        return tuple(outs) + (torch.cat(rest, dim=1),)  # :193-201
    raise KeyError("please type correct phase")  # :203
