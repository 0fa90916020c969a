"""``Model_nefnet2`` (reference: network/model_nefnet2.py:63-203) on the B200 kernels.

The reference applies ONE single-lead trunk (encoder, angular scaling, w_conv, z1 / z2 branches) to every lead in a Python
loop, then two plain k3 convolutions (``single_conv_z1`` after z1_conv, ``single_conv_z2`` after roi_pooling_reverse), the
lead mean / shuffle and the same decoder as ``Model_nefnet``.  Here the leads stay the channel groups of the grouped
kernels (one launch per layer for all leads, as in ``Model_nefnet``) and every group reads -- and accumulates its weight
gradient into -- the SAME weight slice (NefPlan variant 2, include/nefnet_b200.h); the two extra convolutions are linear, so
they run on the lead means and the picked leads (3 x 2 small launches on B segments) instead of on all B x lead_num ones.

state_dict keys, shapes and registration order are the reference's: the lead_num = 1 key set of ``Model_nefnet`` with
``single_conv_z1.0.*`` / ``single_conv_z2.0.*`` before the decoder (checked against the unmodified class in
tests/test_oracle_golden.py through oracle/nefnet2_oracle.py).  Like the reference, ``network.build_model`` does not build
this class (network/__init__.py:7-12); construct it directly.  ``phase='gen'`` / ``gen_ecg`` are not provided: the
reference's ``gen_ecg`` (:205-227) cannot consume what its own ``phase='gen'`` returns."""
from __future__ import annotations

from .model_nefnet import Model_nefnet, _param_specs


class Model_nefnet2(Model_nefnet):
    _variant = 2

    def _make_specs(self):
        base = _param_specs(1)                      # the single-lead trunk: shapes do not depend on lead_num
        extra = [("single_conv_z1.0.weight", (128, 128, 3), "w"), ("single_conv_z1.0.bias", (128,), "b"),
                 ("single_conv_z2.0.weight", (128, 128, 3), "w"), ("single_conv_z2.0.bias", (128,), "b")]
        specs = []
        for sp in base:
            if sp[0] == "decoder.1.double_conv.0.weight":      # model_nefnet2.py:102-107 precede the decoder (:109)
                specs += extra
            specs.append(sp)
        return specs, [sp[0] for sp in base] + [sp[0] for sp in extra]   # C order: nef_param_name_v(G, 2, i)

    def _build_parameters(self):
        lead_num, self.lead_num = self.lead_num, 1     # ConvTranspose fan-in etc. of the single-lead modules
        try:
            super()._build_parameters()
        finally:
            self.lead_num = lead_num

    def forward(self, x, input_thetas, query_theta, rois, rest_theta=None, phase="train"):
        if phase == "gen":
            raise NotImplementedError("Model_nefnet2 (B200): phase 'gen' is not built (the reference's gen_ecg cannot consume it)")
        return super().forward(x, input_thetas, query_theta, rois, rest_theta=rest_theta, phase=phase)

    def gen_ecg(self, z1, z2, query_theta, rois):
        raise NotImplementedError("Model_nefnet2 (B200): gen_ecg is not built (model_nefnet2.py:205-227 feeds lead means to "
                                  "roi_pooling_reverse, which needs the per-ROI codes)")
