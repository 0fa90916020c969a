"""CPU, build container only: the oracle's closed forms of the two ROI operators against the reference's own functions
(network/utils/roi_pooling_1d.py, imported from /root/reference) on randomised ROI tables -- a direct pin of SURVEY F7 (what
`roi_algin` actually computes) and of the `.long()` truncation / empty-ROI rules, beyond the end-to-end golden vectors."""
import importlib.util
import os

import pytest
import torch

from oracle import nefnet_oracle as O

REF = os.path.join(os.environ.get("NEF_REFERENCE", "/root/reference/codes"), "network", "utils", "roi_pooling_1d.py")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout exists only in the build container")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("ref_roi_pooling_1d", REF)   # one file, no package imports of its own
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _tables(gen, B, L, ragged):
    rois = torch.zeros(B, 7, 2, dtype=torch.long)
    for b in range(B):
        if ragged:   # arbitrary sorted cut points: non multiples of 4, repeats (empty ROIs)
            cuts = torch.sort(torch.randint(0, L + 1, (6,), generator=gen)).values
        else:
            cuts = torch.sort(torch.randint(1, L // 4, (6,), generator=gen)).values * 4
        e = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.full((1,), L, dtype=torch.long)])
        rois[b, :, 0], rois[b, :, 1] = e[:-1], e[1:]
    return rois


@pytest.mark.parametrize("L4", [16, 128, 1250, 333])
@pytest.mark.parametrize("ragged", [False, True])
def test_roi_align_closed_form_equals_grid_sample(ref, L4, ragged):
    gen = torch.Generator().manual_seed(L4 + int(ragged))
    B, Cn = 3, 5
    z = torch.randn(B, Cn, L4, generator=gen)
    rois = _tables(gen, B, 4 * L4, ragged)
    want = ref.roi_algin(z, rois.clone(), size=16, spatial_scale=128 / 512)      # model_nefnet.py:136
    got = O.roi_align_center(z, rois)
    assert got.shape == want.shape == (B, Cn, 7, 16)
    torch.testing.assert_close(got, want, rtol=2e-6, atol=2e-7)
    assert rois.dtype == torch.long                                            # int64 rois are not mutated (:50 copies)


@pytest.mark.parametrize("L4", [16, 128, 1250])
def test_roi_reverse_equals_reference_on_tiling_tables(ref, L4):
    gen = torch.Generator().manual_seed(7 * L4)
    B, Cn = 3, 4
    z = torch.randn(B, Cn, 7, 32, generator=gen)
    rois = _tables(gen, B, 4 * L4, ragged=False)
    rois[1, 2, 0] = rois[1, 2, 1] = rois[1, 1, 1]                              # an empty ROI in one segment
    rois[1, 3, 0] = rois[1, 2, 1]
    want = ref.roi_pooling_reverse(z, rois.clone(), spatial_scale=128 / 512)   # model_nefnet.py:143
    got = O.roi_reverse(z, rois, out_len=L4)
    torch.testing.assert_close(got, want, rtol=0, atol=0)


def test_roi_reverse_truncation_cases(ref):
    """Cut points that are not multiples of 4: each ROI's length is long(r1 / 4) - long(r0 / 4) (:83-85)."""
    z = torch.arange(2 * 7 * 32, dtype=torch.float32).view(1, 2, 7, 32)
    rois = torch.tensor([[[0, 17], [17, 18], [18, 30], [30, 33], [33, 47], [47, 47], [47, 64]]])
    want = ref.roi_pooling_reverse(z, rois.clone(), spatial_scale=128 / 512)
    got = O.roi_reverse(z, rois)
    assert got.shape == want.shape == (1, 2, 16)
    torch.testing.assert_close(got, want, rtol=0, atol=0)


def _load(rel, name):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(REF))), *rel.split("/"))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_theta_features_equal_the_reference_encoder():
    """utils/theta_encoder.py:13-29 for the three shapes the model feeds it: (B, G, 2), (B, 2) viewed as (B, 1, 2), (B, V, 2)."""
    enc = _load("network/utils/theta_encoder.py", "ref_theta_encoder").ThetaEncoder(encoder_len=1)
    gen = torch.Generator().manual_seed(5)
    for shape in ((4, 12, 2), (4, 24, 2), (3, 1, 2)):
        th = (torch.rand(*shape, generator=gen) - 0.5) * 6.3
        torch.testing.assert_close(O.theta_features(th), enc(th), rtol=0, atol=0)
    q = torch.rand(5, 2, generator=gen)
    torch.testing.assert_close(O.theta_features(q), enc(q).view(5, -1), rtol=0, atol=0)     # model_nefnet.py:163-164


@pytest.mark.parametrize("reg_loss,using", [("l1_loss", [1, 2, 3]), ("l2_loss", [1, 2, 3]), ("l1_loss", [3]), ("l1_loss", [1, 3])])
def test_standin_loss_equals_the_reference_wrapper(reg_loss, using):
    """loss/losses.py:21-50 incl. the detach of `predict` in the two stand-in terms (gradients compared too)."""
    ref_losses = _load("network/loss/losses.py", "ref_losses")

    class Cfg:
        class SOLVER:
            pass
    Cfg.SOLVER.reg_loss, Cfg.SOLVER.loss_using, Cfg.SOLVER.loss_factor = reg_loss, using, [0.5, 0.25, 2.0]
    gen = torch.Generator().manual_seed(11)
    base = [torch.rand(3, 1, 64, generator=gen) for _ in range(4)]
    rest = torch.rand(3, 4, 64, generator=gen), torch.rand(3, 4, 64, generator=gen)
    grads = []
    vals = []
    for fn in ("ref", "oracle"):
        o, p, l = (t.clone().requires_grad_(True) for t in base[:3])
        if fn == "ref":
            res = ref_losses.losswrapper(o, p, l, base[3], Cfg, rest[0], rest[1])
        else:
            res = O.standin_loss(o, p, l, base[3], factor=(0.5, 0.25, 2.0), loss_using=tuple(using), reg_loss=reg_loss,
                                 rest_out=rest[0], rest_view=rest[1])
        res[0].backward()
        vals.append([float(v.detach()) if torch.is_tensor(v) else float(v) for v in res])
        grads.append([t.grad if t.grad is not None else torch.zeros_like(t) for t in (o, p, l)])
    assert vals[0] == pytest.approx(vals[1], rel=1e-6, abs=1e-9) and len(vals[0]) == 5
    for a, b in zip(*grads):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-9)
