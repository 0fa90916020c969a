"""GPU: mis-tiled ROIs are diagnosed like the reference does (network/utils/roi_pooling_1d.py:83-98: the truncated ROI
lengths of every segment must sum to L / 4, else torch.stack / torch.cat raise a RuntimeError).  The CPU half of this file
pins the rule itself against the oracle's roi_reverse (which is the reference's loop restated)."""
import pytest
import torch

from oracle import nefnet_oracle as O


def _bad_rois(rois, L):
    """Segment 1: roi 5 ends 4 samples before roi 6 starts, so its truncated lengths sum to L/4 - 1."""
    r = rois.clone()
    r[1, 5, 1] -= 4
    return r


def test_the_rule_matches_the_reference_loop_on_the_cpu():
    inp = O.make_inputs(3, 3, 512, 5)
    z = torch.randn(3, 8, 7, 32)
    O.roi_reverse(z, inp["rois"])                      # tiles: fine
    with pytest.raises(RuntimeError):
        O.roi_reverse(z, _bad_rois(inp["rois"], 512))  # torch.stack: unequal lengths


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sync", "deferred", "off"])
def test_mis_tiled_rois_raise(mode):
    import network
    dev = torch.device("cuda:0")
    G, B, L = 3, 3, 512
    torch.manual_seed(0)
    m = network.Model_nefnet(1, G).to(dev).train()
    m.roi_check = mode
    inp = O.make_inputs(B, G, L, 5)
    d = {k: v.to(dev) for k, v in inp.items()}
    bad = _bad_rois(inp["rois"], L).to(dev)
    # well-tiled ROIs never raise
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    sum(o.sum() for o in outs).backward()
    m.check_rois()
    if mode == "sync":
        with pytest.raises(RuntimeError, match="roi_pooling_reverse"):
            m(d["x"], d["input_thetas"], d["query_theta"], bad, phase="train")
    elif mode == "deferred":
        outs = m(d["x"], d["input_thetas"], d["query_theta"], bad, phase="train")    # launches are queued, no stall
        with pytest.raises(RuntimeError, match="segment 1"):
            sum(o.sum() for o in outs).backward()                                     # the same step's backward raises
        # eval: the verdict is due at the next entry point
        m.eval()
        with torch.no_grad():
            m(d["x"], d["input_thetas"], d["query_theta"], bad, phase="test", rest_theta=d["query_theta"][:, None, :])
            with pytest.raises(RuntimeError, match="roi_pooling_reverse"):
                m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="test", rest_theta=d["query_theta"][:, None, :])
    else:
        outs = m(d["x"], d["input_thetas"], d["query_theta"], bad, phase="train")    # clamps silently, as in round 1
        sum(o.sum() for o in outs).backward()
        assert all(bool(torch.isfinite(o).all()) for o in outs)
    # the module stays usable after a diagnosis
    m.train()
    outs = m(d["x"], d["input_thetas"], d["query_theta"], d["rois"], phase="train")
    sum(o.sum() for o in outs).backward()
    m.check_rois()
