"""Golden vectors for the callers next to the hot path, produced by the UNMODIFIED reference in the build container:
    python oracle/make_golden_data.py        (needs /root/reference; writes tests/golden/data_*.npz)

* data_tianchi.npz -- EcgTianChiInterval.__getitem__ (dataset/tianchi.py:84-232) on the two records the reference bundles
  (codes/data/tianchi), lead_num = 12 / super_mode '_12120', jitter off, for several heartbeats each; the heartbeat draw
  (random.sample, :98) is pinned per sample.  The raw records travel as int16 fixtures (they are integer ADC counts).
* data_psnr.npz    -- utils/mertic.py PSNR on seeded predictions, with and without rois, incl. an exact-match row.
numpy 2 removed np.float / np.int, which the reference still uses: they are aliased here, nothing else is patched.
skimage (imported by mertic.py at module scope for SSIM, not used by PSNR) is absent and stubbed."""
import importlib.util
import json
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference/codes"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    np.float, np.int = float, int  # removed aliases the reference relies on
    sk = types.ModuleType("skimage"); skm = types.ModuleType("skimage.metrics")
    skm.structural_similarity = None
    sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
    tianchi = load(os.path.join(REF, "dataset", "tianchi.py"), "ref_tianchi")
    mertic = load(os.path.join(REF, "utils", "mertic.py"), "ref_mertic")

    class Cfg:
        class DATA:
            train_label_path = os.path.join("/tmp", "nef_labels.txt")
            test_label_path = train_label_path
            train_data_root = os.path.join(REF, "data", "tianchi", "npy_data", "tianchi_train_round1")
            train_label_root = os.path.join(REF, "data", "tianchi", "tianchi_interval")
            lead_num = 12
            super_mode = "_12120"
            train_data_mode = "x"

        class MODEL:
            jitter_factor = 0
    ids = ["40723", "11315"]
    with open(Cfg.DATA.train_label_path, "w") as f:
        f.write("\n".join(i + ".json" for i in ids))
    ds = tianchi.EcgTianChiInterval(Cfg, "train")
    out = {}
    real_sample = random.sample
    n = 0
    for ri, rid in enumerate(ids):
        rec = np.load(os.path.join(Cfg.DATA.train_data_root, rid + ".npy"))
        assert rec.dtype == np.int64 and np.abs(rec).max() < 32768
        out["record%d" % ri] = rec.astype(np.int16)
        label = json.load(open(os.path.join(Cfg.DATA.train_label_root, rid + ".json")))
        nb = len(label["P on"])
        for beat in sorted(set([0, 1, nb // 2, nb - 2])):
            state = {"first": True}

            def pinned(pop, k, _beat=beat, _state=state):
                if _state["first"]:          # the heartbeat draw (:98)
                    _state["first"] = False
                    return [_beat]
                return real_sample(pop, k)
            tianchi.random.sample = pinned
            random.seed(100 + n)
            np.random.seed(100 + n)
            meta = ds[ri]
            tianchi.random.sample = real_sample
            end_point = label["P on"][beat + 1] if beat + 1 < nb else rec.shape[-1]
            marks = [label[k][beat] for k in ("P on", "P off", "R on", "R off", "T on", "T off")] + [end_point]
            tgt_idx = int(np.argmax([np.array_equal(meta["target_theta"], ds.theta[k].astype(np.float32)) for k in range(12)]))
            out["s%d/record" % n] = np.int64(ri)
            out["s%d/marks" % n] = np.array(marks, dtype=np.int64)
            out["s%d/target_index" % n] = np.int64(tgt_idx)
            out["s%d/data" % n] = meta["data"]
            out["s%d/rois" % n] = meta["rois"].astype(np.int64)
            out["s%d/target_view" % n] = meta["target_view"]
            out["s%d/ori_data" % n] = meta["ori_data"].astype(np.float32)
            assert np.array_equal(meta["target_view"], meta["ori_data"][tgt_idx].astype(np.float32))
            n += 1
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(OUT, "data_tianchi.npz"), **out)

    rng = np.random.default_rng(5)
    B, V, L = 5, 4, 512
    gt = rng.random((B, V, L), dtype=np.float32)
    pred = (gt + 0.05 * rng.standard_normal((B, V, L)).astype(np.float32)).astype(np.float32)
    pred[2, 1] = gt[2, 1]                      # rmse == 0 -> 100 (:16-17)
    rois = np.zeros((B, 7, 2), dtype=np.int64)
    rois[:, -1, 0] = [300, 512, 17, 480, 1]
    np.savez_compressed(os.path.join(OUT, "data_psnr.npz"), pred=pred, gt=gt, rois=rois,
                        psnr_rois=np.float64(mertic.PSNR(pred, gt, rois)), psnr_full=np.float64(mertic.PSNR(pred, gt)))
    print("wrote", n, "tianchi samples and the PSNR vector")


if __name__ == "__main__":
    main()
