"""CPU: the byte / unit accounting bench.py reports is the one SURVEY 8(d) states (no GPU, no timing)."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    """286.0 MB / segment forward at G = 12, L = 5000, n_dec = 3 (nominal); 1106.7 MB at L = 20000;
    528.3 MB for the 24-view sweep; a train step is 3 x forward -> 219.6 GB at batch 256."""
    nominal = bench.algorithmic_bytes_per_segment(12, 5000, live=False)
    assert nominal / 1e6 == pytest.approx(286.0, abs=0.05)
    assert 3 * nominal * 256 / 1e9 == pytest.approx(219.6, abs=0.1)
    assert bench.algorithmic_bytes_per_segment(12, 20000, live=False) / 1e6 == pytest.approx(1106.7, abs=0.1)
    assert bench.algorithmic_bytes_per_segment(12, 5000, n_dec=24, live=False) / 1e6 == pytest.approx(528.3, abs=0.1)
    assert bench.survey_live_bytes_per_segment(12, 5000) / 1e6 == pytest.approx(241.8, abs=0.05)     # SURVEY's own live figures
    assert bench.survey_live_bytes_per_segment(12, 20000) / 1e6 == pytest.approx(930.1, abs=0.3)
    live = bench.algorithmic_bytes_per_segment(12, 5000, live=True)
    assert live < nominal and (nominal - live) == 4 * 4 * 128 * 12 * 1250   # z2_conv1 on the live window
    assert 3 * live * 256 / 1e9 == pytest.approx(196.04, abs=0.01)


def test_forward_report_and_workload_names():
    r = bench.forward_report(18.6, 1, 256, 12, 5000, 6547.2)
    assert r["hbm_frac_nominal"] == pytest.approx(0.60, abs=0.005)           # SURVEY: 18.6 ms <=> 60 % of 6547 GB/s
    assert r["segments_per_s"] == pytest.approx(256 / 0.0186)
    assert "NCCL" in bench.workload_name(8, 256, 5000) and "NCCL" not in bench.workload_name(1, 256, 5000)
    assert "24-view" in bench.workload_name(8, 64, 5000, "sweep") and set(bench.CONFIGS) == {2, 4, 5}
    assert bench.METRIC.startswith("ECG segments/sec") and bench.UNIT == "segments/s"


def test_dominant_kernel_traffic_fixture_is_the_bench_shape():
    """roofline.traffic comes from the committed ncu capture; it must describe the shape bench.py times by default and agree
    with the kernel's algorithmic bytes (no wasted re-reads)."""
    t = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
    assert (t["B"], t["G"], t["L"]) == (256, 12, 5000)
    traffic = bench.measured_traffic(256, 12, 5000)
    alg = 2.0 * 128 * 12 * 1250 * 256 * 2          # the fp16 copies of dY and X, each read once
    assert traffic is not None and 0.95 < traffic / alg < 1.10
    assert bench.measured_traffic(64, 12, 20000) is None


def test_profile_summaries_regenerate_from_the_committed_launch_lists():
    """profiles/*_summary.txt are tools/ output over the committed ncu launch lists: regenerate and compare, so the per-kernel
    shares quoted in DESIGN.md can be traced to raw captures."""
    import subprocess
    prof = os.path.join(ROOT, "profiles")
    for tool, args, summary in (
            ("kernel_metrics.py", ["r01_step_b256_per_kernel_metrics.csv", "6457.4"], "r01_step_b256_per_kernel_metrics_summary.txt"),
            ("kernel_metrics.py", ["r02_step_b256_per_kernel_metrics.csv", "6457.4"], "r02_step_b256_per_kernel_metrics_summary.txt"),
            ("launch_summary.py", ["r01_launches_step_b256.csv", "0", "--second-half"], "r01_launches_step_b256_summary.txt")):
        cmd = [sys.executable, os.path.join(ROOT, "tools", tool), os.path.join(prof, args[0])] + args[1:]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        want = [l.rstrip() for l in open(os.path.join(prof, summary)).read().strip().splitlines()]
        got = [l.rstrip() for l in out.stdout.strip().splitlines()]
        assert got[:len(want)] == want or want[:len(got)] == got, (tool, got[:3], want[:3])
        assert any("wgrad" in l for l in got[:3])          # the largest share of the step, as DESIGN.md says


def test_bench_without_a_gpu_fails_loudly():
    """The product arm has no CPU fallback: without a CUDA device bench.py exits non-zero and says so (it does not quietly time
    the oracle)."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout) and not r.stdout.strip().startswith("{")
