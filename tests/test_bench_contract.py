"""The driver contract of bench.py: one JSON line with the keys the round-end harness reads."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line(refcpu):
    """`--impl reference`: the reference's own CPU renderer (oracle/_ref) on the host cores, same metric and config."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["config"]["workload"] == "car_only_1080p"
    assert BASE_KEYS <= set(d) and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_b200_arm_line(rt):
    if rt.device_count() < 1:
        pytest.skip("no CUDA device")
    d = run_bench("--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--also", "")
    assert BASE_KEYS <= set(d) and d["metric"] == "Mrays/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3
    assert d["config"]["workload"] == "car_only_1080p" and "model" not in d["config"] and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert d["value"] > 500 and abs(d["value"] - d["rays_per_frame"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    assert d["gpu_launches"] == 4 * 3  # per timed frame: the render kernel + the two kernels that order the next frame's tiles
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0 and e["d2h_bytes_per_step"] == 1920 * 1080 * 4 and e["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    # the shipped scenes are cache resident: the bound is the measured L1 record-gather rate (HBM figure kept beside it)
    assert r["bound"] == "l1_gather" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert r["frac_hbm_stream"] > 0 and d["frame_equals_1gpu"] is None and "run" in d and "ref_gpu_baseline" in d
    g = d["roofline_gather"]
    assert g["l1_resident_64KB"] > g["l2_resident_4MB"] > g["hbm_8GB"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
