"""Host-side pieces of bench.py that need no GPU: the synthetic-COLMAP round trip of the views, the clock-sample
summary (which run is rejected / noted), the kernel-count table and the reference arm's rank gating."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from instascene_b200 import synth  # noqa: E402


def test_workload_views_travel_through_colmap_files(monkeypatch):
    monkeypatch.setitem(bench.WORKLOADS, "cfg3", dict(bench.WORKLOADS["cfg3"], P=500, n_views=12))
    wl, scene, cams = bench.build_workload("cfg3", 5)
    ref = synth.ring_cameras(12, wl["W"], wl["H"])
    assert len(cams) == 12 and scene.xyz.shape == (500, 3)
    for a, b in zip(cams, ref):
        assert np.allclose(a.full_proj_transform, b.full_proj_transform, atol=2e-6)
        assert np.allclose(a.world_view_transform, b.world_view_transform, atol=2e-6)
        assert abs(a.tanfovx - b.tanfovx) < 1e-12 and (a.image_width, a.image_height) == (b.image_width, b.image_height)


def test_clock_summary_flags_throttle_reasons():
    cs = bench.ClockSampler(0, enabled=False)
    assert cs.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    cs._max = 1965.0
    cs.samples = [(1965.0, 0x0, 650.0), (1950.0, 0x4, 990.0), (1800.0, 0x4 | 0x40, 1001.0)]
    s = cs.summary()
    assert s["sm_mhz"] == 1950.0 and s["sm_max_mhz"] == 1965.0 and s["samples"] == 3 and s["power_w_max"] == 1001.0
    assert s["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]


def test_traffic_record_is_stamped_with_its_kernel():
    """roofline.traffic comes from profiles/r2_traffic.json and only when the record was captured on the very kernel
    the bench times; anything else reports null (no stale constants)."""
    t, src = bench.ncu_traffic("cfg3", "no_such_kernel")
    assert t is None and src is None
    rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    for wl, r in rec.items():
        assert wl in bench.WORKLOADS and {"kernel", "dram_bytes_read", "dram_bytes_write", "source"} <= set(r)
        assert os.path.exists(os.path.join(ROOT, r["source"]))
        t, src = bench.ncu_traffic(wl, r["kernel"])
        assert t == r["dram_bytes_read"] + r["dram_bytes_write"] and src == r["source"]


def test_reference_arm_is_silent_on_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_prints_one_json_line(monkeypatch):
    """Bounded sample on a tiny workload: the line carries impl/cpu_baseline/e2e as the contract asks."""
    code = ("import bench, sys; bench.WORKLOADS['cfg3'] = dict(bench.WORKLOADS['cfg3'], P=3000, W=160, H=96);"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and "nothing extrapolated" in d["cpu_baseline"]["sample"]


def test_reference_arm_uses_every_core_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use all host cores."""
    code = ("import bench, sys; bench.WORKLOADS['cfg3'] = dict(bench.WORKLOADS['cfg3'], P=2000, W=96, H=64, samples=512);"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; bench.main()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="2", OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
