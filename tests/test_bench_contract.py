"""The bench line keeps the driver's contract.  Behaviour that runs without a GPU is exercised live: the reference arm
(`bench.py --impl reference`, incl. the rank != 0 exit under torchrun's environment), the CPU-baseline helpers, the parsing
of the newest ncu summary into roofline.traffic / inst_per_px.  The B200 arm itself needs a GPU: its last committed lines
under profiles/ are checked for the contract's keys and internal consistency."""
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    def ver(p):
        m = re.search(r"r(\d+)_v(\d+)_", os.path.basename(p))
        return (int(m.group(1)), int(m.group(2))) if m else (0, 0)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=ver)
    assert files, pattern
    return json.loads(open(files[-1]).read())


def test_b200_line_has_every_contract_key():
    d = _latest("r*_v*_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity_checked"):
        assert k in d, k
    assert d["parity_checked"]["mismatches"] == 0 and d["parity_checked"]["frames"] >= 4 * d["n_gpus"]
    assert abs(d["value"] - d["n_gpus"] * d["config"]["frames_per_step_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"]      # DRAM bytes >= compulsory bytes
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_line_has_every_contract_key():
    d = _latest("r*_v*_bench_reference.json")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    b = _latest("r*_v*_bench.json")
    assert d["metric"] == b["metric"] and d["unit"] == b["unit"] and d["config"]["workload"] == b["config"]["workload"]


def test_reference_arm_runs_live_and_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["cores"] >= 1 and c["kind"] in ("reference", "port")
    from oracle import ref_lib
    assert c["kind"] == ("reference" if ref_lib.available() else "port")      # the reference's own code whenever it is built
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"]["frames_per_step"]) < 1e-6 * d["config"]["frames_per_step"] + 1e-3
    # under torchrun only rank 0 works and prints
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_front_end_and_primitive_timing():
    sys.path.insert(0, ROOT)
    import bench
    fr = bench.make_pairs(2, seed=3)
    fe = bench.CpuFrontEnd(2)
    try:
        f, k, dt = fe.step(fr, 1)
        assert f == 4 and dt > 0 and 4 * 1500 < k < 4 * 2100          # 2 threads x 1 pair x 2 frames, ~2000 keypoints each
        f2, k2, _ = fe.step(fr, 1)
        assert (f2, k2) == (f, k)                                     # persistent detectors: same work, same result
    finally:
        fe.close()
    prim = bench.primitive_times(fr[0], reps=1)
    assert prim["port_scalar"]["total"] > 0
    if prim.get("cv2_simd_1thread"):
        assert 0 < prim["cv2_simd_1thread"]["total"] < prim["port_scalar"]["total"]      # OpenCV's SIMD primitives beat the scalar restatement


def test_roofline_inputs_come_from_the_newest_ncu_summary():
    sys.path.insert(0, ROOT)
    import bench
    path, frames, rows = bench.newest_ncu_summary()
    assert re.match(r"r\d+_v\d+_ncu_full_summary\.md$", os.path.basename(path)) and frames > 0
    # one step = every pyramid level once: 7 eight-pixel launches (resize8) plus the four-pixel launches of the remainder columns
    assert any("fast_band" in k for k in rows) and sum(len(v) for k, v in rows.items() if "resize8" in k) == 7
    assert 7 <= sum(len(v) for k, v in rows.items() if "resize" in k) <= 14
    pix = bench.level_pixels(376, 1241)
    assert pix == 1444097
    traffic, inst_px, src = bench.pyrfast_from_summary(pix)
    assert src == os.path.relpath(path, ROOT)
    assert pix < traffic < 4 * pix and 20 < inst_px < 80              # between compulsory and the un-fused dataflow; issue-bound


def test_bench_cli():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--batch-sweep"):
        assert flag in out.stdout
