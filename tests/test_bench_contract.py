"""CPU test of bench.py's reference arm (the oracle's C / OpenMP port timed on the
host on the SAME mesh as the GPU arm): the JSON line carries the keys the driver
reads, `steps` is what was actually timed and nothing is extrapolated."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20",
         "--warmup", "5", "--mesh-size", "0.8", "--no-c1"],
        capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["unit"] == "iters/s"
    assert line["metric"].startswith("optimizer iters/sec on 1M-elem 3D cantilever")
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    # the CPU leg is bounded: at most 2 timed steps, reported as such
    assert 1 <= line["steps"] <= 2 and line["steps_requested"] == 20 and line["dtype"] == "f64"
    assert line["config"]["workload"].startswith("custom: 3D cantilever toy_base(0.8)")
    assert abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert "no extrapolation" in cb["sample"]
    assert line["details"]["n_elem"] == 10 * 8 * 5
    assert line["e2e"] == {"value": line["value"], "unit": "iters/s",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_both_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.bench_config(bench.C2_MESH_SIZE)
    assert a == bench.bench_config(bench.C2_MESH_SIZE)
    assert a["workload"].startswith("C2: 3D cantilever toy_base(0.0577)")
    assert "52,728 hex" in a["same_config_c1"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
