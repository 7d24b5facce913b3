"""bench.py's JSON contract, checked on CPU through the reference arm (the oracle on the host cores); the GPU
arm prints the same keys plus roofline / clocks / gpu_launches and is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "K2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_beam_scores_per_s"
    assert line["unit"] == "scores/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("K2") and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["data"] == "synthetic" and line["dtype"] == "f64"


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_native_e2e_helper_runs_over_the_c_abi(oracle):
    """csrc/e2e_host.cpp (bench.py's `e2e_native`) is a plain C-ABI client: exercised here against the oracle library,
    which exports the same entry points (the GPU box runs it against libgms.so)."""
    sys.path.insert(0, ROOT)
    import bench

    wl = dict(bench.WORKLOADS["K2"])
    scans = bench.make_scans(wl, 3)
    r = bench.native_e2e(oracle.path, wl, 64, scans, 2, warm=1)
    assert r and "error" not in r, r
    assert r["steps"] == 2 and r["ms_per_step"] > 0 and r["value"] > 0 and r["h2d_bytes_per_step"] == 25 * wl["B"]
