"""bench.py contract pieces that need no GPU: the reference arm (CPU oracle of the same schedule on the host cores) prints
one JSON line with the keys the driver reads; configurations and flags exist as documented."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "decoys_per_sec_L150_dist_only" and d["unit"] == "decoys/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "decoys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "PyRosetta absent" in d["cpu_baseline"]["sample"] or d["cpu_baseline"]["kind"] == "reference"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_configurations_and_flags():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.CONFIGS) == {1, 2, 3, 4}
    assert bench.CONFIGS[2]["L"] == 300 and bench.CONFIGS[2]["two_model"] and bench.CONFIGS[1]["dist_only"] and bench.CONFIGS[3]["mc"]["cycles"] > 0
    assert bench.CONFIGS[4]["n_targets"] == 64 and bench.CONFIGS[4]["decoys_per_target"] == 100
    old = sys.argv
    try:
        sys.argv = ["bench.py", "--gpus", "8", "--scaling", "strong", "--config", "3", "--decoys", "2048", "--resident", "256"]
        a = bench.parse()
    finally:
        sys.argv = old
    assert (a.gpus, a.scaling, a.config, a.decoys, a.resident, a.impl) == (8, "strong", 3, 2048, 256, "b200")
