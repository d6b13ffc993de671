"""bench.py's reference arm (the CPU leg the driver runs next to the GPU arm) prints one JSON line with the contract's
keys; runs here without a GPU (bounded sample: 50 queries x the TVR corpus, one step)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-sample-queries", "50"], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "query-video pairs scored+ranked/sec" and d["unit"] == "pairs/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "tvr_two_scale" and d["config"]["queries"] == 10895
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_frame_head_times_the_unmodified_reference():
    """--workload tvr_frame: the head the reference ships is timed on the UNMODIFIED reference (baseline/_ref, staged
    by oracle/install_ref.py), kind "reference"."""
    import pytest
    from oracle import install_ref
    install_ref.install()
    if not install_ref.staged():
        pytest.skip("the reference is not staged (no /root/reference here and no baseline/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tvr_frame",
                        "--steps", "1", "--warmup", "1", "--cpu-sample-queries", "50"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and d["config"]["workload"] == "tvr_frame"
    assert d["cpu_baseline"]["kind"] == "reference" and "baseline/_ref" in d["cpu_baseline"]["sample"]
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
