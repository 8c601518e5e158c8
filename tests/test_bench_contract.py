"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the oracle restatement on the host cores)
prints one JSON line with the agreed keys, and the roofline traffic is read from the committed ncu summary."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_agreed_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "32", "--subdiv", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "winding queries/sec" and line["unit"] == "Gqueries/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_traffic_comes_from_the_committed_profile():
    sys.path.insert(0, ROOT)
    import bench

    t = bench.profile_traffic()
    assert t is not None and os.path.exists(os.path.join(ROOT, t["source"].split(" ")[0]))
    assert t["queries_per_launch"] == 262144 * 512 and t["static"] is True  # labelled: read from a committed profile, not measured live
    assert 2e8 < t["bytes_per_launch"] < 4e9  # ~10 bytes per query: the records the launch touches, from HBM, plus the output


def test_reference_arm_uses_every_host_thread_and_shares_the_config_keys():
    """under torchrun OMP_NUM_THREADS=1 must not make the CPU arm single-threaded, and both arms carry the
    same `config` keys (the driver compares them)."""
    import argparse

    sys.path.insert(0, ROOT)
    import bench

    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "5", "--points", "20000", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == bench.HOST_THREADS
    args = argparse.Namespace(gpus=1, mode="tree", hierarchy="reference", leaf_size=4)
    assert set(line["config"]) == set(bench.config_dict(args, "x", 1))
    # a non-zero rank of the reference arm exits quietly
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "16", "--subdiv", "1"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_exact_mode_reference_arm():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mode", "exact", "--points", "4000", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["config"]["mode"] == "exact" and "exact mode" in line["config"]["workload"] and line["value"] > 0
