"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference's CPU decoder, compiled by oracle/Makefile,
timed on the host threads) prints ONE JSON line with the keys the driver reads -- checked here without a GPU, with the
smallest step count.  The GPU arm's line carries the same `config` (bench.workload_config is shared by both arms)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_mpix_per_s" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1
    assert d["vs_baseline"] is None and d["dtype"] == "u16" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frames" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert cfg["workload"].startswith("c2") and cfg["frames"] == 240 and (cfg["width"], cfg["height"]) == (1920, 1080)
    assert cfg["compression_type"] == 7
    assert set(d["workloads"]) == {"c1", "c3", "c4"} and all(w["value"] > 0 for w in d["workloads"].values())


def test_both_arms_describe_the_workload_with_the_same_function():
    sys.path.insert(0, ROOT)
    import bench
    streams, _ = bench.make_streams("c4", want_images=False)
    a = bench.workload_config("c4", streams, 1)
    b = bench.workload_config("c4", streams, 1)
    assert a == b and a["compression_type"] == 6 and a["frames"] == 64 and (a["width"], a["height"]) == (4000, 3000)
