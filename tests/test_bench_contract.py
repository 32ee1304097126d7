"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm (the oracle port on the host
cores, one bounded step), the shape of the `e2e` object, and the `roofline.traffic` source."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config")


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in BASE_KEYS:
        assert k in line, k
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["unit"] == "tiles/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - line["value"]) < 1e-9
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert abs(e2e["value"] - line["value"]) < 1e-9 and e2e["unit"] == line["unit"]
    # the metric and workload are the ones BASELINE.json names
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "tiles/sec centerOffsetRes10 infer+decode" in base["metric"]          # its headline clause (the second clause,
    assert "tiles/sec centerOffsetRes10 infer+decode" in line["metric"]          # train samples/sec, is the `train` leg)


def test_e2e_entry_reports_both_forms_and_heads_the_faster_one():
    import bench
    e = bench.e2e_entry(world=2, B=64, K=20, u8_ms=60.0, f32_ms=55.0)
    assert e["form"] == "float32_tiles" and e["unit"] == "tiles/s"
    assert abs(e["value"] - 2 * 64 * 20 / 0.055) < 1e-6
    assert e["h2d_bytes_per_step"] == 64 * 512 * 512 * 4 and e["d2h_bytes_per_step"] == 10 * 64 * 100 * 4
    assert e["grey_bytes"]["h2d_bytes_per_step"] == 64 * 512 * 512
    e = bench.e2e_entry(world=1, B=64, K=20, u8_ms=50.0, f32_ms=55.0)
    assert e["form"] == "grey_bytes" and e["h2d_bytes_per_step"] == 64 * 512 * 512


def test_roofline_traffic_comes_from_the_committed_capture():
    import bench
    traffic, src = bench.heads_traffic()
    assert src == "profiles/ncu_full_r02.json"
    assert 0.9 * 570e6 < traffic < 1.1 * 570e6          # 537 MB of activations read once + weights + 29 MB written
    assert "workload" in bench.workload_config(64)
