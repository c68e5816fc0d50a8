"""The bench.py output contract (CPU): the reference arm runs here end to end, and the committed B200 line of the
default workload (profiles/bench_r1_default.json) carries every key the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-clients", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "client_gan_steps_per_sec" and line["unit"] == "client-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_committed_b200_line_has_the_contract_keys():
    line = json.load(open(os.path.join(ROOT, "profiles", "bench_r1_default.json")))
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(line)
    assert line["n_gpus"] == 1 and line["dtype"] == "f32" and line["data"] == "synthetic" and line["scaling"] == "weak"
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is not None and 0.9 < r["traffic"] / r["bytes_per_launch"] < 1.2     # no wasted re-reads
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 1024 * 100 * 784 * 4 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    assert line["gpu_launches"] > 0
    c = line["clocks"]
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] / line["cpu_baseline"]["value"] > 100   # the north-star target
