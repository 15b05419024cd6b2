"""The JSON line of bench.py: the reference arm runs here (it is CPU work by design) and must carry the contract's keys; the
committed line of the CUDA arm (profiles/, measured on a B200 by the same script) is checked for the same keys plus the
roofline / cpu_baseline / e2e objects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline")


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--candidates", "20000", "--cpu-sample-per-core", "1000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "banded_sw_gcups" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    for key in BASE_KEYS:
        assert key in d, key
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_cuda_arm_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r2_aq_bench_default.json")            # the default command on one B200, final state of round 2
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    for key in BASE_KEYS + ("gpu_launches", "clocks", "roofline"):
        assert key in d, key
    assert d["metric"] == "banded_sw_gcups" and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and 0 < c["value"] < d["value"] / 50
    # the two side measurements of the default line carry their own end-to-end figure and CPU baseline
    p = d["pairs_pipeline"]
    assert p["unit"] == "pairs/s" and 0 < p["e2e"]["value"] < p["value"] and p["e2e"]["h2d_bytes_per_step"] > 0 and p["e2e"]["d2h_bytes_per_step"] > 0
    assert p["cpu_baseline"]["kind"] == "reference" and 0 < p["cpu_baseline"]["value"] < p["value"] / 50
    g = d["gap_realigner"]
    assert g["unit"] == "fragments/s" and 0 < g["e2e"]["value"] < g["value"] and g["cpu_baseline"]["kind"] == "reference"
    assert 0 < g["cpu_baseline"]["value"] < g["e2e"]["value"] / 5 and g["realigned_fragments"] > 0
    assert 0 < d["roofline"]["frac_vs_16x2_peak"] < d["roofline"]["frac"]
