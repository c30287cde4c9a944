"""bench.py's JSON contract (CPU-side checks): the reference arm runs here (it only needs oracle/_ref), and the committed GPU lines of this
round (profiles/bench_r02_n*.json, produced on B200 boxes by `python bench.py [--gpus N]`) carry every key the contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhobbit_ref.so")):
        pytest.skip("oracle/_ref not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d) and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_gpu_lines_carry_the_contract(n):
    path = os.path.join(ROOT, "profiles", "bench_r02_n%d.json" % n)
    d = json.load(open(path))
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert d["n_gpus"] == n and d["metric"] == "pc_commit_field_elems_per_s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 16 << 26 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "reference"
        assert any(k.startswith("config1") for k in d["extras"]) and any(k.startswith("config3") for k in d["extras"])
    else:
        assert d["parity_check"]["commit"] is True and d["parity_check"]["sumcheck"] is True
