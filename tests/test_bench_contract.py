"""bench.py's reference arm runs without a GPU: its JSON line must carry the contract's keys (CPU test, tiny shape)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--m", "400", "--n", "800",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["higher_is_better"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "400x800" in cb["sample"]
    # the line's time per step is the MEASURED time of the steps it ran (not an extrapolation)
    assert abs(line["ms_per_step"] * 2 / 1e3 - cb["seconds"]) < 1e-6
    assert abs(line["value"] - 2 / cb["seconds"]) < 1e-9 * line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "400x800" in line["config"]["workload"]
