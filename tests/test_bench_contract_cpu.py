"""The reference arm of bench.py runs on CPU only: its JSON line must carry the contract's keys (the driver computes the
speed-up from it).  Small n so that the whole CPU suite stays within minutes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--n", "512"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["higher_is_better"] is True
    assert line["metric"] == "fp64_gflops_mm_solve_cholesky" and line["value"] > 0
    assert line["config"]["n"] == 512 and "n=512" in line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data"):
        assert k in line
