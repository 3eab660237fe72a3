"""bench.py --impl reference (the CPU arm: oracle port on host cores) on a small workload: one JSON line with the
contract's keys.  Runs on CPU; the default 512^3 workload is exercised on the GPU box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "cg_p2d256", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "solver_iterations_per_s" and d["unit"] == "iterations/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["name"] == "cg_p2d256" and "workload" in d["config"]


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    n, nnz = 16777216, 117047296
    b_spmv, cg = bench.algorithmic_bytes("cg", 0, n, nnz)
    assert b_spmv == 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    assert cg == b_spmv + 72.0 * n and abs(cg / n - 175.7) < 0.1     # "≈ 176·N B" per CG iteration on the 7-point stencil
    _, ks = bench.algorithmic_bytes("kskipmrr", 8, n, nnz)
    assert abs(ks - ((26 * b_spmv + 8.0 * n * 19 + 72.0 * n * 9) / 9)) < 1.0
