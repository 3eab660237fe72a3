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
    # the reference's own files when build() staged them (oracle/_ref), else the port
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    expect = "reference" if importlib.import_module("ref_loader").available() else "port"
    assert d["cpu_baseline"]["kind"] == expect and d["cpu_baseline"]["cores"] >= 1
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
    _, mr = bench.algorithmic_bytes("mrr", 0, n, nnz)
    assert mr == b_spmv + 80.0 * n                                    # §8d's primary MrR figure
    # bytes the code really moves: CG == formula; the k-skip trips make 2k+1 passes over A instead of 3k+2
    assert bench.actual_bytes("cg", 0, n, nnz) == cg
    assert bench.actual_bytes("mrr", 0, n, nnz) == b_spmv + 104.0 * n
    b_a = 12.0 * nnz + 4.0 * (n + 1)
    assert abs(bench.actual_bytes("kskipcg", 4, n, nnz) * 5 - (9 * b_a + 4 * 32.0 * n + 8.0 * n * 11 + 56.0 * n + 4 * 48.0 * n + 16.0 * n)) < 1.0
    assert bench.actual_bytes("kskipcg", 4, n, nnz) < bench.algorithmic_bytes("kskipcg", 4, n, nnz)[1]
