"""Print, for every golden case, how far the GPU solve is from the reference: iteration counts, max relative
history deviation over the first 50 solver iterations and overall, true residual.  Run on the GPU box."""
import os, sys, json
os.environ.setdefault("PK_QUIET", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import parallel_krylov_b200 as pk
import krylov_oracle as oracle
from golden_util import CASES, inputs, load

rows = []
for c in CASES:
    g = load(c); mat, b = inputs(c)
    kw = {"tol": c["tol"], "maxiter": c["maxiter"]}
    if c["k"] is not None: kw["k"] = c["k"]
    x, info = getattr(pk, c["solver"])(mat, b, **kw)
    res = info["residual"].cpu().numpy(); nosl = info["nosl"].cpu().numpy()
    m = min(len(res), len(g["residual"]))
    m50 = min(m, int(np.searchsorted(g["nosl"], 50, side="right")))
    rel = np.abs(res[:m] - g["residual"][:m]) / g["residual"][:m]
    rows.append({"id": c["id"], "it_ref": int(g["nosl"][-1]), "it": int(nosl[-1]),
                 "dev50": float(rel[:m50].max()), "dev_all": float(rel.max()),
                 "true": oracle.true_relres(mat, b, x.cpu().numpy()), "conv": info["converged"]})
    r = rows[-1]
    print(f"{r['id']:58s} it {r['it_ref']:4d}/{r['it']:4d} dev50 {r['dev50']:.1e} all {r['dev_all']:.1e} true {r['true']:.2e}")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_table.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"), indent=1)
