"""Micro-benchmarks of the individual kernels through the C-ABI (CUDA events, warm-up, inputs >> L2).
    python tools/kernel_bench.py [p3d256|p3d512|band32m|p3d128] [reps]
"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PK_QUIET", "1")
import torch
from parallel_krylov_b200 import device_problems as dp, _lib
from parallel_krylov_b200._core import Context, Operator, _ptr

which = sys.argv[1] if len(sys.argv) > 1 else "p3d256"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = Context.get(0)
lib = ctx.lib
if which.startswith("dense"):
    nd = int(which[5:] or 16384)
    a = torch.empty(nd, nd, dtype=torch.float64, device="cuda").normal_()
    op = Operator.from_dense_tensor(a, ctx)
    ld = op.ld
    xs = [dp.hash_normal(i, ld) for i in range(2)]
    y0 = torch.empty(ld, dtype=torch.float64, device="cuda"); y1 = torch.empty_like(y0)
    sums = torch.zeros(3, dtype=torch.float64, device="cuda")
    N0 = C.c_void_p(0)
    def timeit(fn, nbytes, label):
        for _ in range(3): fn()
        ctx.sync(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        print(f"{label:34s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s  {100*nbytes/ms/1e6/6556.5:5.1f}% of measured peak")
    bts = 8.0 * nd * nd + 16.0 * nd
    timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(xs[0]), _ptr(y0), N0, N0, N0, N0)), bts, f"gemv {nd}x{nd}")
    timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(xs[0]), _ptr(y0), N0, N0, _ptr(xs[0]), _ptr(sums))), bts, "gemv + fused dots")
    timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(xs[0]), _ptr(y0), _ptr(xs[1]), _ptr(y1), N0, N0)), bts, "gemv two chains (one pass over A)")
    ref = a @ xs[0][:nd]
    print("max rel err vs torch:", float((y0[:nd] - (a @ xs[0][:nd])).abs().max() / ref.abs().max()))
    sys.exit(0)
if which.startswith("p3d"):
    d = int(which[3:]); rowptr, col, val, n = dp.stencil_csr(d, d, d)
elif which.startswith("p2d"):
    d = int(which[3:]); rowptr, col, val, n = dp.stencil_csr(d, d, 1)
else:
    n = 1 << 25 if which == "band32m" else 1 << 22
    rowptr, col, val, n = dp.banded_csr(n, 13, 0)
if os.environ.get("KB_HALF"):          # first half of the rows only (what one of two ranks owns)
    h = n // 2
    nz = int(rowptr[h].item())
    rowptr, col, val = rowptr[: h + 1].clone(), col[:nz].clone(), val[:nz].clone()
nnz = val.numel()
op = Operator.from_csr_tensors(rowptr, col, val, n, ctx)
n_rows = rowptr.numel() - 1
ld = op.ld
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(f"{which}: n={n} nnz={nnz} {op.kernel_info()} peak={PEAK}")
vecs = [dp.hash_normal(i, ld) for i in range(4)]
y0 = torch.empty(ld, dtype=torch.float64, device="cuda"); y1 = torch.empty_like(y0)
sums = torch.zeros(3, dtype=torch.float64, device="cuda")

def timeit(fn, nbytes, label):
    for _ in range(3): fn()
    ctx.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    gbs = nbytes / ms / 1e6
    print(f"{label:34s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured peak")
    return ms

b_spmv = 12.0 * nnz + 4.0 * (n_rows + 1) + 16.0 * n_rows
N0 = C.c_void_p(0)
timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(vecs[0]), _ptr(y0), N0, N0, N0, N0)), b_spmv, "spmv")
timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(vecs[0]), _ptr(y0), N0, N0, _ptr(vecs[0]), _ptr(sums))), b_spmv, "spmv + fused dots (w = x)")
timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(vecs[0]), _ptr(y0), N0, N0, _ptr(vecs[1]), _ptr(sums))), b_spmv + 8.0 * n, "spmv + fused dots (w != x)")
timeit(lambda: _lib.check(lib.pk_spmv(ctx.handle, op.handle, _ptr(vecs[0]), _ptr(y0), _ptr(vecs[1]), _ptr(y1), N0, N0)), b_spmv + 16.0 * n, "spmv two chains (bytes of ONE pass over A)")
timeit(lambda: _lib.check(lib.pk_dot(ctx.handle, n, _ptr(vecs[0]), _ptr(vecs[1]), _ptr(sums))), 16.0 * n, "dot")
for k in (2, 4, 8):
    nu, nv = k + 2, k + 1
    if (nu + nv) * ld * 8 > 60e9: continue
    U = torch.empty(nu * ld, dtype=torch.float64, device="cuda").normal_()
    V = torch.empty(nv * ld, dtype=torch.float64, device="cuda").normal_()
    g = torch.zeros(6 * (k + 2), dtype=torch.float64, device="cuda")
    timeit(lambda: _lib.check(lib.pk_gram(ctx.handle, 0, n, ld, _ptr(U), nu, _ptr(V), nv, _ptr(g))), 8.0 * n * (2 * k + 3), f"gram MrR k={k} ({2*k+3} vectors)")
    del U, V
