"""Register / spill budget of the hot kernels, read from the ptxas -v log of the in-tree build (CPU test).

The persistent grids are sized `SMs x resident blocks`; a kernel that silently grows past its register budget loses a
resident block per SM and a quarter of its bandwidth (r02: a `__launch_bounds__(256, 1)` on the plain SpMV made ptxas
spend 101 registers instead of 64 and cost 28 %).  This pins the budgets the measured numbers were taken with."""
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LOGDIR = os.path.join(HERE, "..", "parallel_krylov_b200", "csrc", "build")


def _resources(log):
    path = os.path.join(LOGDIR, log)
    if not os.path.exists(path):
        pytest.skip(f"{log} not found (library not built in-tree)")
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores.*?"
                         r"Used (\d+) registers", txt, flags=re.S):
        out[m.group(1)] = (int(m.group(4)), int(m.group(3)))
    return out


def _find(res, pattern):
    hits = [(k, v) for k, v in res.items() if re.search(pattern, k)]
    assert hits, pattern
    return hits


def test_spmv_register_budgets():
    res = _resources("pk_spmv.cu.log")
    # plain single-vector SpMV, 256-row tiles: 64 registers, no spills -> 4 blocks / SM
    for stages in (2, 3, 4):
        for name, (regs, spill) in _find(res, rf"k_spmv_tmaILi1ELi256ELi{stages}ELb0ELi0E"):
            assert regs <= 64 and spill == 0, (name, regs, spill)
    # fused-exchange variant of the same kernel must also fit 4 blocks / SM
    for name, (regs, spill) in _find(res, r"k_spmv_tmaILi1ELi256ELi2ELb1ELi0E"):
        assert regs <= 64, (name, regs, spill)
    # two-chain and fused-step forms: 3 blocks / SM (<= 85 registers), plain and fused-exchange
    for pat in (r"k_spmv_tmaILi2ELi256ELi2ELb[01]ELi0E", r"k_spmv_tmaILi1ELi256ELi2ELb[01]ELi[12]E"):
        for name, (regs, spill) in _find(res, pat):
            assert regs <= 85, (name, regs, spill)


def test_vector_kernel_register_budgets():
    res = _resources("pk_kernels.cu.log")
    for pat, cap in ((r"k_cg_xr", 40), (r"k_cg_pE", 32), (r"k_mrr_update", 64)):
        for name, (regs, spill) in _find(res, pat):
            assert regs <= cap, (name, regs, spill)


def test_matrix_powers_register_budgets():
    res = _resources("pk_matpow.cu.log")
    # dense-band kernel: 384 threads keep two rows of 27 values each in registers (108 registers of values); it must
    # neither spill nor exceed 65536 / 384 = 170 registers, or the block does not launch
    for name, (regs, spill) in _find(res, r"k_matpow_bandILi13E"):
        assert regs <= 168 and spill == 0, (name, regs, spill)
    for name, (regs, spill) in _find(res, r"k_matpow_bandILi0E"):
        assert regs <= 168, (name, regs, spill)
    # general kernel: 640 threads, one row in registers: <= 102 registers (65536 / 640)
    for name, (regs, spill) in _find(res, r"8k_matpowILb[01]E"):
        assert regs <= 96, (name, regs, spill)


def test_dense_band_row_pointer_closed_form_matches_the_generator():
    """csrc/pk_matpow.cu::mb_rowptr (the dense-band kernels never read A's row pointers: they follow from (n, bw)) —
    the same formula restated here against the row pointers of problems.banded_spd, incl. the clipped end rows."""
    import numpy as np
    from parallel_krylov_b200 import problems

    def mb_rowptr(r, n, bw):
        q = r * (2 * bw + 1)
        t = min(r, bw)
        q -= t * bw - t * (t - 1) // 2
        m = r - (n - bw)
        if m > 0:
            q -= m * (m + 1) // 2
        return q

    for n, bw in ((27, 13), (28, 13), (100, 1), (101, 2), (1000, 13), (777, 6)):
        rowptr = np.asarray(problems.banded_spd(n, bw, 0)[0], dtype=np.int64)
        got = np.array([mb_rowptr(r, n, bw) for r in range(n + 1)], dtype=np.int64)
        assert np.array_equal(got, rowptr), (n, bw)
        assert rowptr[-1] == n * (2 * bw + 1) - bw * (bw + 1)           # the count mpi/_dist.py::_setup_band_ext tests
