"""Pin the oracle to the reference's only numeric golden vector.

tests/golden/reference_tests_output.txt is a verbatim copy of the reference's
tests/output.txt (the printed output of tests/test_sparse_solvers.f90).  The
oracle has to reproduce every digit: residual histories of BiCGStab(DILU), ICCG
and DPCG on the two 5x5 systems and the three solution vectors.
"""
import os
import re

import numpy as np

from oracle import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_tests_output.txt")

# matrix literals are default-real (single precision) in the Fortran source and get
# promoted to double on assignment (tests/test_sparse_solvers.f90:605-612, 646-653)
f32 = lambda v: np.array(v, dtype=np.float32).astype(np.float64)

A1 = f32([6.80, -6.05, -0.45, 8.32, -9.67,
          -2.11, -3.30, 2.58, 2.71, -5.14,
          5.66, 5.36, -2.70, 4.35, -7.26,
          5.97, -4.44, 0.27, -7.17, 6.08,
          8.23, 1.08, 9.04, 2.14, -6.87])
B1 = f32([4.02, 6.19, -8.22, -7.57, -3.03])
A2 = f32([3.14, 0.17, -0.90, 1.65, -0.72,
          0.17, 0.79, 0.83, -0.65, 0.28,
          -0.90, 0.83, 4.53, -3.70, 1.60,
          1.65, -0.65, -3.70, 5.32, -1.37,
          -0.72, 0.28, 1.60, -1.37, 1.98])
B2 = f32([-7.29, 9.25, 5.99, -1.94, -8.30])


def dense5_csr():
    ja = np.tile(np.arange(1, 6, dtype=np.int32), 5)
    ioffset = np.array([1, 6, 11, 16, 21, 26], dtype=np.int32)
    diag = np.array([1, 7, 13, 19, 25], dtype=np.int32)
    return oracle.Csr(ioffset, ja, diag)


def parse_golden():
    blocks, cur = [], None
    for line in open(GOLD):
        m = re.search(r"res0 =\s*(\S+)", line)
        if m:
            cur = {"res0": m.group(1), "iters": [], "sol": []}
            blocks.append(cur)
            continue
        m = re.search(r"iter =\s*(\d+) resl =\s*(\S+) rsm =\s*(\S+)", line)
        if m:
            cur["iters"].append((int(m.group(1)), m.group(2), m.group(3)))
            continue
        m = re.match(r"\s+(-?\d+\.\d\d)\s+(-?\d+\.\d\d)\s*$", line)
        if m and cur is not None:
            cur["sol"].append(m.group(1))
    return blocks


def fmt(x):  # Fortran 1PE10.3
    return f"{x:10.3E}".strip()


def run_case(name, a, b, x):
    res0, resl, iters, res, hist = oracle.solve(name, dense5_csr(), a, b, x, sor=float(np.float32(1e-13)),
                                                nsw=100, small=oracle.SMALL_TEST, tol=-1.0, history=True)
    return res0, hist, iters


def test_reference_golden_output_reproduced_digit_for_digit():
    gold = parse_golden()
    assert len(gold) == 3
    x = np.zeros(5)
    got = []
    for name, a, b in (("bicgstab", A1, B1), ("iccg", A2, B2), ("dpcg", A2, B2)):
        res0, hist, iters = run_case(name, a, b, x)   # x is chained exactly like the Fortran program
        got.append((res0, hist, iters, x.copy()))
    for g, (res0, hist, iters, sol) in zip(gold, got):
        assert fmt(res0) == g["res0"]
        assert iters == len(g["iters"])
        for (it, resl_s, rsm_s), resl in zip(g["iters"], hist):
            assert fmt(resl) == resl_s, (it, fmt(resl), resl_s)
            rsm = resl / (res0 + oracle.SMALL_TEST)
            assert fmt(rsm) == rsm_s, (it, fmt(rsm), rsm_s)
        assert [f"{v:5.2f}".strip() for v in sol] == g["sol"]


def test_iteration_counts_match_baseline_md():
    gold = parse_golden()
    assert [len(g["iters"]) for g in gold] == [6, 5, 5]
