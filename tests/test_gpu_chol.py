"""Reduced-system solver alone (the linear solve inside ceres::Solve, src/optimizer.cc:1222,1602,1840): dense SPD systems with
different tile patterns through tslam_dev_chol_solve, fused persistent kernel (default path of tslam_solve) and wave kernels,
against numpy.linalg.solve."""
import ctypes as C

import numpy as np
import pytest

from textslam_b200._lib import lib, check
from chol_cases import NB, nd_pattern, spd_with_pattern, dev_chol_solve

pytestmark = pytest.mark.gpu


CASES = ["one_tile", "two_tiles_dense", "dense_5", "nd3", "nd4_c5_shape", "random_sparse", "ragged_n", "dense_12"]


def make_case(case):
    rng = np.random.default_rng(abs(hash(case)) % 1000 + 7)
    if case == "one_tile":
        n, pat = 42, np.ones((1, 1), bool)
    elif case == "two_tiles_dense":
        n, pat = 100, np.tril(np.ones((2, 2), bool))
    elif case == "dense_5":
        n, pat = 5 * NB, np.tril(np.ones((5, 5), bool))
    elif case == "dense_12":
        n, pat = 12 * NB - 5, np.tril(np.ones((12, 12), bool))
    elif case == "nd3":
        pat = nd_pattern(3); n = pat.shape[0] * NB
    elif case == "nd4_c5_shape":
        pat = nd_pattern(4); n = pat.shape[0] * NB - 11
    elif case == "random_sparse":
        Tn = 14
        pat = np.tril(rng.random((Tn, Tn)) < 0.18)
        for j in range(Tn - 1):
            pat[j + 1, j] |= (j % 3 != 2)
        n = Tn * NB
    else:
        Tn = 7
        pat = np.tril(rng.random((Tn, Tn)) < 0.4); n = Tn * NB - 37
    S = spd_with_pattern(rng, n, pat)
    return S, rng.standard_normal(n), pat


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("case", CASES)
def test_chol_solve_matches_numpy(ctx, case, mode):
    S, b, pat = make_case(case)
    ref = np.linalg.solve(S, b)
    x, ms, info, _ = dev_chol_solve(ctx, S, b, pat, mode, reps=2)
    assert info[3] == 0
    assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (case, mode, np.abs(x - ref).max())


def test_pattern_from_nonzeros_and_ill_conditioned(ctx):
    """tile pattern derived from S itself; a Jacobi-scaled LM-like system (unit diagonal + damping, correlation close to 1)"""
    rng = np.random.default_rng(5)
    n = 300
    J = rng.standard_normal((n + 20, n))
    S = J.T @ J
    d = 1.0 / np.sqrt(np.diag(S)); S = S * d[:, None] * d[None, :] + 1e-4 * np.eye(n)
    b = rng.standard_normal(n)
    ref = np.linalg.solve(S, b)
    for mode in (1, 0):
        x, _, info, _ = dev_chol_solve(ctx, S, b, None, mode)
        assert info[3] == 0
        assert np.abs(x - ref).max() <= 1e-7 * np.abs(ref).max(), mode


def test_not_positive_definite_raises_fail_flag(ctx):
    rng = np.random.default_rng(6)
    S, b, pat = make_case("dense_5")
    S[70, 70] = -1.0
    for mode in (1, 0):
        _, _, info, _ = dev_chol_solve(ctx, S, b, pat, mode)
        assert info[3] == 1, mode
