"""Fused reduced-system solve: the task queue built on the host (analysis.cpp: chol_fused_schedule, chol_sched.hpp) is replayed
in numpy, strictly in queue order. Every task must find its dependency counters satisfied by EARLIER tasks (the queue is a valid
topological order: the device kernel never waits for something that is popped later) and the replay must reproduce
numpy.linalg.solve. Runs on the CPU: the schedule export touches no device."""
import numpy as np
import pytest

from chol_cases import NB, HB, FT_F, FT_S, FT_U, FT_B, get_schedule, replay, spd_with_pattern, nd_pattern


@pytest.mark.parametrize("case", ["one_tile", "two_tiles_dense", "dense_5", "nd3", "nd4_c5_shape", "random_sparse", "ragged_n"])
def test_schedule_replay_matches_numpy(case):
    rng = np.random.default_rng(hash(case) % 1000)
    if case == "one_tile":
        n, pat = 42, np.ones((1, 1), bool)
    elif case == "two_tiles_dense":
        n, pat = 100, np.tril(np.ones((2, 2), bool))
    elif case == "dense_5":
        n, pat = 5 * NB, np.tril(np.ones((5, 5), bool))
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
    b = rng.standard_normal(n)
    x, n_by_type, Tn = replay(S, b, pat)
    ref = np.linalg.solve(S, b)
    assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    assert n_by_type[FT_B] == Tn and n_by_type[FT_F] >= 1


def test_c5_shape_schedule_pairs_every_node():
    """the tile-aligned nested-dissection layout of the 500-keyframe global BA: 16 leaves + 15 separators of two tiles each ->
    31 F tasks (every node is a pair), the forward solve rides on the b row (one S task per tile)"""
    pat = nd_pattern(4)
    tasks, deps, srcs, below, nsync, Tn = get_schedule(pat.shape[0] * NB, pat)
    f = tasks[tasks[:, 0] == FT_F]
    assert Tn == 62 and len(f) == 31 and (f[:, 6] == 2).all()
    s = tasks[tasks[:, 0] == FT_S]
    assert (s[s[:, 6] == Tn][:, 8] == 8).all() and len(s[s[:, 6] == Tn]) == Tn
    # the queue is sorted by need: the first task is a leaf factorisation, the last ones are the backward solve
    assert tasks[0, 0] == FT_F and (tasks[-Tn:, 0] == FT_B).all()
