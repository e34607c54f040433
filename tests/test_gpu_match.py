"""GPU parity of the descriptor-matching core (tracking::SearchFrom3D inner loop) vs the numpy oracle: exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_match_hamming_exact(ctx, oracle):
    import textslam_b200 as T
    rng = np.random.default_rng(81)
    nt, nq = 1500, 700
    train = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    query = train[rng.integers(0, nt, nq)].copy()
    flips = rng.integers(0, 256, (nq, 32), dtype=np.uint8) & (rng.random((nq, 32)) < 0.15).astype(np.uint8) * 0xFF
    query ^= (flips & rng.integers(0, 256, (nq, 32), dtype=np.uint8))
    sizes = rng.integers(0, 60, nq); sizes[:5] = 0; sizes[5] = 1; sizes[6] = 400
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    idx = rng.integers(0, nt, ptr[-1]).astype(np.int32)
    idx[ptr[7]:ptr[8]] = idx[ptr[7]] if sizes[7] else 0          # duplicates -> distance ties: first candidate must win
    bi, bd, sd = T.match_hamming(ctx, query, train, ptr, idx)
    oi, od, os_ = oracle.match_hamming(query, train, ptr, idx)
    assert np.array_equal(bd, od) and np.array_equal(sd, os_)
    assert np.array_equal(bi, oi)
    assert np.all(bi[:5] == -1)
