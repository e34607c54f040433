"""CPU pinning of the descriptor-matching oracle: the reference's DescriptorDistance (src/tracking.cc:2762-2778, the SWAR popcount over
eight 32-bit words) restated literally in numpy uint32 arithmetic, and OpenCV's own Hamming norm as a third opinion."""
import numpy as np
import pytest


def swar_distance(a, b):
    """tracking::DescriptorDistance as written: v = pa ^ pb; v -= (v >> 1) & 0x55555555; ... ; dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24."""
    pa, pb = a.view(np.uint32).astype(np.uint64), b.view(np.uint32).astype(np.uint64)
    M = np.uint64(0xFFFFFFFF)
    dist = 0
    for i in range(8):
        v = pa[i] ^ pb[i]
        v = (v - ((v >> np.uint64(1)) & np.uint64(0x55555555))) & M
        v = ((v & np.uint64(0x33333333)) + ((v >> np.uint64(2)) & np.uint64(0x33333333))) & M
        dist += int((((((v + (v >> np.uint64(4))) & np.uint64(0xF0F0F0F)) * np.uint64(0x1010101)) & M) >> np.uint64(24)))
    return dist


def test_match_oracle_against_reference_popcount_and_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    train = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    query = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    query[:10] = train[rng.integers(0, 300, 10)]                       # exact matches (distance 0)
    sizes = rng.integers(0, 25, len(query)); sizes[0] = 0; sizes[1] = 1
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    idx = rng.integers(0, 300, ptr[-1]).astype(np.int32)
    bi, bd, sd = oracle.match_hamming(query, train, ptr, idx)
    for q in range(len(query)):
        cand = idx[ptr[q]:ptr[q + 1]]
        if len(cand) == 0:
            assert bi[q] == -1 and bd[q] == 2147483647
            continue
        # the reference's scan: strict '<' keeps the first minimum (src/tracking.cc:1161-1175)
        best, best_d, second = -1, 256 * 8 + 1, 2147483647
        ds = []
        for c in cand:
            d = swar_distance(query[q], train[c])
            assert d == int(cv2.norm(query[q], train[c], cv2.NORM_HAMMING))
            ds.append(d)
            if d < best_d:
                best_d, best = d, int(c)
        assert bi[q] == best and bd[q] == best_d
        if len(cand) > 1:
            k = int(np.argmin(ds))
            assert sd[q] == min(ds[:k] + ds[k + 1:])
