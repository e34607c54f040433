"""CPU: the SearchFrom3D restatement (oracle/pyoracle.py) against an independent brute force over all keypoints (no grid), and the
grid built by the product's host mirror (vectorised) against the oracle's loop version."""
import numpy as np
import textslam_b200 as T
from search3d_cases import make_case


def test_grid_mirror_equals_loop_version(oracle):
    c = make_case(3)
    g = oracle.frame_grid(c["kp_xy"], c["width"], c["height"])
    G = T.FrameGrid(c["kp_xy"], c["width"], c["height"])
    assert G.inv_w == g["inv_w"] and G.inv_h == g["inv_h"]
    for ix in range(G.cols):
        for iy in range(G.rows):
            cid = ix * G.rows + iy
            assert list(G.cell_idx[G.cell_ptr[cid]:G.cell_ptr[cid + 1]]) == g["cells"][ix][iy]
    assert G.cell_ptr[-1] < len(c["kp_xy"])      # some keypoints fall off the grid (PosInGrid false)


def test_oracle_equals_brute_force(oracle):
    c = make_case(4, n_kp=600, n_pts=200)
    g = oracle.frame_grid(c["kp_xy"], c["width"], c["height"])
    th = 15
    bi, bd, uv = oracle.search_from_3d(c["Tcw"], c["K"], c["pt_ray"], c["pt_rho"], c["poses"], c["pt_host"], c["pt_query"], c["query_desc"],
                                       c["kp_xy"], c["kp_octave"], c["train_desc"], g, th)
    r = np.float32(th) * np.float32(1.2)
    on_grid = np.zeros(len(c["kp_xy"]), bool)
    for col in g["cells"]:
        for cell in col:
            on_grid[cell] = True
    n_found = 0
    for i in range(len(bi)):
        if c["pt_query"][i] < 0:
            assert bi[i] == -1
            continue
        u, v = uv[i]
        if u < 0 or u > c["width"] or v < 0 or v > c["height"]:
            assert bi[i] == -1
            continue
        x, y = np.float32(u), np.float32(v)
        ok = on_grid & (c["kp_octave"] <= 1) & (np.abs(c["kp_xy"][:, 0] - x) < r) & (np.abs(c["kp_xy"][:, 1] - y) < r)
        if not ok.any():
            assert bi[i] == -1
            continue
        d = np.unpackbits(c["query_desc"][c["pt_query"][i]][None] ^ c["train_desc"][ok], axis=1).sum(1)
        assert bd[i] == d.min()                                   # the window is a subset of the visited cells: same minimum
        assert d[list(np.nonzero(ok)[0]).index(bi[i])] == d.min()
        n_found += 1
    assert n_found > 50
    m32, m23, n = T.resolve_matches(bi, bd, c["pt_query"], len(c["kp_xy"]), len(c["query_desc"]))
    assert n == (m32 >= 0).sum() == (m23 >= 0).sum() and n > 10
    assert len(set(m32[m32 >= 0])) == n


def test_add_variant_bookkeeping():
    """SearchFrom3DAdd (src/tracking.cc:1196-1270): points matched by the first search are skipped, its 2D->3D table stays in force"""
    bi = np.array([3, 3, 5, 7, -1]); bd = np.array([10, 20, 30, 200, 2147483647]); pq = np.array([0, 1, 2, 3, 4])
    m32, m23, n = T.resolve_matches(bi, bd, pq, 10, 5)
    assert n == 2 and list(m32) == [3, -1, 5, -1, -1]
    bi2 = np.array([9, 4, 8, 6, 3]); bd2 = np.array([1, 50, 60, 70, 80])
    a32, a23, n2 = T.resolve_matches(bi2, bd2, pq, 10, 5, m32=m32, m23=m23)
    assert n2 == 2 and list(a32) == [3, 4, 5, 6, -1] and a23[3] == 0     # point 4 wants key point 3, which point 0 holds
