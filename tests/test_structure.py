"""Host-side structure analysis (csrc/analysis.cpp) against a brute-force numpy/set restatement; runs without a GPU."""
import numpy as np
import pytest
import textslam_b200 as T
from textslam_b200 import synth


def brute(prob):
    cf = np.asarray(prob.cam_fixed, bool)
    rf = np.asarray(prob.rho_fixed, bool) if prob.rho_fixed is not None else np.zeros(len(prob.rho), bool)
    tf = np.asarray(prob.theta_fixed, bool) if prob.theta_fixed is not None else np.zeros(len(prob.theta), bool)
    free_cams, free_pts, free_pl = set(), set(), set()
    blocks, slots_p, slots_t = set(), set(), set()
    lm_cams = {}
    n_direct = 0

    def visit(c, h, l, lf, tag, slots, free_l):
        nonlocal n_direct
        if cf[c] and cf[h] and lf:
            return
        cams = [k for k in (c, h) if not cf[k]]
        free_cams.update(cams)
        for k in cams:
            blocks.add((k, k)); n_direct += 1
        if len(cams) == 2 and c != h:
            blocks.add((min(c, h), max(c, h))); n_direct += 1
        if not lf:
            free_l.add(l)
            for k in cams:
                slots.add((l, k))
            lm_cams.setdefault((tag, l), set()).update(cams)

    for c, h, l in zip(prob.p_cam, prob.p_host, prob.p_lm):
        visit(int(c), int(h), int(l), rf[l], 0, slots_p, free_pts)
    for c, h, l in zip(prob.t_cam, prob.t_host, prob.t_plane):
        visit(int(c), int(h), int(l), tf[l], 1, slots_t, free_pl)
    n_schur = 0
    for cams in lm_cams.values():
        cs = sorted(cams)
        n_schur += len(cs) * (len(cs) + 1) // 2
        for i, a in enumerate(cs):
            for b in cs[i:]:
                blocks.add((a, b))
    return dict(n_free_cams=len(free_cams), n_free_points=len(free_pts), n_free_planes=len(free_pl), n_blocks=len(blocks),
                n_slots_point=len(slots_p), n_slots_text=len(slots_t), n_schur_entries=n_schur, n_direct_entries=n_direct)


@pytest.mark.parametrize("maker", [synth.c3_pose_only, synth.c4_local_ba,
                                   lambda: synth.make_ba_problem(seed=5, n_kf=40, n_lm=600, obs_per_lm=4, n_planes=6, feats_per_plane=9, n_ext=3, frac_ext_lm=0.2)])
def test_structure_counts_match_brute_force(maker):
    prob = maker()
    info = T.analyze_structure(prob)
    ref = brute(prob)
    for k, v in ref.items():
        assert info[k] == v, (k, info[k], v)
    assert info["reduced_dim"] == 6 * ref["n_free_cams"]
    assert info["n_tiles"] == (info["reduced_dim"] + 63) // 64


def test_structure_sharding_is_a_partition():
    prob = synth.make_ba_problem(seed=9, n_kf=150, n_lm=4000, obs_per_lm=4, n_planes=10, feats_per_plane=4, fixed_cams=(0,))
    whole = T.analyze_structure(prob)
    assert whole["n_waves"] < whole["n_tiles"]          # the nested-dissection order shortens the tile elimination chain
    for world in (2, 3):
        parts = [T.analyze_structure(prob, r, world) for r in range(world)]
        for p in parts:   # every rank builds the same reduced system
            for k in ("n_free_cams", "n_free_points", "n_free_planes", "n_blocks", "n_tiles", "n_waves", "n_tile_updates"):
                assert p[k] == whole[k], (k, world)
        for k in ("n_local_pobs", "n_local_tobs", "n_owned_points", "n_owned_planes", "n_slots_point", "n_slots_text", "n_schur_entries", "n_direct_entries"):
            assert sum(p[k] for p in parts) == whole[k], (k, world)


def test_structure_rejects_bad_indices():
    prob = synth.c4_local_ba()
    prob.p_lm = prob.p_lm.copy(); prob.p_lm[3] = len(prob.rho) + 7
    with pytest.raises(T.TslamError):
        T.analyze_structure(prob)


def test_banded_problems_get_a_shallow_tile_elimination_tree():
    """nd_layout.h: on a banded keyframe graph the tile-aligned nested-dissection layout must turn the chain of T panel steps into
    L * tiles(separator) + tiles(leaf) waves (C5 shape: 498 free cameras, band 20 -> 4 levels of 2-tile separators + 2-tile leaves)."""
    import textslam_b200 as T
    info = T.analyze_structure(synth.c5_global_ba(seed=0))
    assert info["n_free_cams"] == 498 and info["reduced_dim"] == 2988
    assert info["n_tiles"] == 62 and info["n_waves"] == 10     # 15 separators x 2 tiles + 16 leaves x 2 tiles; natural order: 47 waves
    small = T.analyze_structure(synth.c5_global_ba(seed=3, n_kf=100, n_lm=3000))   # < 128 free cameras: natural order, no padding
    assert small["n_tiles"] == (6 * small["n_free_cams"] + 63) // 64
    big = T.analyze_structure(synth.make_ba_problem(seed=8, n_kf=1000, n_lm=12000, obs_per_lm=4, band=10, fixed_cams=(0, 1), w_point=1.0))
    assert big["n_waves"] <= 12 < big["n_tiles"]
    # sharded analysis sees the same global structure on every rank
    a, b = T.analyze_structure(synth.c5_global_ba(seed=0), rank=0, world=2), T.analyze_structure(synth.c5_global_ba(seed=0), rank=1, world=2)
    for k in ("n_free_cams", "n_blocks", "n_tiles", "n_waves", "n_tile_updates"):
        assert a[k] == b[k] == info[k]
