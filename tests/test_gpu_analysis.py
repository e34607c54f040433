"""The device-side structure analysis (csrc/analysis_dev.cu) must produce the same index structures as the host analysis."""
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,maker", [
    ("c3_pose_only", synth.c3_pose_only),
    ("c4_local_ba", synth.c4_local_ba),
    ("c5_global_ba", synth.c5_global_ba),
    ("ext_fixed", lambda: synth.make_ba_problem(seed=5, n_kf=40, n_lm=600, obs_per_lm=4, n_planes=6, feats_per_plane=9, n_ext=3, frac_ext_lm=0.2)),
    ("mid_nd", lambda: synth.make_ba_problem(seed=9, n_kf=150, n_lm=4000, obs_per_lm=4, n_planes=10, feats_per_plane=4, fixed_cams=(0,))),
    ("no_text", lambda: synth.make_ba_problem(seed=2, n_kf=12, n_lm=300, obs_per_lm=3)),
])
def test_device_analysis_equals_host_analysis(ctx, name, maker):
    prob = maker()
    assert ctx.compare_analysis(prob) == []


def test_device_analysis_with_shuffled_observations(ctx):
    """Observation order is arbitrary in the reference (AddResidualBlock order): the index structures must still agree."""
    prob = synth.make_ba_problem(seed=11, n_kf=30, n_lm=800, obs_per_lm=4, n_planes=5, feats_per_plane=4)
    rng = np.random.default_rng(0)
    perm = rng.permutation(prob.n_pobs)
    for f in ("p_uv", "p_ray", "p_cam", "p_host", "p_lm"):
        setattr(prob, f, np.ascontiguousarray(getattr(prob, f)[perm]))
    assert ctx.compare_analysis(prob) == []
