"""GPU parity of tool::CalTextinfo (mu / sigma of projected text quads) vs the oracle."""
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_text_info_matches_oracle(ctx, oracle):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=91, n=3)
    rng = np.random.default_rng(91)
    quads, qimg = [], []
    for k in range(300):
        if k % 3 == 2:   # partially outside the image
            q = np.stack([rng.uniform(-150, 790, 4), rng.uniform(-120, 600, 4)], 1)
        else:            # text-box-like convex quad inside the image
            c = np.array([rng.uniform(60, 580), rng.uniform(40, 440)]); a = rng.uniform(0, np.pi); hw, hh = rng.uniform(8, 55), rng.uniform(4, 25)
            R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            q = c + (np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) * rng.uniform(0.8, 1.2, (4, 1))) @ R.T
        quads.append(q); qimg.append(k % 3)
    quads[0] = np.array([[700.0, 500.0], [710.0, 500.0], [710.0, 510.0], [700.0, 510.0]])   # entirely outside -> empty
    ok, mu, sg = T.text_info(ctx, imgs, np.array(quads), qimg)
    for k in range(300):
        oko, muo, sgo = oracle.text_info(imgs[qimg[k]], quads[k])
        assert bool(ok[k]) == oko, k
        if oko:
            assert mu[k] == muo, (k, mu[k], muo)
            assert abs(sg[k] - sgo) <= 1e-12 * sgo, (k, sg[k], sgo)
    assert not ok[0]
