"""GPU parity of the direct-method frame pyramid (frame::GetPyrMat) vs the oracle: bit-exact."""
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_frame_pyramid_bit_exact(ctx, oracle):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=71, n=3)
    fp = T.FramePyramid(ctx, 8)
    fp.build(imgs)
    for i in range(3):
        for l in range(8):
            for what in (0, 1, 2, 3):
                assert np.array_equal(fp.get(i, l, what), oracle.frame_pyramid(imgs[i], l, what)), (i, l, what)
    odd = synth.orb_images(seed=72, n=1, w=645, h=487)   # odd sizes: ((w+1)/2, (h+1)/2) chain
    fp.build(odd)
    for l in range(8):
        for what in (0, 1, 2, 3):
            assert np.array_equal(fp.get(0, l, what), oracle.frame_pyramid(odd[0], l, what)), (l, what)
    fp.close()
