"""GPU parity of the ORB extractor (through the C-ABI) against the CPU oracle: bit-exact keypoints
(coordinates, size, angle, response, octave, order) and descriptors, per image of a batch."""
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu


def _check(kp, desc, kpo, desco, tag=""):
    assert len(kp) == len(kpo), (tag, len(kp), len(kpo))
    for f in ("octave", "response", "x", "y", "size", "angle", "class_id"):
        assert np.array_equal(kp[f], kpo[f]), (tag, f, np.nonzero(kp[f] != kpo[f])[0][:5])
    assert np.array_equal(desc, desco), (tag, np.nonzero((desc != desco).any(1))[0][:5])


def test_pyramid_levels_bit_exact(ctx, oracle):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=31, n=2)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    orb.extract_batch(imgs)
    for i in range(2):
        for l in range(8):
            assert orb.level_size(l) == oracle.orb_level_size(640, 480, 1.2, 8, l)
            assert np.array_equal(orb.get_level(i, l), oracle.orb_pyramid_level(imgs[i], 1.2, 8, l)), (i, l)
    orb.close()


def test_batch_bit_exact_vs_oracle(ctx, oracle):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=32, n=6)
    imgs[4][:] = 77                                   # flat image: no keypoints at all
    imgs[5][:] = 90; imgs[5][200:230, 300:340] = 200  # one rectangle: a handful of corners (threshold fallback, tiny quad-tree)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    res = orb.extract_batch(imgs)
    for i, (kp, desc) in enumerate(res):
        kpo, desco = oracle.orb_extract(imgs[i], 1000, 1.2, 8, 20, 7)
        _check(kp, desc, kpo, desco, f"img{i}")
    assert len(res[4][0]) == 0 and 0 < len(res[5][0]) < 200
    orb.close()


def test_operator_call_and_3000_features(ctx, oracle):
    import textslam_b200 as T
    img = synth.orb_images(seed=33, n=1)[0]
    orb = T.ORBextractor(ctx, 3000, 1.2, 8, 20, 7)   # the extractor used for the first frames (src/tracking.cc:39)
    kp, desc = orb(img)
    kpo, desco = oracle.orb_extract(img, 3000, 1.2, 8, 20, 7)
    _check(kp, desc, kpo, desco, "3000")
    orb.close()


def test_other_sizes_and_blur_variant(ctx, oracle):
    import textslam_b200 as T
    img = synth.orb_images(seed=34, n=1, w=752, h=480)[0]
    orb = T.ORBextractor(ctx, 500, 1.2, 6, 20, 7, blur_variant=1)
    kp, desc = orb(img)
    kpo, desco = oracle.orb_extract(img, 500, 1.2, 6, 20, 7, blur_variant=1)
    _check(kp, desc, kpo, desco, "752x480")
    orb.close()


def test_real_texture_like_image(ctx, oracle):
    """Smooth gradients + sparse structure: most cells fall back to the low threshold or stay empty."""
    import textslam_b200 as T
    rng = np.random.default_rng(35)
    yy, xx = np.mgrid[0:480, 0:640]
    img = (96 + 40 * np.sin(xx / 37.0) + 30 * np.cos(yy / 23.0)).astype(np.float32)
    for _ in range(40):
        x0, y0 = int(rng.integers(0, 600)), int(rng.integers(0, 440))
        img[y0:y0 + int(rng.integers(5, 40)), x0:x0 + int(rng.integers(5, 40))] += float(rng.uniform(-60, 60))
    img = np.clip(img + rng.normal(0, 2.0, img.shape), 0, 255).astype(np.uint8)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    kp, desc = orb(img)
    kpo, desco = oracle.orb_extract(img, 1000, 1.2, 8, 20, 7)
    _check(kp, desc, kpo, desco, "texture")
    assert len(kp) > 50
    orb.close()


def test_dev_bench_hook(ctx):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=36, n=4)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    ms, nkp = orb.dev_bench(imgs, reps=2)
    assert ms > 0 and nkp >= 4 * 1000
    orb.close()


def test_batch_64_bit_exact_vs_oracle(ctx, oracle):
    """BASELINE configuration C2 at its full batch: 64 images of 640x480 in one call, every keypoint field and descriptor byte
    equal to the oracle's per-image extraction."""
    import textslam_b200 as T
    from concurrent.futures import ThreadPoolExecutor
    imgs = synth.orb_images(seed=37, n=64)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    res = orb.extract_batch(imgs)
    with ThreadPoolExecutor(8) as ex:   # the oracle is a ctypes call that releases the GIL
        ref = list(ex.map(lambda im: oracle.orb_extract(im, 1000, 1.2, 8, 20, 7), imgs))
    assert len(res) == 64
    for i, ((kp, desc), (kpo, desco)) in enumerate(zip(res, ref)):
        _check(kp, desc, kpo, desco, f"img{i}")
    orb.close()


def test_row_stride_wider_than_the_image(ctx, oracle):
    """cv::Mat rows need not be packed (step > cols: a region of interest, an aligned allocation): the extractor takes the row
    stride in bytes (tslam_orb_extract's `stride`) and must read exactly the w columns of every row."""
    import textslam_b200 as T
    packed = synth.orb_images(seed=38, n=3)
    n, h, w = packed.shape
    wide = np.full((n, h, w + 96), 255, np.uint8)   # the padding holds a different value than any border handling would produce
    wide[:, :, :w] = packed
    view = wide[:, :, :w]
    assert view.strides[1] == w + 96 and not view.flags["C_CONTIGUOUS"]
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    res = orb.extract_batch(view)
    assert orb._stride == w + 96
    for i, (kp, desc) in enumerate(res):
        kpo, desco = oracle.orb_extract(packed[i], 1000, 1.2, 8, 20, 7)
        _check(kp, desc, kpo, desco, f"img{i}")
    orb.close()
