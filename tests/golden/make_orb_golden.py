"""Generates tests/golden/orb_cv2_golden.npz with Python cv2 (run in the build container; cv2 4.13.0).
The vectors pin the oracle's restatement of the OpenCV primitives the reference calls
(src/ORBextractor.cc:810-817 FAST, :1131 resize, :1097 GaussianBlur, :103 fastAtan2)."""
import os, sys
import numpy as np
import cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from textslam_b200 import synth

img = synth.orb_images(seed=5, n=1, w=320, h=240)[0]
out = {"cv2_version": np.array(cv2.__version__), "img": img}
lvl1 = cv2.resize(img, (267, 200), interpolation=cv2.INTER_LINEAR)
lvl2 = cv2.resize(lvl1, (222, 167), interpolation=cv2.INTER_LINEAR)
out["resize_267x200"] = lvl1
out["resize_222x167"] = lvl2
out["blur7"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
for t in (20, 7):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    roi = img[30:30 + 41, 50:50 + 38]  # a cell-sized ROI like ComputeKeyPointsOctTree's
    out[f"fast{t}_full"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(img)], dtype=np.int32).reshape(-1, 3)
    out[f"fast{t}_roi"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(np.ascontiguousarray(roi))], dtype=np.int32).reshape(-1, 3)
rng = np.random.default_rng(0)
yx = rng.integers(-200000, 200000, (2000, 2)).astype(np.float32)
yx[:4] = [[0, 0], [0, 5], [5, 0], [-3, -3]]
out["atan2_in"] = yx
out["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], dtype=np.float32)
vals = np.concatenate([np.arange(-8, 8) + 0.5, rng.normal(0, 30, 200)])
out["round_in"] = vals
out["round_out"] = np.array([int(np.rint(v)) for v in vals], dtype=np.int32)  # cvRound == rint (round-half-even)
small = img[:120, :160].copy()
pd = cv2.pyrDown(small)
out["n2_src"] = small
out["n2_pyrdown"] = pd
out["n2_sobel_x"] = cv2.Sobel(small, cv2.CV_8U, 1, 0, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
out["n2_sobel_y"] = cv2.Sobel(small, cv2.CV_8U, 0, 1, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
out["n2_grad"] = cv2.addWeighted(out["n2_sobel_x"], 0.5, out["n2_sobel_y"], 0.5, 0)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_cv2_golden.npz"), **out)
print("wrote", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
