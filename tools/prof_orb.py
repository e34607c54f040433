"""ORB extractor: device time per 64-image batch with and without the second-stream overlap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
imgs = synth.orb_images(seed=0, n=64)
orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
orb.dev_bench(imgs, reps=3)
ms, nkp = orb.dev_bench(imgs, reps=20)
print("overlap", os.environ.get("TSLAM_ORB_OVERLAP", "1"), "ms/batch", round(ms, 4), "kpts", nkp)
