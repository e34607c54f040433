"""Small driver for ncu: C2 ORB batch (64 x 640x480) through the device-resident bench hook."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = T.Context(0)
imgs = synth.orb_images(seed=0, n=n)
orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
ms, nkp = orb.dev_bench(imgs, reps=3)
print(json.dumps({"ms_per_batch": ms, "kpts": nkp, "kpts_per_s": nkp / ms * 1e3, "images_per_s": n / ms * 1e3}))
