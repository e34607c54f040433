"""Per-phase device time of one LM iteration on C5 with the text branch on (25 000 nume_BAText blocks beside the 100 000 point blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
prob = synth.c5_global_ba(seed=0, n_planes=1000)
dev = ctx.upload(prob)
for _ in range(3):
    dev.lm_iterations(20)
acc = np.zeros(8); n = 5
for _ in range(n):
    ph, s = dev.lm_iterations(20)
    acc += np.asarray(ph)
names = ["eval", "lm_prep", "reduced_build", "allreduce", "cholesky", "backsub", "model_cand", "whole"]
print({k: round(float(v) / n, 4) for k, v in zip(names, acc)}, "its", s["iterations"])
