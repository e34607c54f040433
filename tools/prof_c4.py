import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
prob = synth.c4_local_ba(seed=0)
dev = ctx.upload(prob)
dev.lm_iterations(3)
dev.lm_iterations(3)
