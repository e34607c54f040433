"""Multi-GPU global BA check (run under torchrun, one rank per GPU): the landmark-sharded solve with one NCCL
all-reduce per LM iteration must reproduce the single-GPU solve (same iteration sequence, parameters to 1e-8)."""
import os, sys, json, time
os.environ["TSLAM_SMALL"] = "0"   # the sharded solve runs the general path: the single-GPU reference takes the same one
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import textslam_b200 as T
from textslam_b200 import synth
from textslam_b200.dist import env_rank_world, broadcast_unique_id

rank, world, local = env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.Context(local)
ctx.init_comm(rank, world, broadcast_unique_id(T.Context.nccl_unique_id, rank, dist))
ref_ctx = T.Context(local)   # plain single-GPU context on the same device
ok = True
for name, prob, its in (("c4+text", synth.c4_local_ba(seed=51), 10), ("medium", synth.c5_global_ba(seed=52, n_kf=60, n_lm=3000, n_planes=20, text_kf_stride=2), 8),
                        ("c5", synth.c5_global_ba(seed=0), 20)):
    a, b = prob.copy(), prob.copy()
    t0 = time.perf_counter(); sm, frm, trm = ctx.solve(a, its); tm = time.perf_counter() - t0
    t0 = time.perf_counter(); s1, fr1, tr1 = ref_ctx.solve(b, its); t1 = time.perf_counter() - t0
    rel = lambda x, y: float(np.abs(x - y).max() / (np.abs(y).max() + 1e-300)) if y.size else 0.0
    res = {"case": name, "rank": rank, "iters": (sm["iterations"], s1["iterations"]), "cost": (sm["final_cost"], s1["final_cost"]),
           "rel_cams": rel(a.cams, b.cams), "rel_rho": rel(a.rho, b.rho), "rel_theta": rel(a.theta, b.theta),
           "rel_resid": rel(frm, fr1), "ms_multi": 1e3 * tm, "ms_single": 1e3 * t1}
    good = sm["iterations"] == s1["iterations"] and res["rel_cams"] < 1e-8 and res["rel_rho"] < 1e-8 and res["rel_theta"] < 1e-8 and res["rel_resid"] < 1e-7
    ok &= good
    if rank == 0 or not good:
        print(json.dumps(res), flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU_CHECK", "PASS" if flag.item() == 1.0 else "FAIL", "world", world, flush=True)
dist.destroy_process_group()
