"""One-process-per-GPU plumbing for multi-GPU global BA. torch.distributed carries the
rendezvous (NCCL on GPUs, gloo in the CPU tests); the data-path collective itself is the library's own
ncclAllReduce inside tslam_solve."""
import os
import numpy as np
from ._lib import lib, check


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_owner(landmark_is_free, landmark_index, obs_index, world):
    r = lib().tslam_shard_owner(int(bool(landmark_is_free)), int(landmark_index), int(obs_index), int(world))
    if r < 0:
        check(r)
    return r


def shard_indices(prob, rank, world):
    """Indices of the point / text observations rank `rank` owns (mirror of the sharded upload)."""
    p = [i for i in range(prob.n_pobs) if shard_owner(not prob.rho_fixed[prob.p_lm[i]], prob.p_lm[i], i, world) == rank]
    t = [i for i in range(prob.n_tobs) if shard_owner(not prob.theta_fixed[prob.t_plane[i]], prob.t_plane[i], i, world) == rank]
    return np.array(p, dtype=np.int64), np.array(t, dtype=np.int64)


def broadcast_unique_id(make_id, rank, dist):
    """Rank 0 creates the 128-byte NCCL unique id, everybody receives it (any torch.distributed backend)."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def max_over_ranks(value, dist, device="cpu"):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
