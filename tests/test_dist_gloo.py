"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: rendezvous helpers and the landmark sharding
rule — every observation has exactly one owner, all observations of a free landmark are co-located, and the
per-rank partial costs (evaluated with the CPU oracle) all-reduce to the full cost."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from textslam_b200 import synth, dist as tdist
    from textslam_b200._abi import PT_BA, TX_BA, JAC_ANALYTIC
    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert tdist.env_rank_world() == (rank, world, rank)
        uid = tdist.broadcast_unique_id(lambda: bytes(range(128)), rank, dist)
        assert uid == bytes(range(128))
        assert tdist.max_over_ranks(10.0 + rank, dist) == 10.0 + world - 1
        prob = synth.make_ba_problem(seed=41, n_kf=8, n_lm=300, obs_per_lm=3, band=8, fixed_cams=(0, 1), n_ext=2, frac_ext_lm=0.2, n_planes=6)
        ip, it = tdist.shard_indices(prob, rank, world)
        # ownership is a partition
        mask = torch.zeros(prob.n_pobs + prob.n_tobs, dtype=torch.int32)
        mask[torch.from_numpy(ip)] += 1
        mask[prob.n_pobs + torch.from_numpy(it)] += 1
        dist.all_reduce(mask)
        assert bool((mask == 1).all())
        # observations of a free landmark never straddle ranks
        owned_lm = set(prob.p_lm[ip][prob.rho_fixed[prob.p_lm[ip]] == 0].tolist())
        gathered = [None] * world
        dist.all_gather_object(gathered, owned_lm)
        assert not (gathered[0] & gathered[1])
        # partial costs sum to the full cost
        r, _ = po.eval_points(prob, PT_BA, want_J=False)
        rt, _ = po.eval_text(prob, TX_BA, JAC_ANALYTIC, want_J=False)
        part = torch.tensor([(r[ip] ** 2).sum() + (rt[it] ** 2).sum()], dtype=torch.float64)
        dist.all_reduce(part)
        full = (r ** 2).sum() + (rt ** 2).sum()
        assert abs(part.item() - full) <= 1e-9 * full
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
