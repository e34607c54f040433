#!/usr/bin/env python
"""bench.py — BASELINE.json metric on the C5 global-BA configuration (500 KF x 100k point observations).

One "step" = one GlobalBA call (src/optimizer.cc:334-453: level 0, max 20 LM iterations) on the
seeded synthetic C5 problem. `value` = residual+Jacobian block evaluations per second over whole
solves with the problem resident in HBM; `ms_per_step` = one solve; `lm_iter_ms` is reported beside
it. `e2e` = the same metric through the public C-ABI call `tslam_solve` with HOST buffers (upload,
structure analysis, LM loop, download inside the timed region). `--impl reference` times the CPU
oracle (Ceres-faithful restatement; the reference's Ceres/OpenCV path cannot be built here, see
DESIGN.md) on the box's host cores: the same whole solve per step, all host threads (num_threads = 1 beside it).

Launch: python bench.py [--gpus N --steps K --warmup W] or, for N > 1, under torchrun (one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BA_ITERS = 20  # src/optimizer.cc:411-414
POINT_EVAL_BYTES = 268  # 44 B in + 224 B out per auto_BAScene(NW) evaluation
TEXT_EVAL_BYTES = 1280  # 256 B in + 64 B r + 960 B J per nume_BAText block
ORB_IMAGE_BYTES = 307200 + 1158012 + 1000 * (28 + 32)   # image in + pyramid + keypoints/descriptors out
FP64_TENSOR_PEAK = 37.0e12   # measured DMMA rate of this part (tools/ubench/fp64_rates.cu, profiles/r1_notes.md); MEASURED_PEAKS.json has no FP64 entry
WORKLOAD = "C5 global BA: 500 KF x 100k auto_BASceneNW obs (25k landmarks x 4 obs, band +-10, text off as src/optimizer.cc:1707), <=20 LM its"
# ncu --set full captures of this round, exported with `ncu -i ... --page raw --csv` (tools/gpu_prof2.sh); dram traffic is read from them
NCU_RAW = {"point_eval_x16": "profiles/r2_ncu_point_eval_x16_raw.csv", "text_eval": "profiles/r2_ncu_text_eval_raw.csv"}


def bench_config(world):
    """The `config` object of both arms (ours and --impl reference): identical by construction."""
    return {"workload": WORKLOAD,
            "l2": "the solve is not L2-flushed between iterations (its working set, 21 MB of J + 8 MB of factor tiles + index lists, is what a real solve keeps in L2); the stand-alone kernel timings of `roofline` flush L2 or exceed it",
            "landmark_sharding": f"landmark % {world}" if world > 1 else "none"}


def ncu_dram_bytes(key, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch of a kernel in a committed ncu raw-page CSV; None if absent."""
    import csv
    path = os.path.join(ROOT, NCU_RAW[key])
    try:
        with open(path, newline="") as f:
            rows = list(csv.reader(f))
    except OSError:
        return None
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        return None
    units = rows[rows.index(hdr) + 1]
    try:
        kn, cr, cw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[rows.index(hdr) + 2:]:
        if len(r) > max(kn, cr, cw) and kernel_substr in r[kn]:
            try:
                return float(r[cr].replace(",", "")) * scale.get(units[cr], 1.0) + float(r[cw].replace(",", "")) * scale.get(units[cw], 1.0)
            except ValueError:
                return None
    return None


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_running(self, timeout=3.0):
        """Blocks until the sampler has produced its first line (nvidia-smi takes ~0.1 s to start)."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return True
            time.sleep(0.01)
        return False

    def stop(self, t_begin=None, t_end=None):
        """Samples inside [t_begin, t_end] (time.time() of the timed region) are used; if the region was shorter than the
        sampling period, all samples taken since the sampler started (warm-up + timed region, GPU under load) are."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                row = [float(c[0]), float(c[1]), c[3:7], None]
            except ValueError:
                continue
            if len(c) >= 8:
                try:
                    row[3] = datetime.datetime.strptime(c[7], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    pass
            rows.append(row)
        inside = [r for r in rows if r[3] is not None and t_begin is not None and t_begin - 0.02 <= r[3] <= t_end + 0.02]
        used = inside if inside else rows
        reasons = set()
        for r in used:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if used:
            out.update(sm_mhz=float(np.median([r[0] for r in used])), sm_max_mhz=float(max(r[1] for r in used)), reasons=sorted(reasons), samples=len(used),
                       window="timed region" if inside else "warm-up + timed region (region shorter than the sampling period)")
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


class NvmlSampler:
    """SM clock + clock-event reasons sampled through NVML from a thread of this process every ~2 ms (the timed region of
    a default run is only ~0.1-0.3 s long, too short for `nvidia-smi -lms`). The solver calls run inside ctypes with the
    GIL released, so the thread keeps sampling while the GPU is busy."""
    NAMES = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index=0, period=0.002):
        import threading
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        # NVML enumerates physical devices; map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        phys = int(ids[index]) if ids and all(v.strip().isdigit() for v in ids) and index < len(ids) else index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.rows, self.period, self._stop = [], period, threading.Event()
        self._sample()   # fails here (-> nvidia-smi fallback) if the queries are unsupported
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((time.time(), mhz, mask))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period)

    def wait_running(self, timeout=3.0):
        return True

    def stop(self, t_begin=None, t_end=None):
        self._stop.set()
        self.t.join(timeout=2)
        inside = [r for r in self.rows if t_begin is not None and t_begin <= r[0] <= t_end]
        used = inside if inside else self.rows
        reasons = set()
        for _, _, mask in used:
            for name, attr in self.NAMES:
                if mask & int(getattr(self.nv, attr, 0)):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median([r[1] for r in used])) if used else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
               "samples": len(used), "source": "nvml", "window": "timed region" if inside else "warm-up + timed region"}
        try:
            self.nv.nvmlShutdown()
        except Exception:
            pass
        return out


def make_clock_sampler(index):
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)


PROBLEM_ARRAYS = ("cams", "cam_fixed", "rho", "rho_fixed", "theta", "theta_fixed", "p_uv", "p_ray", "p_cam", "p_host", "p_lm",
                  "t_rays", "t_iref", "t_musigma", "t_cam", "t_host", "t_plane", "t_img", "imgs")


def pinned_copy(prob, torch):
    """A copy of the host problem whose arrays live in page-locked memory (what a SLAM front end that reuses its
    staging buffers would pass to tslam_solve)."""
    q = prob.copy()
    q._pins = []
    for name in PROBLEM_ARRAYS:
        a = getattr(q, name)
        if a.size:
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            q._pins.append(t)
            setattr(q, name, t.numpy())
    return q


def problem_bytes(p):
    arrs = [p.cams, p.cam_fixed, p.rho, p.rho_fixed, p.theta, p.theta_fixed, p.p_uv, p.p_ray, p.p_cam, p.p_host, p.p_lm,
            p.t_rays, p.t_iref, p.t_musigma, p.t_cam, p.t_host, p.t_plane, p.t_img, p.imgs]
    return int(sum(a.nbytes for a in arrs))


def jac_evals(summ, prob):
    return (summ["successful_steps"] + 1) * (prob.n_pobs + prob.n_tobs)


def oracle_full_solve(po, prob, threads):
    q = prob.copy()
    t0 = time.perf_counter()
    summ, _, _ = po.solve(q, GLOBAL_BA_ITERS, n_threads=threads, want_trace=False)
    return time.perf_counter() - t0, summ


def run_reference(args, rank, world):
    """CPU arm: the reference's own CPU implementation of the path is Ceres, which cannot be built here, so this
    times the oracle restatement (Ceres-faithful LM loop, oracle/ba_lm.cpp) on the box's host cores. One step = the SAME whole
    GlobalBA solve (<= 20 LM iterations) on the same problem as our arm, with all host threads; the reference's own setting,
    num_threads = 1 (src/optimizer.cc:1838), is timed once beside it."""
    if rank != 0:
        return
    from textslam_b200 import synth
    from oracle import pyoracle as po
    po.build()
    cores = os.cpu_count() or 1
    threads = min(cores, 32)
    prob = synth.c5_global_ba(seed=0)
    times, evals, its = [], [], []
    for s in range(args.warmup + args.steps):
        dt, summ = oracle_full_solve(po, prob, threads)
        if s >= args.warmup:
            times.append(dt); evals.append(jac_evals(summ, prob)); its.append(summ["iterations"])
    total = sum(times)
    value = sum(evals) / total / 1e6
    dt1, summ1 = oracle_full_solve(po, prob, 1)
    sample = f"whole C5 solve per step ({its[0]} LM iterations, the same solve our arm runs), {threads} host threads"
    line = {"impl": "reference", "metric": "ba_resjac_mevals_per_s", "value": value, "unit": "M-evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.gpus),
            "lm_iter_ms": 1e3 * total / max(1, sum(its)), "lm_iterations_per_step": sum(its) / len(its),
            "cpu_baseline": {"value": value, "unit": "M-evals/s", "cores": threads, "kind": "port", "sample": sample,
                             "single_thread": {"note": "num_threads = 1 as src/optimizer.cc:1838 sets it, one whole solve", "value": jac_evals(summ1, prob) / dt1 / 1e6,
                                               "ms_per_solve": 1e3 * dt1, "lm_iter_ms": 1e3 * dt1 / max(1, summ1["iterations"])}},
            "e2e": {"value": value, "unit": "M-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the eval-kernel / ORB / CPU side measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import textslam_b200 as T
    from textslam_b200 import synth
    from textslam_b200._lib import lib
    assert torch.cuda.is_available(), "bench.py needs a B200; textslam_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = T.Context(local_rank)
    if world > 1:
        uid = [T.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.init_comm(rank, world, uid[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prob = synth.c5_global_ba(seed=0)
    n_obs = prob.n_pobs + prob.n_tobs
    dev = ctx.upload(prob)
    W, K = max(args.warmup, 3), args.steps
    clocks = make_clock_sampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.wait_running()
    for _ in range(W):
        dev.lm_iterations(GLOBAL_BA_ITERS)
    launches0 = lib().tslam_launch_count()
    barrier()
    t_region0 = time.time()
    t_wall0 = time.perf_counter()
    dev_ms, evals, its = 0.0, 0, 0
    phases_acc = np.zeros(8)
    for _ in range(K):
        phases, summ = dev.lm_iterations(GLOBAL_BA_ITERS)
        dev_ms += phases[7] * max(1, summ["iterations"])  # whole-call device time (CUDA events on the library stream)
        phases_acc += np.array(phases) * max(1, summ["iterations"])
        evals += jac_evals(summ, prob); its += summ["iterations"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    launches = lib().tslam_launch_count() - launches0
    clk = clocks.stop(t_region0, time.time()) if clocks else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = evals / (dev_ms * 1e-3) / 1e6
    hbm_peak, peak_src = read_peaks()

    line = {"metric": "ba_resjac_mevals_per_s", "value": value, "unit": "M-evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(world),
            "lm_iter_ms": dev_ms / max(1, its), "lm_iterations_per_step": its / K, "wall_ms_per_step": wall_ms / K,
            "gpu_launches": int(launches), "clocks": clk}
    phase_names = ["eval_resjac", "landmark_prep", "reduced_build", "allreduce", "cholesky", "backsub", "model_candidate"]
    line["lm_phase_ms_per_iter"] = {n: float(phases_acc[i] / max(1, its)) for i, n in enumerate(phase_names)}

    if rank == 0 and not args.no_extras and world == 1:
        # ---- the residual+Jacobian kernel alone (north_star roofline target), inputs resident in HBM.
        # Primary roofline: a launch whose footprint (429 MB) exceeds L2 (C5 x16 replica, same structure); the C5-sized
        # launch (26.8 MB, L2 flushed with a write+read sweep of 2x L2 between launches) is reported beside it.
        big = synth.make_ba_problem(seed=1, n_kf=500, n_lm=400000, obs_per_lm=4, band=10, fixed_cams=(0, 1), w_point=1.0, perturb=False)
        dbig = ctx.upload(big)
        dbig.eval_points(T.PT_BA_NW, reps=3)
        msb = dbig.eval_points(T.PT_BA_NW, reps=10, flush_l2=True)
        achb = POINT_EVAL_BYTES * big.n_pobs / (msb * 1e-3) / 1e9
        line["roofline"] = {"kernel": "point_eval_kernel<0,13,J> (auto_BASceneNW residual+Jacobian), C5 x16 replica = 1.6M evals / launch (429 MB > L2)",
                            "bound": "hbm", "achieved": achb, "peak": hbm_peak, "unit": "GB/s", "frac": achb / hbm_peak,
                            "traffic": ncu_dram_bytes("point_eval_x16", "point_eval_kernel"), "traffic_source": NCU_RAW["point_eval_x16"],
                            "peak_source": peak_src, "ms_per_launch": msb,
                            "mevals_per_s": big.n_pobs / (msb * 1e-3) / 1e6,
                            "algorithmic_bytes_per_launch": POINT_EVAL_BYTES * big.n_pobs}
        dbig.free()
        dev.eval_points(T.PT_BA_NW, reps=5, flush_l2=True)
        ms = dev.eval_points(T.PT_BA_NW, reps=20, flush_l2=True)
        ach = POINT_EVAL_BYTES * prob.n_pobs / (ms * 1e-3) / 1e9
        line["roofline_c5"] = {"kernel": "same kernel at the C5 size (100k evals, 26.8 MB / launch, L2 flushed between launches)", "bound": "hbm",
                               "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "ms_per_launch": ms,
                               "mevals_per_s": prob.n_pobs / (ms * 1e-3) / 1e6,
                               "note": "one partial wave (782 CTAs on 148 SMs): launch/latency bound, see DESIGN.md"}
        info = T.analyze_structure(prob)
        chol_ms = line["lm_phase_ms_per_iter"]["cholesky"]
        chol_flops = 2.0 * 64 ** 3 * info["n_tile_updates"]   # the factor's GEMM work (64^3 tile multiply-adds of the symbolic factorisation)
        line["dominant_kernel"] = {"name": "chol_fused_kernel (persistent, flag-ordered: tile Cholesky of the reduced camera system + forward and backward solve in one launch)",
                                   "share_of_lm_iteration": chol_ms / line["lm_iter_ms"],
                                   "ms_per_launch": chol_ms, "tile_updates": int(info["n_tile_updates"]), "flops": chol_flops,
                                   "achieved_tflops": chol_flops / (chol_ms * 1e-3) / 1e12, "peak_tflops": FP64_TENSOR_PEAK / 1e12,
                                   "frac": chol_flops / (chol_ms * 1e-3) / FP64_TENSOR_PEAK,
                                   "bound": "latency: 5 elimination-tree levels x 128 dependent pivots (reciprocal -> fma -> shuffle, ~190 cycles each in "
                                            "potrf32_sym, tools/ubench/potrf_sym_bench.cu) + 3 flag hand-offs per level; the FP64 tensor work (mma.sync.m8n8k4.f64) is "
                                            "a few percent of the pipe's capacity by construction"}
        # ---- CPU baseline beside it: the oracle port on the same whole solve (all host threads; the reference's num_threads = 1 beside it)
        from oracle import pyoracle as po
        po.build()
        threads = min(os.cpu_count() or 1, 32)
        dt, summ = oracle_full_solve(po, prob, threads)
        dt1, summ1 = oracle_full_solve(po, prob, 1)
        line["cpu_baseline"] = {"value": jac_evals(summ, prob) / dt / 1e6, "unit": "M-evals/s", "cores": threads, "kind": "port",
                                "sample": f"one whole C5 solve ({summ['iterations']} LM iterations, oracle/ba_lm.cpp), {threads} host threads",
                                "ms_per_solve": 1e3 * dt, "lm_iter_ms": 1e3 * dt / max(1, summ["iterations"]),
                                "single_thread": {"note": "num_threads = 1 as src/optimizer.cc:1838 sets it, one whole solve", "value": jac_evals(summ1, prob) / dt1 / 1e6,
                                                  "ms_per_solve": 1e3 * dt1, "lm_iter_ms": 1e3 * dt1 / max(1, summ1["iterations"])}}
        line["text_on"] = text_on_bench(ctx, T, synth, po, threads, hbm_peak, peak_src)
        line["small_problems"] = small_problem_bench(ctx, T, synth, po, threads)
        try:
            line["orb"] = orb_bench(ctx, T, synth, po, hbm_peak)
        except Exception as e:  # ORB is reported beside the BA metric; its absence must not hide the BA line
            line["orb"] = {"error": str(e)[:200]}
    dev.free()
    # ---- end to end through the public C-ABI with HOST buffers, at every N: all ranks call tslam_solve on the same problem
    # (sharded by landmark inside the library when the context has a communicator); wall clock, max over ranks.
    fr_bytes = 8 * (2 * prob.n_pobs + 8 * prob.n_tobs)
    copies = [pinned_copy(prob, torch) for _ in range(2 + K)]
    fr_host = torch.empty(2 * prob.n_pobs + 8 * prob.n_tobs, dtype=torch.float64).pin_memory().numpy()
    e2e_evals = 0
    for s in range(2 + K):
        if s == 2:
            barrier()
            t0 = time.perf_counter()
        summ, _, _ = ctx.solve(copies[s], GLOBAL_BA_ITERS, want_trace=False, final_residuals=fr_host)
        if s >= 2:
            e2e_evals += jac_evals(summ, prob)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_t = float(t.item())
    line["e2e"] = {"value": e2e_evals / e2e_t / 1e6, "unit": "M-evals/s", "h2d_bytes_per_step": problem_bytes(prob),
                   "d2h_bytes_per_step": int(prob.cams.nbytes + prob.rho.nbytes + prob.theta.nbytes + fr_bytes),
                   "ms_per_step": 1e3 * e2e_t / K,
                   "call": "tslam_solve (caller-owned page-locked host buffers: index validation, upload, structure analysis, LM loop, "
                           "download of parameters + final residuals); every rank passes the whole problem, the library shards it"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def small_problem_bench(ctx, T, synth, po, threads):
    """C3 pose-only and C4 local BA (BASELINE.json configs 2-3): one pyramid level = one tslam_solve call with HOST buffers
    (10 LM iterations allowed), through the persistent small-problem kernel (csrc/ba_small.cu) and, beside it, through the general
    path (TSLAM_SMALL=0) and the single-threaded oracle; then the reference's three-level PoseOptim loop with its chi^2 gates
    (src/optimizer.cc:172-186) end to end."""
    import torch
    from textslam_b200.api import PyramidLevel, run_pyramid
    out = {}

    def timed(prob_of, n=20):
        cps = [prob_of() for _ in range(n + 3)]
        for c in cps[:3]:
            summ, _, _ = ctx.solve(c, 10, want_trace=False)
        t0 = time.perf_counter(); inner = 0.0
        for c in cps[3:]:
            summ, _, _ = ctx.solve(c, 10, want_trace=False)
            inner += summ["total_ms"]
        return 1e3 * (time.perf_counter() - t0) / n, inner / n, summ

    for name, prob in (("c3_pose_only_2k_pts_250_text", synth.c3_pose_only(seed=0)), ("c4_local_ba_10kf_3k_pts_750_text", synth.c4_local_ba(seed=0))):
        wall, inner, summ = timed(prob.copy)
        wall_pin, inner_pin, _ = timed(lambda: pinned_copy(prob, torch))
        os.environ["TSLAM_SMALL"] = "0"
        try:
            wall_gen, inner_gen, sgen = timed(prob.copy)
        finally:
            os.environ.pop("TSLAM_SMALL", None)
        t0 = time.perf_counter()
        so, _, _ = po.solve(prob.copy(), 10, n_threads=1, want_trace=False)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        out[name] = {"gpu_ms_per_solve_e2e": wall, "gpu_ms_per_solve_inside_the_c_abi": inner, "gpu_ms_per_solve_e2e_pinned_host_buffers": wall_pin,
                     "general_path_ms_per_solve_e2e": wall_gen, "lm_iterations": summ["iterations"], "general_path_lm_iterations": sgen["iterations"],
                     "oracle_1thread_ms_per_solve": cpu_ms, "oracle_iterations": so["iterations"],
                     "note": "e2e = wall clock around Context.solve (ctypes marshalling included); kernel launches per solve: 1"}

    # three-level PoseOptim (levels 2, 1, 0; gates 12.25 / 0.5, 0.5, 0.95; 10 iterations each)
    def levels():
        lv = []
        for l in (2, 1, 0):
            p = synth.c3_pose_only(seed=5, level=l)
            lv.append(PyramidLevel(p, t_obj=p.t_plane.copy(), t_feat=np.tile(np.arange(25), len(p.theta))))
        return lv

    def flags(lv):
        n_obj = len(lv[0].prob.theta)
        return np.ones(lv[0].prob.n_pobs, bool), np.ones(n_obj, bool), np.ones((n_obj, 25), bool)

    def oracle_gated(prob, gate, t_obj, obj_size, max_iters):
        summ, fr, tr = po.solve(prob, max_iters, n_threads=1, want_trace=False)
        pb, tb, ob, cnt = po.gate_residuals(fr, prob.n_pobs, prob.n_tobs, gate, t_obj, obj_size)
        return summ, fr, tr, pb, tb, ob, cnt

    opt = T.Optimizer(ctx)
    for _ in range(3):
        lv = levels(); opt.PoseOptim(lv, 10, *flags(lv))
    sets = [levels() for _ in range(10)]
    t0 = time.perf_counter()
    for lv in sets:
        res = opt.PoseOptim(lv, 10, *flags(lv))
    gpu_ms = 1e3 * (time.perf_counter() - t0) / len(sets)
    lv = levels()
    t0 = time.perf_counter()
    ro = run_pyramid(oracle_gated, lv, (12.25,) * 3, (0.5, 0.5, 0.95), (10,) * 3, *flags(lv))
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    out["pose_optim_3_levels"] = {"workload": "optimizer::PoseOptim: levels 2, 1, 0 of C3 (2000 point + 250 text blocks each), chi^2 gates between the levels",
                                  "gpu_ms_e2e": gpu_ms, "oracle_1thread_ms": cpu_ms, "lm_iterations": [r["summary"]["iterations"] for r in res],
                                  "oracle_lm_iterations": [r["summary"]["iterations"] for r in ro],
                                  "note": "includes the Python level loop (sub-problem assembly in numpy) on both sides"}
    return out


def text_on_bench(ctx, T, synth, po, threads, hbm_peak, peak_src):
    """C5 with the text branch of PyrGlobalBA switched on (src/optimizer.cc:1766-1822, w_T = 1; BASELINE.md reports C5 twice):
    1000 planes x 25 features = 25 000 nume_BAText blocks beside the 100k point blocks; the text kernel's own roofline beside it."""
    prob = synth.c5_global_ba(seed=0, n_planes=1000)
    dev = ctx.upload(prob)
    for _ in range(2):
        dev.lm_iterations(GLOBAL_BA_ITERS)
    ms_tot, evals, its = 0.0, 0, 0
    for _ in range(5):
        phases, summ = dev.lm_iterations(GLOBAL_BA_ITERS)
        ms_tot += phases[7] * max(1, summ["iterations"]); evals += jac_evals(summ, prob); its += summ["iterations"]
    out = {"workload": "C5 + text: 500 KF x (100k auto_BASceneNW + 25k nume_BAText blocks, 1000 planes), analytic text Jacobian",
           "value": evals / (ms_tot * 1e-3) / 1e6, "unit": "M-evals/s", "ms_per_step": ms_tot / 5, "lm_iter_ms": ms_tot / max(1, its), "lm_iterations_per_step": its / 5}
    dev.eval_text(T.TX_BA, T.JAC_ANALYTIC, reps=3, flush_l2=True)
    ms = dev.eval_text(T.TX_BA, T.JAC_ANALYTIC, reps=20, flush_l2=True)
    ach = TEXT_EVAL_BYTES * prob.n_tobs / (ms * 1e-3) / 1e9
    out["roofline_text_eval"] = {"kernel": "text_eval_kernel<0,15,analytic> (nume_BAText residual + 8x15 Jacobian), 25k blocks = 32 MB / launch, L2 flushed",
                                 "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "ms_per_launch": ms,
                                 "traffic": ncu_dram_bytes("text_eval", "text_eval_kernel"), "traffic_source": NCU_RAW["text_eval"], "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": TEXT_EVAL_BYTES * prob.n_tobs,
                                 "note": "one partial wave at this size (latency bound like the point kernel at the C5 size)"}
    dev.eval_text(T.TX_BA, T.JAC_ANALYTIC_TMA, reps=3, flush_l2=True)
    ms_tma = dev.eval_text(T.TX_BA, T.JAC_ANALYTIC_TMA, reps=20, flush_l2=True)
    out["tma_staged_variant"] = {"kernel": "text_eval_tma_kernel (one CTA per text object and keyframe; image window by cp.async.bulk.tensor.2d, parameter blocks in shared memory)",
                                 "ms_per_launch": ms_tma, "achieved": TEXT_EVAL_BYTES * prob.n_tobs / (ms_tma * 1e-3) / 1e9, "unit": "GB/s",
                                 "vs_ldg_taps": ms / ms_tma}
    ms_cd = dev.eval_text(T.TX_BA, T.JAC_CENTRAL_DIFF, reps=5, flush_l2=True)
    out["central_diff_ms_per_launch"] = ms_cd   # Ceres-faithful mode: 35 exact functor evaluations per block
    dev.free()
    dt, summ = oracle_full_solve(po, prob, threads)
    out["cpu_baseline"] = {"value": jac_evals(summ, prob) / dt / 1e6, "unit": "M-evals/s", "cores": threads, "kind": "port",
                           "sample": f"one whole solve ({summ['iterations']} LM iterations), analytic text Jacobian", "ms_per_solve": 1e3 * dt}
    return out


def orb_bench(ctx, T, synth, po, hbm_peak):
    imgs = synth.orb_images(seed=0, n=64)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    ms, nkp = orb.dev_bench(imgs, reps=5)
    t0 = time.perf_counter()
    res = orb.extract_batch(imgs)
    dt = time.perf_counter() - t0
    n = sum(len(k) for k, _ in res)
    out = {"workload": "C2: 64 x 640x480, 8 levels x1.2, 1000 feat, FAST 20/7", "kpts_per_s": nkp / (ms * 1e-3), "ms_per_batch": ms,
           "images_per_s": 64 / (ms * 1e-3), "e2e_kpts_per_s": n / dt, "keypoints": int(nkp)}
    ach = ORB_IMAGE_BYTES * 64 / (ms * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm (nominal; what binds is the integer ALU work of the FAST measure and the per-level launch latency of the small pyramid levels)",
                       "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "algorithmic_bytes_per_image": ORB_IMAGE_BYTES}
    # CPU baselines: the C++ restatement of ORBextractor.cc (1 core, then one image per host thread), and OpenCV's own cv::ORB for scale
    t0 = time.perf_counter(); k1 = 0
    for i in range(8):
        kp, _ = po.orb_extract(imgs[i], 1000, 1.2, 8, 20, 7); k1 += len(kp)
    dt1 = time.perf_counter() - t0
    from concurrent.futures import ThreadPoolExecutor
    threads = min(os.cpu_count() or 1, 32)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        ka = sum(len(kp) for kp, _ in ex.map(lambda im: po.orb_extract(im, 1000, 1.2, 8, 20, 7), list(imgs)))
    dta = time.perf_counter() - t0
    out["cpu_baseline"] = {"kind": "port", "unit": "kpts/s", "oracle_1core": k1 / dt1, "oracle_1core_ms_per_image": 1e3 * dt1 / 8,
                           "oracle_all_cores": ka / dta, "cores": threads, "sample": "8 images on 1 core; the 64-image batch over a thread pool"}
    try:
        import cv2
        cv2.setNumThreads(1)
        cvo = cv2.ORB_create(nfeatures=1000, scaleFactor=1.2, nlevels=8, fastThreshold=20)
        t0 = time.perf_counter(); kc = 0
        for i in range(16):
            kp, _ = cvo.detectAndCompute(imgs[i], None); kc += len(kp)
        dtc = time.perf_counter() - t0
        out["cpu_baseline"]["cv2_orb_1core"] = kc / dtc
        out["cpu_baseline"]["cv2_note"] = f"cv2 {cv2.__version__} cv::ORB (Harris ranking, no quad-tree: a different selection rule, listed for scale only)"
    except Exception as e:
        out["cpu_baseline"]["cv2_orb_1core"] = None
        out["cpu_baseline"]["cv2_note"] = str(e)[:100]
    return out


if __name__ == "__main__":
    main()
