"""ctypes binding of the CPU oracle (oracle/libtslam_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (textslam_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np
from textslam_b200._abi import (BAProblemC, SolveOptionsC, SolveSummaryC, PT_NCOLS, TX_NCOLS, TRACE_COLS,
                                solve_options, c_dp, KP_DTYPE)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtslam_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.tso_solve.restype = C.c_int
        _LIB.tso_eval_points.restype = C.c_int
        _LIB.tso_eval_text.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def eval_points(prob, kind, want_J=True, n_threads=1):
    n, nc = prob.n_pobs, PT_NCOLS[kind]
    r = np.zeros((n, 2)); J = np.zeros((n, 2, nc)) if want_J else None
    pc = prob.as_c()
    lib().tso_eval_points(C.c_int(kind), C.byref(pc), _dp(r), _dp(J) if want_J else None, C.c_int(n_threads))
    return r, J


def eval_text(prob, kind, jac_mode, want_J=True, n_threads=1):
    n, nc = prob.n_tobs, TX_NCOLS[kind]
    r = np.zeros((n, 8)); J = np.zeros((n, 8, nc)) if want_J else None
    pc = prob.as_c()
    lib().tso_eval_text(C.c_int(kind), C.c_int(jac_mode), C.byref(pc), _dp(r), _dp(J) if want_J else None, C.c_int(n_threads))
    return r, J


def solve(prob, max_iters=10, text_jac_mode=0, n_threads=1, dense_full=0, want_trace=True, **kw):
    """Runs the Ceres-faithful LM loop; updates prob parameters in place. Returns (summary dict, final_residuals, trace)."""
    opt = solve_options(max_iters, text_jac_mode, n_threads, dense_full, **kw)
    summ = SolveSummaryC()
    fr = np.zeros(2 * prob.n_pobs + 8 * prob.n_tobs)
    tr = np.full((max_iters + 1, TRACE_COLS), np.nan)
    pc = prob.as_c()
    rc = lib().tso_solve(C.byref(pc), C.byref(opt), C.byref(summ), _dp(fr), _dp(tr) if want_trace else None)
    assert rc == 0
    return summ.as_dict(), fr, tr


def point_ambient(cam, host, rho, ray_xy, uv, K4, w):
    r = np.zeros(2); J = np.zeros((2, 15))
    f = lambda a: _dp(np.ascontiguousarray(a, dtype=np.float64))
    lib().tso_point_ambient(f(cam), f(host), C.c_double(rho), f(ray_xy), f(uv), f(K4), f(w), _dp(r), _dp(J))
    return r, J


def quat_plus(x, d):
    out = np.zeros(4)
    f = lambda a: _dp(np.ascontiguousarray(a, dtype=np.float64))
    lib().tso_quat_plus(f(x), f(d), _dp(out))
    return out


# ---- ORB ----------------------------------------------------------------------------------------
def orb_extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, blur_variant=0, max_kp=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    max_kp = max_kp or (nfeatures + 4 * nlevels + 64)
    kp = np.zeros(max_kp, dtype=KP_DTYPE)
    desc = np.zeros((max_kp, 32), dtype=np.uint8)
    n = lib().tso_orb_extract(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(nfeatures), C.c_float(scale),
                              C.c_int(nlevels), C.c_int(ini_th), C.c_int(min_th), C.c_int(blur_variant), C.c_int(max_kp),
                              kp.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p))
    assert n >= 0, "max_kp too small"
    return kp[:n].copy(), desc[:n].copy()


def orb_level_size(w, h, scale, nlevels, level):
    lw, lh = C.c_int(), C.c_int()
    lib().tso_orb_level_size(C.c_int(w), C.c_int(h), C.c_float(scale), C.c_int(nlevels), C.c_int(level), C.byref(lw), C.byref(lh))
    return lw.value, lh.value


def orb_pyramid_level(img, scale, nlevels, level):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    lw, lh = orb_level_size(w, h, scale, nlevels, level)
    out = np.zeros((lh, lw), dtype=np.uint8)
    lib().tso_orb_pyramid_level(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(w), C.c_float(scale), C.c_int(nlevels),
                                C.c_int(level), out.ctypes.data_as(C.c_void_p))
    return out


def orb_features_per_level(nfeatures, scale, nlevels):
    out = np.zeros(nlevels, dtype=np.int32)
    lib().tso_orb_features_per_level(C.c_int(nfeatures), C.c_float(scale), C.c_int(nlevels), out.ctypes.data_as(C.c_void_p))
    return out


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.zeros((dh, dw), dtype=np.uint8)
    lib().tso_resize_linear(src.ctypes.data_as(C.c_void_p), C.c_int(src.shape[1]), C.c_int(src.shape[0]), out.ctypes.data_as(C.c_void_p),
                            C.c_int(dw), C.c_int(dh))
    return out


def gaussian7(src, variant=0):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.zeros_like(src)
    lib().tso_gaussian7(src.ctypes.data_as(C.c_void_p), C.c_int(src.shape[1]), C.c_int(src.shape[0]), out.ctypes.data_as(C.c_void_p), C.c_int(variant))
    return out


def fast(img, threshold, max_kp=200000):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    xyr = np.zeros((max_kp, 3), dtype=np.int32)
    n = lib().tso_fast(img.ctypes.data_as(C.c_void_p), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(threshold), C.c_int(max_kp),
                       xyr.ctypes.data_as(C.c_void_p))
    assert n <= max_kp
    return xyr[:n].copy()


def fast_atan2(y, x):
    f = lib().tso_fast_atan2
    f.restype = C.c_float
    return f(C.c_float(y), C.c_float(x))


def orb_debug(img, what, level, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    """Intermediate stages of the oracle extractor: 0 FAST measure plane, 1 candidates, 2 quad-tree winners."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    if what == 0:
        lw, lh = orb_level_size(w, h, scale, nlevels, level)
        out = np.zeros((lh, lw), dtype=np.uint8)
        lib().tso_orb_debug(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(nfeatures), C.c_float(scale), C.c_int(nlevels),
                            C.c_int(ini_th), C.c_int(min_th), C.c_int(0), C.c_int(level), out.ctypes.data_as(C.c_void_p), C.c_int(out.size))
        return out
    buf = np.zeros((140000, 3), dtype=np.int32)
    n = lib().tso_orb_debug(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(nfeatures), C.c_float(scale), C.c_int(nlevels),
                            C.c_int(ini_th), C.c_int(min_th), C.c_int(what), C.c_int(level), buf.ctypes.data_as(C.c_void_p), C.c_int(len(buf)))
    return buf[:n].copy()


def frame_pyramid(img, level, what):
    """frame::GetPyrMat restatement: what = 0 level image, 1 grad, 2 grad_x, 3 grad_y."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    lw, lh = C.c_int(), C.c_int()
    lib().tso_frame_pyramid(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(8), C.c_int(level), C.c_int(what), None,
                            C.byref(lw), C.byref(lh))
    out = np.zeros((lh.value, lw.value), dtype=np.uint8)
    lib().tso_frame_pyramid(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(8), C.c_int(level), C.c_int(what),
                            out.ctypes.data_as(C.c_void_p), C.byref(lw), C.byref(lh))
    return out


def match_hamming(query_desc, train_desc, cand_ptr, cand_idx):
    """tracking::SearchFrom3D inner loop (src/tracking.cc:1161-1175) with DescriptorDistance (:2762-2778), restated in numpy:
    sequential scan, strict `dist < bestDist` keeps the FIRST minimum. Returns (best_idx, best_dist, second_dist)."""
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    nq = len(q); INT_MAX = 2147483647
    bi = np.full(nq, -1, dtype=np.int32); bd = np.full(nq, INT_MAX, dtype=np.int32); sd = np.full(nq, INT_MAX, dtype=np.int32)
    for i in range(nq):
        c = np.asarray(cand_idx[cand_ptr[i]:cand_ptr[i + 1]], dtype=np.int64)
        if len(c) == 0:
            continue
        d = np.unpackbits(q[i][None, :] ^ t[c], axis=1).sum(1).astype(np.int64)
        k = int(np.argmin(d))   # first minimum
        bi[i] = c[k]; bd[i] = d[k]
        if len(c) > 1:
            sd[i] = np.delete(d, k).min()
    return bi, bd, sd


def frame_grid(kp_xy, width, height, cols=64, rows=48):
    """frame::frame bounds / cell sizes (src/frame.cc:118-125) and AssignFeaturesToGrid + PosInGrid (:376-406), restated with
    Python loops: mGrid[ix][iy] as lists of keypoint indices in insertion order. float members -> np.float32 arithmetic."""
    f32 = np.float32
    g = {"cols": cols, "rows": rows, "min_x": f32(0.0), "min_y": f32(0.0), "max_x": f32(width), "max_y": f32(height)}
    g["inv_w"] = f32(float(cols) / float(g["max_x"] - g["min_x"])); g["inv_h"] = f32(float(rows) / float(g["max_y"] - g["min_y"]))
    cells = [[[] for _ in range(rows)] for _ in range(cols)]
    rnd = lambda v: int(np.floor(float(v) + 0.5)) if v >= 0 else -int(np.floor(-float(v) + 0.5))   # C round(): half away from zero
    for i, (x, y) in enumerate(np.asarray(kp_xy, dtype=f32).reshape(-1, 2)):
        px = rnd((x - g["min_x"]) * g["inv_w"]); py = rnd((y - g["min_y"]) * g["inv_h"])
        if px < 0 or px >= cols or py < 0 or py >= rows:
            continue
        cells[px][py].append(i)
    g["cells"] = cells
    return g


def _quat_R(q):
    w, x, y, z = (float(v) for v in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def search_from_3d(Tcw, K, pt_ray, pt_rho, poses, pt_host, pt_query, query_desc, kp_xy, kp_octave, train_desc, grid, th, min_level=-1, max_level=1):
    """tracking::SearchFrom3D (src/tracking.cc:1124-1176) per map point, with frame::GetFeaturesInArea (src/frame.cc:415-468) written
    out: projection in double (T_cr = T_cw T_rw^-1 in closed form for rigid transforms; the reference goes through Eigen's general
    4x4 inverse, equal up to rounding), bounds test in double, query window and distance tests in float, cells ix-outer / iy-inner,
    strict `dist < bestDist`. Returns (best_idx, best_dist, uv)."""
    f32 = np.float32
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    kp_xy = np.asarray(kp_xy, dtype=f32).reshape(-1, 2)
    n = len(pt_rho); INT_MAX = 2147483647
    bi = np.full(n, -1, np.int32); bd = np.full(n, INT_MAX, np.int32); uv = np.zeros((n, 2))
    Rc = _quat_R(Tcw[:4]); tc = np.asarray(Tcw[4:], dtype=np.float64)
    fx, fy, cx, cy = (float(v) for v in K)
    r = f32(th) * f32(1.2)
    for i in range(n):
        if pt_query[i] < 0:
            continue
        P = poses[pt_host[i]]
        Rr = _quat_R(P[:4]); tr = P[4:]
        R = np.array([[(Rc[a, 0] * Rr[b, 0] + Rc[a, 1] * Rr[b, 1]) + Rc[a, 2] * Rr[b, 2] for b in range(3)] for a in range(3)])
        ir = 1.0 / float(pt_rho[i])
        p = np.zeros(3)
        for a in range(3):
            Rt = (R[a, 0] * tr[0] + R[a, 1] * tr[1]) + R[a, 2] * tr[2]
            Rray = (R[a, 0] * pt_ray[i][0] + R[a, 1] * pt_ray[i][1]) + R[a, 2]
            p[a] = ir * Rray + (tc[a] + -Rt)
        X = fx * p[0] + cx * p[2]; Y = fy * p[1] + cy * p[2]
        u = X / p[2]; v = Y / p[2]
        uv[i] = (u, v)
        if u < float(grid["min_x"]) or u > float(grid["max_x"]) or v < float(grid["min_y"]) or v > float(grid["max_y"]):
            continue
        x, y = f32(u), f32(v)
        c0 = max(0, int(np.floor((x - grid["min_x"] - r) * grid["inv_w"])))
        if c0 >= grid["cols"]:
            continue
        c1 = min(grid["cols"] - 1, int(np.ceil((x - grid["min_x"] + r) * grid["inv_w"])))
        if c1 < 0:
            continue
        r0 = max(0, int(np.floor((y - grid["min_y"] - r) * grid["inv_h"])))
        if r0 >= grid["rows"]:
            continue
        r1 = min(grid["rows"] - 1, int(np.ceil((y - grid["min_y"] + r) * grid["inv_h"])))
        if r1 < 0:
            continue
        check = min_level > 0 or max_level >= 0
        best, best_k = INT_MAX, -1
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for k in grid["cells"][ix][iy]:
                    if check:
                        if kp_octave[k] < min_level:
                            continue
                        if max_level >= 0 and kp_octave[k] > max_level:
                            continue
                    if abs(kp_xy[k, 0] - x) < r and abs(kp_xy[k, 1] - y) < r:
                        d = int(np.unpackbits(q[pt_query[i]] ^ t[k]).sum())
                        if d < best:
                            best, best_k = d, k
        if best_k >= 0:
            bi[i] = best_k; bd[i] = best
    return bi, bd, uv


def search_in_area(uv, pt_query, query_desc, kp_xy, kp_octave, train_desc, grid, th, kp_skip=None, min_level=-1, max_level=-1):
    """tracking::SearchFrom3DLocalTrack (src/tracking.cc:1296-1329) per map point: GetFeaturesInArea at the given projection (double
    narrowed to float at the call), candidates flagged in kp_skip left out (:1311-1313), best and runner-up distance with the
    reference's update rule. Returns (best_idx, best_dist, second_dist)."""
    f32 = np.float32
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    kp_xy = np.asarray(kp_xy, dtype=f32).reshape(-1, 2)
    n = len(uv); INT_MAX = 2147483647
    bi = np.full(n, -1, np.int32); bd = np.full(n, INT_MAX, np.int32); sd = np.full(n, INT_MAX, np.int32)
    r = f32(th) * f32(1.2)
    for i in range(n):
        if pt_query[i] < 0:
            continue
        x, y = f32(uv[i][0]), f32(uv[i][1])
        c0 = max(0, int(np.floor((x - grid["min_x"] - r) * grid["inv_w"])))
        c1 = min(grid["cols"] - 1, int(np.ceil((x - grid["min_x"] + r) * grid["inv_w"])))
        r0 = max(0, int(np.floor((y - grid["min_y"] - r) * grid["inv_h"])))
        r1 = min(grid["rows"] - 1, int(np.ceil((y - grid["min_y"] + r) * grid["inv_h"])))
        if c0 >= grid["cols"] or c1 < 0 or r0 >= grid["rows"] or r1 < 0:
            continue
        check = min_level > 0 or max_level >= 0
        best, best2, best_k = INT_MAX, INT_MAX, -1
        for ix in range(c0, c1 + 1):
            for iy in range(r0, r1 + 1):
                for k in grid["cells"][ix][iy]:
                    if check:
                        if kp_octave[k] < min_level:
                            continue
                        if max_level >= 0 and kp_octave[k] > max_level:
                            continue
                    if not (abs(kp_xy[k, 0] - x) < r and abs(kp_xy[k, 1] - y) < r):
                        continue
                    if kp_skip is not None and kp_skip[k]:
                        continue
                    d = int(np.unpackbits(q[pt_query[i]] ^ t[k]).sum())
                    if d < best:
                        best2 = best; best = d; best_k = k
                    elif d < best2:
                        best2 = d
        if best_k >= 0:
            bi[i] = best_k; bd[i] = best; sd[i] = best2
    return bi, bd, sd


def gate_residuals(final_residuals, n_pobs, n_tobs, gate, t_obj=None, obj_size=None):
    """The reference's outlier loops (src/optimizer.cc:1236-1302, 1616-1684) on a final residual vector.
    Returns (pt_bad, tf_bad, obj_bad, (nBadS, nBadFeat, nBadT)); raises if the reference's asserts would fire."""
    fr = np.ascontiguousarray(final_residuals, dtype=np.float64)
    t_obj = np.ascontiguousarray(t_obj if t_obj is not None else np.zeros(n_tobs), dtype=np.int32)
    obj_size = np.ascontiguousarray(obj_size if obj_size is not None else [], dtype=np.int32)
    pb, tb, ob = np.zeros(n_pobs, np.uint8), np.zeros(n_tobs, np.uint8), np.zeros(len(obj_size), np.uint8)
    cnt = (C.c_int32 * 3)()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().tso_gate_residuals(_dp(fr), C.c_int(n_pobs), C.c_int(n_tobs), vp(t_obj), vp(obj_size), C.c_int(len(obj_size)),
                                  C.byref(gate), vp(pb), vp(tb), vp(ob), cnt)
    if rc != 0:
        raise ValueError(f"tso_gate_residuals: inconsistent object bookkeeping ({rc})")
    return pb, tb, ob, tuple(cnt)


def theta_covariance(prob, jac_mode=0):
    cov = np.zeros((len(prob.theta), 3, 3))
    pc = prob.as_c()
    ns = lib().tso_theta_covariance(C.byref(pc), C.c_int(jac_mode), _dp(cov))
    return cov, ns


def text_info(img, quad, want_mask=False):
    """tool::CalTextinfo restatement: (ok, mu, sigma[, fillPoly mask]) for one quad (4x2 doubles) on a u8 image."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    q = np.ascontiguousarray(quad, dtype=np.float64).reshape(8)
    mu, sg = C.c_double(), C.c_double()
    mask = np.zeros((h, w), dtype=np.uint8) if want_mask else None
    ok = lib().tso_text_info(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), _dp(q), C.byref(mu), C.byref(sg),
                             mask.ctypes.data_as(C.c_void_p) if want_mask else None)
    return (bool(ok), mu.value, sg.value, mask) if want_mask else (bool(ok), mu.value, sg.value)
