"""Host-side mirror of the reference interface for the hot path, on top of the C-ABI.

Names follow the reference: `Optimizer.PoseOptim / LocalBundleAdjustment / GlobalBA` correspond to
src/optimizer.h:57-70 (here they take the flattened SoA problem that optimizer.cc:213-279 builds out
of the map instead of the pointer graph), `ORBextractor.__call__` to src/ORBextractor.h:51-61.
"""
import ctypes as C
import numpy as np
from ._lib import lib, check, TslamError  # noqa: F401
from ._abi import (BAProblem, SolveSummaryC, KeyPointC, KP_DTYPE, PT_NCOLS, TX_NCOLS, TRACE_COLS, solve_options,
                   c_dp, c_bp, c_ip, PT_BA, PT_BA_NW, PT_POSE, PT_RHO, TX_BA, TX_POSE, TX_THETA, JAC_ANALYTIC,
                   JAC_CENTRAL_DIFF, JAC_ANALYTIC_TMA, GateOptionsC, gate_options)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


class Context:
    """One per host thread / GPU (tslam_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().tslam_ctx_create(C.c_int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().tslam_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU -------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().tslam_nccl_unique_id(buf))
        return bytes(buf)

    def init_comm(self, rank, world, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().tslam_ctx_init_comm(self._h, C.c_int(rank), C.c_int(world), buf))

    # ---- evaluation --------------------------------------------------------------------------
    def eval_points(self, prob, kind, want_J=True):
        n, nc = prob.n_pobs, PT_NCOLS[kind]
        r = np.zeros((n, 2))
        J = np.zeros((n, 2, nc)) if want_J else None
        pc = prob.as_c()
        check(lib().tslam_eval_points(self._h, C.c_int(kind), C.byref(pc), _dp(r), _dp(J)))
        return r, J

    def eval_text(self, prob, kind, jac_mode=JAC_ANALYTIC, want_J=True):
        n, nc = prob.n_tobs, TX_NCOLS[kind]
        r = np.zeros((n, 8))
        J = np.zeros((n, 8, nc)) if want_J else None
        pc = prob.as_c()
        check(lib().tslam_eval_text(self._h, C.c_int(kind), C.c_int(jac_mode), C.byref(pc), _dp(r), _dp(J)))
        return r, J

    # ---- solve ---------------------------------------------------------------------------------
    def solve(self, prob, max_iters=10, text_jac_mode=JAC_ANALYTIC, want_trace=True, final_residuals=None, **kw):
        """ceres::Solve + Problem::Evaluate replacement; updates `prob` parameters in place.
        `final_residuals`: optional caller-owned float64 buffer of 2*n_pobs + 8*n_tobs (e.g. page-locked) to receive the residuals."""
        opt = solve_options(max_iters, text_jac_mode, **kw)
        summ = SolveSummaryC()
        nfr = 2 * prob.n_pobs + 8 * prob.n_tobs
        if final_residuals is not None:
            fr = final_residuals
            assert fr.dtype == np.float64 and fr.size == nfr and fr.flags.c_contiguous
        else:
            fr = np.zeros(nfr)
        tr = np.full((max_iters + 1, TRACE_COLS), np.nan)
        pc = prob.as_c()
        check(lib().tslam_solve(self._h, C.byref(pc), C.byref(opt), C.byref(summ), _dp(fr), _dp(tr) if want_trace else None))
        return summ.as_dict(), fr, tr

    # ---- chi^2 gates (src/optimizer.cc:1236-1302, 1616-1684) ------------------------------------
    @staticmethod
    def _gate_buffers(n_pobs, n_tobs, t_obj, obj_size):
        t_obj = np.ascontiguousarray(t_obj if t_obj is not None else np.zeros(n_tobs), dtype=np.int32)
        obj_size = np.ascontiguousarray(obj_size if obj_size is not None else [], dtype=np.int32)
        assert len(t_obj) == n_tobs
        return t_obj, obj_size, np.zeros(n_pobs, np.uint8), np.zeros(n_tobs, np.uint8), np.zeros(len(obj_size), np.uint8), (C.c_int32 * 3)()

    def gate_residuals(self, final_residuals, n_pobs, n_tobs, gate, t_obj=None, obj_size=None):
        """Outlier flags from the final residual vector. Returns (pt_bad, tf_bad, obj_bad, (nBadS, nBadFeat, nBadT))."""
        fr = np.ascontiguousarray(final_residuals, dtype=np.float64)
        assert fr.size == 2 * n_pobs + 8 * n_tobs
        t_obj, obj_size, pb, tb, ob, cnt = self._gate_buffers(n_pobs, n_tobs, t_obj, obj_size)
        check(lib().tslam_gate_residuals(self._h, _dp(fr), C.c_int(n_pobs), C.c_int(n_tobs), t_obj.ctypes.data_as(c_ip),
                                         obj_size.ctypes.data_as(c_ip), C.c_int(len(obj_size)), C.byref(gate),
                                         pb.ctypes.data_as(c_bp), tb.ctypes.data_as(c_bp), ob.ctypes.data_as(c_bp), cnt))
        return pb, tb, ob, tuple(cnt)

    def solve_gated(self, prob, gate, t_obj=None, obj_size=None, max_iters=10, text_jac_mode=JAC_ANALYTIC, want_trace=True, **kw):
        """One pyramid level of PoseOptim / LocalBundleAdjustment: ceres::Solve + Problem::Evaluate + the chi^2 loops in one call.
        Returns (summary, final_residuals, trace, pt_bad, tf_bad, obj_bad, counts)."""
        opt = solve_options(max_iters, text_jac_mode, **kw)
        summ = SolveSummaryC()
        fr = np.zeros(2 * prob.n_pobs + 8 * prob.n_tobs)
        tr = np.full((max_iters + 1, TRACE_COLS), np.nan)
        t_obj, obj_size, pb, tb, ob, cnt = self._gate_buffers(prob.n_pobs, prob.n_tobs, t_obj, obj_size)
        pc = prob.as_c()
        check(lib().tslam_solve_gated(self._h, C.byref(pc), C.byref(opt), C.byref(gate), t_obj.ctypes.data_as(c_ip),
                                      obj_size.ctypes.data_as(c_ip), C.c_int(len(obj_size)), C.byref(summ), _dp(fr),
                                      _dp(tr) if want_trace else None, pb.ctypes.data_as(c_bp), tb.ctypes.data_as(c_bp),
                                      ob.ctypes.data_as(c_bp), cnt))
        return summ.as_dict(), fr, tr, pb, tb, ob, tuple(cnt)

    def theta_covariance(self, prob, jac_mode=JAC_ANALYTIC):
        """ceres::Covariance of every theta block (src/optimizer.cc:2219-2238). Returns (cov (n_planes,3,3), n_singular)."""
        cov = np.zeros((len(prob.theta), 3, 3)); ns = C.c_int32(0)
        pc = prob.as_c()
        check(lib().tslam_theta_covariance(self._h, C.byref(pc), C.c_int(jac_mode), _dp(cov), C.byref(ns)))
        return cov, ns.value

    def compare_analysis(self, prob):
        """Test hook: names of the index arrays on which the device-side structure analysis differs from the host one."""
        buf = C.create_string_buffer(4096)
        pc = prob.as_c()
        check(lib().tslam_debug_compare_analysis(self._h, C.byref(pc), buf, C.c_int(4096)))
        return buf.value.decode().split()

    # ---- device-resident handles (benchmarks) --------------------------------------------------
    def upload(self, prob):
        return DeviceProblem(self, prob)


class DeviceProblem:
    def __init__(self, ctx, prob):
        self.ctx, self.prob = ctx, prob
        self._h = C.c_void_p()
        pc = prob.as_c()
        check(lib().tslam_dev_upload(ctx._h, C.byref(pc), C.byref(self._h)))

    def free(self):
        if self._h:
            lib().tslam_dev_free(self.ctx._h, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def eval_points(self, kind, reps=1, flush_l2=False):
        ms = C.c_float()
        check(lib().tslam_dev_eval_points(self.ctx._h, self._h, C.c_int(kind), C.c_int(reps), C.c_int(int(flush_l2)), C.byref(ms)))
        return ms.value

    def eval_text(self, kind, jac_mode=JAC_ANALYTIC, reps=1, flush_l2=False):
        ms = C.c_float()
        check(lib().tslam_dev_eval_text(self.ctx._h, self._h, C.c_int(kind), C.c_int(jac_mode), C.c_int(reps),
                                        C.c_int(int(flush_l2)), C.byref(ms)))
        return ms.value

    def download_eval(self, which, ncols):
        n = self.prob.n_pobs if which == 0 else self.prob.n_tobs
        rows = 2 if which == 0 else 8
        r = np.zeros((n, rows)); J = np.zeros((n, rows, ncols))
        check(lib().tslam_dev_download_eval(self.ctx._h, self._h, C.c_int(which), _dp(r), _dp(J), C.c_int(ncols)))
        return r, J

    def lm_iterations(self, iters, max_iters=None, text_jac_mode=JAC_ANALYTIC):
        opt = solve_options(max_iters if max_iters is not None else iters, text_jac_mode)
        phase = (C.c_float * 8)()
        summ = SolveSummaryC()
        check(lib().tslam_dev_lm_iterations(self.ctx._h, self._h, C.byref(opt), C.c_int(iters), phase, C.byref(summ)))
        return list(phase), summ.as_dict()

    def download_params(self):
        p = self.prob
        cams = np.zeros_like(p.cams); rho = np.zeros_like(p.rho); theta = np.zeros_like(p.theta)
        check(lib().tslam_dev_download_params(self.ctx._h, self._h, _dp(cams), _dp(rho), _dp(theta)))
        return cams, rho, theta


class PyramidLevel:
    """The candidate residual blocks of ONE PyrPoseOptim / PyrBA invocation (one pyramid level), before the Good-flag filter,
    with the reference's index maps back to the observation-quality vectors:
    p_raw[i]  -> entry of vObvGoodPts the point block i clears (vIdx2vPtsGood, src/optimizer.cc:1145, 1433);
    t_obj[j]  -> text object of block j (vIdx2vTextsGood :1201);  t_feat[j] -> its feature index (vIdx2vTextFeatsGood :1202)."""

    def __init__(self, prob, p_raw=None, t_obj=None, t_feat=None):
        self.prob = prob
        self.p_raw = np.arange(prob.n_pobs) if p_raw is None else np.asarray(p_raw, np.int64)
        self.t_obj = np.zeros(prob.n_tobs, np.int64) if t_obj is None else np.asarray(t_obj, np.int64)
        self.t_feat = np.arange(prob.n_tobs) if t_feat is None else np.asarray(t_feat, np.int64)
        assert len(self.p_raw) == prob.n_pobs and len(self.t_obj) == prob.n_tobs and len(self.t_feat) == prob.n_tobs


def run_pyramid(solve_gated, levels, chi2_mono, chi2_text, its, pts_good, texts_good, feats_good, rapid=False, on_level=None):
    """The level loop of optimizer::PoseOptim (src/optimizer.cc:172-186) / LocalBundleAdjustment (:282-289): per level, assemble
    the problem from the blocks whose Good flags are still set, solve, evaluate, gate, clear flags, carry the parameters on.
    `solve_gated(prob, gate, t_obj, obj_size, max_iters)` -> (summary, final_residuals, trace, pt_bad, tf_bad, obj_bad, counts) is
    Context.solve_gated for the product; the parity tests pass the CPU oracle's equivalent. Flags are modified in place.
    rapid (bFlag_rapid): gates off (src/optimizer.cc:1077-1080, 1343-1346). Returns the per-level summaries."""
    carried, out = None, []
    for li, lvl in enumerate(levels):
        sel_p = pts_good[lvl.p_raw]
        sel_t = texts_good[lvl.t_obj] & feats_good[lvl.t_obj, lvl.t_feat] if lvl.prob.n_tobs else np.zeros(0, bool)
        sub = lvl.prob.subset(sel_p, sel_t)
        if carried is not None:
            sub.set_params(*carried)
        if on_level is not None:
            on_level(li, sub)   # e.g. refresh mu / sigma with text_info() at the carried pose (src/optimizer.cc:1179-1184)
        t_obj = lvl.t_obj[sel_t]
        obj_size = np.bincount(t_obj, minlength=len(texts_good)).astype(np.int32)   # vSizeEachObj
        gate = gate_options(w_point=sub.w_point, chi2_mono=chi2_mono[li], w_text=sub.w_text, chi2_text=chi2_text[li],
                            gate_points=not rapid, gate_text=not rapid)
        summ, fr, _, pb, tb, ob, cnt = solve_gated(sub, gate, t_obj, obj_size, its[li])
        if not rapid:
            pts_good[lvl.p_raw[sel_p][pb == 1]] = False
            feats_good[t_obj[tb == 1], lvl.t_feat[sel_t][tb == 1]] = False
            texts_good[np.nonzero(ob == 1)[0]] = False
        carried = sub.params()
        out.append({"summary": summ, "n_point_blocks": sub.n_pobs, "n_text_blocks": sub.n_tobs, "bad": cnt, "final_residuals": fr})
    if carried is not None:
        for lvl in levels:
            lvl.prob.set_params(*carried)
    return out


class Optimizer:
    """Mirror of TextSLAM::optimizer's solve entry points (src/optimizer.h:57-70) on flattened problems.
    Iteration counts are the reference's: 10 (pose, local BA; src/optimizer.cc:180,286), 20 (global BA, :413).
    PoseOptim / LocalBundleAdjustment accept either one flattened problem (a single ceres::Solve) or the list of
    PyramidLevel inputs of the reference's coarse-to-fine loop together with the Good-flag arrays it maintains."""
    CHI2_MONO = (12.25, 12.25, 12.25, 12.25)   # src/optimizer.cc:175, 285
    CHI2_TEXT = (0.5, 0.5, 0.5, 0.95)          # :176, 286

    def __init__(self, ctx, text_jac_mode=JAC_ANALYTIC, rapid=False):
        self.ctx, self.text_jac_mode, self.rapid = ctx, text_jac_mode, rapid

    def _solve_gated(self, prob, gate, t_obj, obj_size, max_iters):
        return self.ctx.solve_gated(prob, gate, t_obj, obj_size, max_iters, self.text_jac_mode, want_trace=False)

    def _pyramid(self, levels, pts_good, texts_good, feats_good, its, on_level):
        # the reference runs pyramid levels 2, 1, 0 (and 3 first when bFlag_rapid): the last len(levels) table entries
        n = len(levels)
        assert 1 <= n <= 4
        return run_pyramid(self._solve_gated, levels, self.CHI2_MONO[4 - n:], self.CHI2_TEXT[4 - n:], [its] * n,
                           pts_good, texts_good, feats_good, rapid=self.rapid, on_level=on_level)

    def PoseOptim(self, prob, its=10, pts_good=None, texts_good=None, feats_good=None, on_level=None):
        if isinstance(prob, (list, tuple)):
            return self._pyramid(prob, pts_good, texts_good, feats_good, its, on_level)
        return self.ctx.solve(prob, its, self.text_jac_mode)

    def LocalBundleAdjustment(self, prob, its=10, pts_good=None, texts_good=None, feats_good=None, on_level=None):
        if isinstance(prob, (list, tuple)):
            return self._pyramid(prob, pts_good, texts_good, feats_good, its, on_level)
        return self.ctx.solve(prob, its, self.text_jac_mode)

    def GlobalBA(self, prob, its=20):
        return self.ctx.solve(prob, its, self.text_jac_mode)

    def InitBA(self, prob, its=10):
        """optimizer::InitBA (src/optimizer.cc:56-133): the two-view problem of the initialiser. The auto_IniBAScene / nume_IniBAText
        functors are the BA functors with the first keyframe held at the identity pose, so the flattened problem must carry that
        keyframe as a constant camera (checked here) and unit point weights."""
        assert prob.cam_fixed[0] == 1 and np.allclose(prob.cams[0], [1, 0, 0, 0, 0, 0, 0]), "InitBA: keyframe 0 is the fixed identity frame"
        assert prob.w_point == (1.0, 1.0), "auto_IniBAScene is unweighted (ScaleScene = 1)"
        return self.ctx.solve(prob, its, self.text_jac_mode)

    def OptimizeLandmarker(self, prob, its=50):
        """optimizer::OptimizeLandmarker / PyrLandmarkers (src/optimizer.cc:456-562, 1853-2168): every pose constant, inverse
        depths (auto_RhoScene) and planes (nume_thetaText) free, 50 iterations per level."""
        assert prob.cam_fixed.all(), "OptimizeLandmarker keeps every keyframe pose constant"
        return self.ctx.solve(prob, its, self.text_jac_mode)

    def ThetaOptimMultiFs(self, prob, its=50):
        """optimizer::ThetaOptimMultiFs / PyrThetaOptim (src/optimizer.cc:2170-2242): plane parameters of text objects from several
        frames with constant poses, followed by the 3x3 covariance of every plane (ceres::Covariance). PyrThetaOptim never sets
        max_num_iterations (:2203-2209), so Ceres' default of 50 applies; it passes loss_function = nullptr (:2176) and
        nume_thetaText is unweighted (include/nume_thetaText.h:67), hence the two preconditions checked here.
        Returns (summary, final_residuals, trace, covariances (n_planes, 3, 3), flag_variance): flag_variance = False when a
        covariance is singular — PyrThetaOptim itself still returns true in that case (:2219-2241), it only leaves
        thetaVariance untouched, so the flag is reported beside the result and does not mean the solve failed."""
        assert prob.cam_fixed.all(), "ThetaOptimMultiFs keeps every pose constant"
        assert prob.w_text == 1.0, "nume_thetaText is unweighted (include/nume_thetaText.h:67)"
        assert prob.huber_text <= 0, "PyrThetaOptim uses no loss function (src/optimizer.cc:2176)"
        summ, fr, tr = self.ctx.solve(prob, its, self.text_jac_mode)
        cov, n_singular = self.ctx.theta_covariance(prob, self.text_jac_mode)
        return summ, fr, tr, cov, n_singular == 0


class ORBextractor:
    """Mirror of TextSLAM::ORBextractor (src/ORBextractor.h:45-114): ctor args and operator()."""

    def __init__(self, ctx, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, blur_variant=0):
        self.ctx = ctx
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = C.c_void_p()
        check(lib().tslam_orb_create(ctx._h, C.c_int(nfeatures), C.c_float(scaleFactor), C.c_int(nlevels),
                                     C.c_int(iniThFAST), C.c_int(minThFAST), C.c_int(blur_variant), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().tslam_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptrs(self, imgs):
        """Row pointers + the row stride in bytes: a view whose rows are spaced wider than the image (cv::Mat with step > cols, a
        region of interest of a larger frame) is passed as it lies, without a repacking copy."""
        imgs = np.asarray(imgs, dtype=np.uint8)
        if imgs.ndim == 2:
            imgs = imgs[None]
        n, h, w = imgs.shape
        if imgs.strides[2] != 1 or imgs.strides[1] < w:
            imgs = np.ascontiguousarray(imgs)
        self._stride = int(imgs.strides[1])
        ptrs = (C.c_void_p * n)(*[imgs[i].ctypes.data for i in range(n)])
        return imgs, ptrs, n, h, w

    def extract_batch(self, imgs, max_kp=None):
        """imgs: (n,h,w) u8. Returns list of (keypoints structured array, descriptors (k,32) u8)."""
        imgs, ptrs, n, h, w = self._ptrs(imgs)
        max_kp = max_kp or (self.nfeatures + 4 * self.nlevels + 64)
        kp = np.zeros((n, max_kp), dtype=KP_DTYPE)
        desc = np.zeros((n, max_kp, 32), dtype=np.uint8)
        cnt = np.zeros(n, dtype=np.int32)
        check(lib().tslam_orb_extract(self._h, ptrs, C.c_int(n), C.c_int(w), C.c_int(h), C.c_int(self._stride), C.c_int(max_kp),
                                      kp.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(c_bp),
                                      cnt.ctypes.data_as(C.POINTER(C.c_int32))))
        return [(kp[i, :cnt[i]].copy(), desc[i, :cnt[i]].copy()) for i in range(n)]

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors) — mask ignored like the reference (src/ORBextractor.cc:1054)."""
        return self.extract_batch(image)[0]

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(lib().tslam_orb_level_size(self._h, C.c_int(level), C.byref(w), C.byref(h)))
        return w.value, h.value

    def get_level(self, img, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), dtype=np.uint8)
        check(lib().tslam_orb_get_level(self._h, C.c_int(img), C.c_int(level), out.ctypes.data_as(c_bp)))
        return out

    def debug_get(self, what, img, level):
        """Intermediate stages of the last extract call (parity tests): 0 measure plane, 1 candidates, 2 winners."""
        if what == 0:
            w, h = self.level_size(level)
            out = np.zeros((h, w), dtype=np.uint8)
            check(lib().tslam_orb_debug_get(self._h, C.c_int(0), C.c_int(img), C.c_int(level), out.ctypes.data_as(C.c_void_p), C.c_int(out.nbytes)))
            return out
        buf = np.zeros(1 + 3 * 140000, dtype=np.int32)
        check(lib().tslam_orb_debug_get(self._h, C.c_int(what), C.c_int(img), C.c_int(level), buf.ctypes.data_as(C.c_void_p), C.c_int(buf.nbytes)))
        return buf[1:1 + 3 * buf[0]].reshape(-1, 3).copy()

    def dev_bench(self, imgs, reps=5):
        imgs, ptrs, n, h, w = self._ptrs(imgs)
        ms = C.c_float(); nk = C.c_int64()
        check(lib().tslam_orb_dev_bench(self._h, ptrs, C.c_int(n), C.c_int(w), C.c_int(h), C.c_int(self._stride), C.c_int(reps),
                                        C.byref(ms), C.byref(nk)))
        return ms.value, nk.value


def analyze_structure(prob, rank=0, world=1):
    """Host-only structure analysis of a problem (no GPU needed): what tslam_solve derives before its first kernel."""
    from ._abi import StructureInfoC
    info = StructureInfoC()
    pc = prob.as_c()
    check(lib().tslam_analyze_structure(C.byref(pc), C.c_int(rank), C.c_int(world), C.byref(info)))
    return info.as_dict()


def text_info(ctx, imgs, quads, quad_img):
    """tool::CalTextinfo for a batch of projected text quads: returns (ok, mu, sigma) arrays."""
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    if imgs.ndim == 2:
        imgs = imgs[None]
    n, h, w = imgs.shape
    q = np.ascontiguousarray(quads, dtype=np.float64).reshape(-1, 8)
    qi = np.ascontiguousarray(quad_img, dtype=np.int32)
    mu = np.zeros(len(q)); sg = np.zeros(len(q)); ok = np.zeros(len(q), dtype=np.int32)
    check(lib().tslam_text_info(ctx._h, imgs.ctypes.data_as(c_bp), C.c_int(n), C.c_int(w), C.c_int(h), _dp(q), qi.ctypes.data_as(C.POINTER(C.c_int32)),
                                C.c_int(len(q)), _dp(mu), _dp(sg), ok.ctypes.data_as(C.POINTER(C.c_int32))))
    return ok.astype(bool), mu, sg


def match_hamming(ctx, query_desc, train_desc, cand_ptr, cand_idx):
    """Core of tracking::SearchFrom3D*: per query the first minimum-Hamming-distance candidate. Returns (idx, dist, second)."""
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    cp = np.ascontiguousarray(cand_ptr, dtype=np.int32); ci = np.ascontiguousarray(cand_idx, dtype=np.int32)
    nq = len(q)
    bi = np.zeros(nq, dtype=np.int32); bd = np.zeros(nq, dtype=np.int32); sd = np.zeros(nq, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    check(lib().tslam_match_hamming(ctx._h, q.ctypes.data_as(c_bp), C.c_int(nq), t.ctypes.data_as(c_bp), C.c_int(len(t)), ip(cp),
                                    ip(ci) if len(ci) else None, ip(bi), ip(bd), ip(sd)))
    return bi, bd, sd


# ---- projection-guided matching (tracking::SearchFrom3D*, src/tracking.cc:1114-1345) ---------------------------------
FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48     # src/frame.h:26-27
TH_HIGH, TH_LOW = 100, 50                     # src/tracking.cc:21-22


class FrameGridC(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("inv_w", C.c_float), ("inv_h", C.c_float), ("cell_ptr", C.POINTER(C.c_int32)), ("cell_idx", C.POINTER(C.c_int32))]


class FrameGrid:
    """frame::AssignFeaturesToGrid / PosInGrid (src/frame.cc:376-406): keypoints binned into the 64 x 48 grid, insertion order
    inside a cell; bounds and cell sizes as frame::frame sets them (src/frame.cc:118-125)."""

    def __init__(self, kp_xy, width, height, cols=FRAME_GRID_COLS, rows=FRAME_GRID_ROWS):
        f32 = np.float32
        self.cols, self.rows = cols, rows
        self.min_x, self.min_y, self.max_x, self.max_y = f32(0.0), f32(0.0), f32(width), f32(height)
        self.inv_w = f32(float(cols) / float(self.max_x - self.min_x)); self.inv_h = f32(float(rows) / float(self.max_y - self.min_y))
        kp_xy = np.ascontiguousarray(kp_xy, dtype=f32).reshape(-1, 2)
        # round(): half away from zero, on the float product
        px = (kp_xy[:, 0] - self.min_x) * self.inv_w; py = (kp_xy[:, 1] - self.min_y) * self.inv_h
        rnd = lambda v: np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), -np.floor(-v.astype(np.float64) + 0.5)).astype(np.int64)
        gx, gy = rnd(px), rnd(py)
        ok = (gx >= 0) & (gx < cols) & (gy >= 0) & (gy < rows)
        cell = np.where(ok, gx * rows + gy, -1)
        idx = np.nonzero(ok)[0]
        order = idx[np.argsort(cell[idx], kind="stable")]           # stable: insertion (keypoint index) order inside a cell
        self.cell_idx = np.ascontiguousarray(order, dtype=np.int32)
        self.cell_ptr = np.zeros(cols * rows + 1, np.int32)
        np.cumsum(np.bincount(cell[idx], minlength=cols * rows), out=self.cell_ptr[1:])

    def as_c(self):
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        return FrameGridC(self.cols, self.rows, float(self.min_x), float(self.min_y), float(self.max_x), float(self.max_y), float(self.inv_w), float(self.inv_h),
                          ip(self.cell_ptr), ip(self.cell_idx) if len(self.cell_idx) else None)


def search_from_3d(ctx, Tcw, K, pt_ray, pt_rho, poses, pt_host, pt_query, query_desc, kp_xy, kp_octave, train_desc, grid, th, min_level=-1, max_level=1):
    """Per map point: projection, bounds test, GetFeaturesInArea(u, v, th * 1.2f, min_level, max_level), first best candidate.
    Returns (best_idx, best_dist, uv)."""
    f64 = lambda a, shape: np.ascontiguousarray(a, dtype=np.float64).reshape(shape)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    Tcw = f64(Tcw, 7); K = f64(K, 4); pt_ray = f64(pt_ray, (-1, 2)); pt_rho = f64(pt_rho, -1); poses = f64(poses, (-1, 7))
    pt_host, pt_query, kp_octave = i32(pt_host), i32(pt_query), i32(kp_octave)
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    kp_xy = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
    n = len(pt_rho)
    bi = np.zeros(n, np.int32); bd = np.zeros(n, np.int32); uv = np.zeros((n, 2))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    g = grid.as_c()
    radius = np.float32(th) * np.float32(1.2)      # float radius = th*1.2f (src/tracking.cc:1153)
    check(lib().tslam_search_from_3d(ctx._h, _dp(Tcw), _dp(K), C.c_int(n), _dp(pt_ray), _dp(pt_rho), _dp(poses), C.c_int(len(poses)), ip(pt_host), ip(pt_query),
                                     q.ctypes.data_as(c_bp), C.c_int(len(q)), kp_xy.ctypes.data_as(C.POINTER(C.c_float)), ip(kp_octave),
                                     t.ctypes.data_as(c_bp), C.c_int(len(t)), C.byref(g), C.c_float(float(radius)), C.c_int(min_level), C.c_int(max_level),
                                     ip(bi), ip(bd), _dp(uv)))
    return bi, bd, uv


def search_in_area(ctx, uv, pt_query, query_desc, kp_xy, kp_octave, train_desc, grid, th, kp_skip=None, min_level=-1, max_level=-1):
    """tracking::SearchFrom3DLocalTrack's candidate search (src/tracking.cc:1296-1329): given projections (LocalTrackProj), window
    th * 1.2f, level range (-1, -1) = no level check, key points flagged in kp_skip left out. Returns (best_idx, best_dist, second_dist)."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
    pt_query, kp_octave = i32(pt_query), i32(kp_octave)
    q = np.ascontiguousarray(query_desc, dtype=np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train_desc, dtype=np.uint8).reshape(-1, 32)
    kp_xy = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
    skip = None if kp_skip is None else np.ascontiguousarray(kp_skip, dtype=np.uint8)
    n = len(uv)
    bi = np.zeros(n, np.int32); bd = np.zeros(n, np.int32); sd = np.zeros(n, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    g = grid.as_c()
    radius = np.float32(th) * np.float32(1.2)
    check(lib().tslam_search_in_area(ctx._h, C.c_int(n), _dp(uv), ip(pt_query), q.ctypes.data_as(c_bp), C.c_int(len(q)),
                                     kp_xy.ctypes.data_as(C.POINTER(C.c_float)), ip(kp_octave), skip.ctypes.data_as(c_bp) if skip is not None else None,
                                     t.ctypes.data_as(c_bp), C.c_int(len(t)), C.byref(g), C.c_float(float(radius)), C.c_int(min_level), C.c_int(max_level),
                                     ip(bi), ip(bd), ip(sd)))
    return bi, bd, sd


def resolve_matches(best_idx, best_dist, pt_query, n_kp, n_query, th_high=TH_HIGH, m32=None, m23=None):
    """The sequential bookkeeping of tracking::SearchFrom3D (src/tracking.cc:1178-1186): a point keeps its best keypoint if neither
    that keypoint nor the point's observation in the last key frame has been taken by an earlier point. SearchFrom3DAdd (:1196-1270)
    is the same loop on the vMatch3D2D / vMatch2D3D of an earlier search (pass them as m32 / m23; its points with m32 >= 0 are
    skipped — give them pt_query = -1 for the device search as well). Returns (vMatch3D2D, vMatch2D3D, nMatches)."""
    m32 = np.full(len(best_idx), -1, np.int64) if m32 is None else np.array(m32, np.int64)
    m23 = np.full(n_kp, -1, np.int64) if m23 is None else np.array(m23, np.int64)
    m12 = np.full(n_query, -1, np.int64)
    n = 0
    for i in range(len(best_idx)):
        if m32[i] >= 0 or pt_query[i] < 0 or best_idx[i] < 0 or best_dist[i] > th_high:
            continue
        j = best_idx[i]
        if m23[j] < 0 and m12[pt_query[i]] < 0:
            n += 1; m32[i] = j; m23[j] = i; m12[pt_query[i]] = j
    return m32, m23, n


def resolve_local_track(best_idx, best_dist, second_dist, th_high=TH_HIGH):
    """Acceptance rule of tracking::SearchFrom3DLocalTrack (src/tracking.cc:1331-1340): bestDist <= TH_HIGH and not
    bestDist > 0.9 * bestDist2. Returns the boolean mask of the points that get an observation (F.AddSceneObserv)."""
    bd = np.asarray(best_dist, np.int64); sd = np.asarray(second_dist, np.int64)
    return (np.asarray(best_idx) >= 0) & (bd <= th_high) & ~(bd.astype(np.float64) > 0.9 * sd.astype(np.float64))


class FramePyramid:
    """Mirror of frame::GetPyrMat (src/frame.cc:178-202): vFrameImg / vFrameGrad / vFrameGradX / vFrameGradY per level."""
    IMG, GRAD, GRAD_X, GRAD_Y = 0, 1, 2, 3

    def __init__(self, ctx, nlevels=8):
        self.ctx, self.nlevels = ctx, nlevels
        self._h = C.c_void_p()
        check(lib().tslam_frame_pyr_create(ctx._h, C.c_int(nlevels), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().tslam_frame_pyr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, imgs):
        imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
        if imgs.ndim == 2:
            imgs = imgs[None]
        n, h, w = imgs.shape
        ptrs = (C.c_void_p * n)(*[imgs[i].ctypes.data for i in range(n)])
        check(lib().tslam_frame_pyr_build(self._h, ptrs, C.c_int(n), C.c_int(w), C.c_int(h), C.c_int(w)))

    def get(self, img, level, what=0):
        w, h = C.c_int(), C.c_int()
        check(lib().tslam_frame_pyr_level_size(self._h, C.c_int(level), C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), dtype=np.uint8)
        check(lib().tslam_frame_pyr_get(self._h, C.c_int(img), C.c_int(level), C.c_int(what), out.ctypes.data_as(c_bp)))
        return out
