"""ctypes mirror of include/tslam_b200.h (struct layouts + constants).

Shared by the product binding (textslam_b200._lib) and by the test-only oracle binding
(oracle/pyoracle.py) so that both sides are driven by the *same* problem description.
"""
import ctypes as C
import numpy as np

PT_BA, PT_BA_NW, PT_POSE, PT_RHO = 0, 1, 2, 3
TX_BA, TX_POSE, TX_THETA = 0, 1, 2
JAC_ANALYTIC, JAC_CENTRAL_DIFF, JAC_ANALYTIC_TMA = 0, 1, 2
PT_NCOLS = {PT_BA: 13, PT_BA_NW: 13, PT_POSE: 6, PT_RHO: 1}
TX_NCOLS = {TX_BA: 15, TX_POSE: 6, TX_THETA: 3}
TRACE_COLS = 4

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_bp = C.POINTER(C.c_uint8)


class BAProblemC(C.Structure):
    _fields_ = [
        ("n_cams", C.c_int32), ("cams", c_dp), ("cam_fixed", c_bp),
        ("n_points", C.c_int32), ("rho", c_dp), ("rho_fixed", c_bp),
        ("n_planes", C.c_int32), ("theta", c_dp), ("theta_fixed", c_bp),
        ("n_pobs", C.c_int32), ("p_uv", c_dp), ("p_ray", c_dp), ("p_cam", c_ip), ("p_host", c_ip), ("p_lm", c_ip),
        ("K_point", C.c_double * 4), ("w_point", C.c_double * 2), ("huber_point", C.c_double),
        ("n_tobs", C.c_int32), ("t_rays", c_dp), ("t_iref", c_dp), ("t_musigma", c_dp),
        ("t_cam", c_ip), ("t_host", c_ip), ("t_plane", c_ip), ("t_img", c_ip),
        ("n_imgs", C.c_int32), ("img_w", C.c_int32), ("img_h", C.c_int32), ("imgs", c_bp),
        ("K_text", C.c_double * 4), ("w_text", C.c_double), ("huber_text", C.c_double),
    ]


class SolveOptionsC(C.Structure):
    _fields_ = [
        ("max_iters", C.c_int32), ("text_jac_mode", C.c_int32), ("n_threads", C.c_int32), ("dense_full", C.c_int32),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("initial_radius", C.c_double),
    ]


class SolveSummaryC(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("successful_steps", C.c_int32), ("unsuccessful_steps", C.c_int32),
        ("termination", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("total_ms", C.c_double), ("solve_ms", C.c_double), ("setup_ms", C.c_double),
        ("n_free_cams", C.c_int32), ("n_free_points", C.c_int32), ("n_free_planes", C.c_int32),
        ("reduced_dim", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class StructureInfoC(C.Structure):
    _fields_ = [
        ("n_free_cams", C.c_int32), ("n_free_points", C.c_int32), ("n_free_planes", C.c_int32), ("reduced_dim", C.c_int32),
        ("n_blocks", C.c_int32), ("n_local_pobs", C.c_int32), ("n_local_tobs", C.c_int32),
        ("n_owned_points", C.c_int32), ("n_owned_planes", C.c_int32), ("n_slots_point", C.c_int32), ("n_slots_text", C.c_int32),
        ("n_tiles", C.c_int32), ("n_waves", C.c_int32),
        ("n_tile_updates", C.c_int64), ("n_schur_entries", C.c_int64), ("n_direct_entries", C.c_int64),
        ("analysis_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KeyPointC(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28 and C.sizeof(KeyPointC) == 28


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class BAProblem:
    """Host-side SoA problem description; the analogue of the `ceres::Problem` that
    optimizer.cc builds (reference: src/optimizer.cc:1106-1208, 1359-1588, 1716-1830)."""

    def __init__(self, cams, cam_fixed, rho, rho_fixed=None, theta=None, theta_fixed=None,
                 p_uv=None, p_ray=None, p_cam=None, p_host=None, p_lm=None,
                 K_point=(384.396, 382.826, 315.636, 249.183), w_point=(1.0, 1.0), huber_point=0.0,
                 t_rays=None, t_iref=None, t_musigma=None, t_cam=None, t_host=None, t_plane=None, t_img=None,
                 imgs=None, K_text=None, w_text=1.0, huber_text=0.0):
        self.cams = _f64(cams).reshape(-1, 7).copy()
        self.cam_fixed = _u8(cam_fixed).reshape(-1).copy()
        self.rho = _f64(rho if rho is not None else []).reshape(-1).copy()
        self.rho_fixed = _u8(rho_fixed if rho_fixed is not None else np.zeros(len(self.rho))).reshape(-1).copy()
        self.theta = _f64(theta if theta is not None else []).reshape(-1, 3).copy()
        self.theta_fixed = _u8(theta_fixed if theta_fixed is not None else np.zeros(len(self.theta))).reshape(-1).copy()
        e2 = np.zeros((0, 2))
        self.p_uv = _f64(p_uv if p_uv is not None else e2).reshape(-1, 2)
        self.p_ray = _f64(p_ray if p_ray is not None else e2).reshape(-1, 2)
        self.p_cam = _i32(p_cam if p_cam is not None else [])
        self.p_host = _i32(p_host if p_host is not None else [])
        self.p_lm = _i32(p_lm if p_lm is not None else [])
        self.K_point = tuple(float(v) for v in K_point)
        self.w_point = tuple(float(v) for v in w_point)
        self.huber_point = float(huber_point)
        self.t_rays = _f64(t_rays if t_rays is not None else np.zeros((0, 8, 2))).reshape(-1, 8, 2)
        self.t_iref = _f64(t_iref if t_iref is not None else np.zeros((0, 8))).reshape(-1, 8)
        self.t_musigma = _f64(t_musigma if t_musigma is not None else e2).reshape(-1, 2)
        self.t_cam = _i32(t_cam if t_cam is not None else [])
        self.t_host = _i32(t_host if t_host is not None else [])
        self.t_plane = _i32(t_plane if t_plane is not None else [])
        self.t_img = _i32(t_img if t_img is not None else [])
        self.imgs = _u8(imgs if imgs is not None else np.zeros((0, 1, 1)))
        assert self.imgs.ndim == 3
        self.K_text = tuple(float(v) for v in (K_text if K_text is not None else K_point))
        self.w_text = float(w_text)
        self.huber_text = float(huber_text)
        n = len(self.p_uv)
        assert len(self.p_ray) == n and len(self.p_cam) == n and len(self.p_host) == n and len(self.p_lm) == n
        m = len(self.t_rays)
        assert len(self.t_iref) == m and len(self.t_musigma) == m and len(self.t_cam) == m and len(self.t_host) == m
        assert len(self.t_plane) == m and len(self.t_img) == m
        assert len(self.cam_fixed) == len(self.cams)

    @property
    def n_pobs(self):
        return len(self.p_uv)

    @property
    def n_tobs(self):
        return len(self.t_rays)

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def params(self):
        return self.cams.copy(), self.rho.copy(), self.theta.copy()

    def subset(self, sel_p=None, sel_t=None):
        """The same problem restricted to the selected residual blocks (boolean masks, insertion order kept) — what the
        reference's `if(!vPtsGood[..]) continue;` / `if(!vTextFeatsGood[..][..]) continue;` filters do while the ceres::Problem
        is assembled (src/optimizer.cc:1128-1130, 1168-1169, 1188-1189). Parameter blocks and images are shared by value."""
        sp = np.ones(self.n_pobs, bool) if sel_p is None else np.asarray(sel_p, bool)
        st = np.ones(self.n_tobs, bool) if sel_t is None else np.asarray(sel_t, bool)
        assert sp.shape == (self.n_pobs,) and st.shape == (self.n_tobs,)
        return BAProblem(self.cams, self.cam_fixed, self.rho, self.rho_fixed, self.theta, self.theta_fixed,
                         self.p_uv[sp], self.p_ray[sp], self.p_cam[sp], self.p_host[sp], self.p_lm[sp],
                         self.K_point, self.w_point, self.huber_point,
                         self.t_rays[st], self.t_iref[st], self.t_musigma[st], self.t_cam[st], self.t_host[st], self.t_plane[st],
                         self.t_img[st], self.imgs, self.K_text, self.w_text, self.huber_text)

    def set_params(self, cams, rho, theta):
        self.cams[...] = cams
        self.rho[...] = rho
        self.theta[...] = theta

    def as_c(self):
        """Return a BAProblemC whose pointers alias this object's numpy arrays (keep `self` alive)."""
        s = BAProblemC()

        def dp(a):
            return a.ctypes.data_as(c_dp) if a.size else C.cast(None, c_dp)

        def ip(a):
            return a.ctypes.data_as(c_ip) if a.size else C.cast(None, c_ip)

        def bp(a):
            return a.ctypes.data_as(c_bp) if a.size else C.cast(None, c_bp)

        s.n_cams, s.cams, s.cam_fixed = len(self.cams), dp(self.cams), bp(self.cam_fixed)
        s.n_points, s.rho, s.rho_fixed = len(self.rho), dp(self.rho), bp(self.rho_fixed)
        s.n_planes, s.theta, s.theta_fixed = len(self.theta), dp(self.theta), bp(self.theta_fixed)
        s.n_pobs = self.n_pobs
        s.p_uv, s.p_ray, s.p_cam, s.p_host, s.p_lm = dp(self.p_uv), dp(self.p_ray), ip(self.p_cam), ip(self.p_host), ip(self.p_lm)
        s.K_point = (C.c_double * 4)(*self.K_point)
        s.w_point = (C.c_double * 2)(*self.w_point)
        s.huber_point = self.huber_point
        s.n_tobs = self.n_tobs
        s.t_rays, s.t_iref, s.t_musigma = dp(self.t_rays), dp(self.t_iref), dp(self.t_musigma)
        s.t_cam, s.t_host, s.t_plane, s.t_img = ip(self.t_cam), ip(self.t_host), ip(self.t_plane), ip(self.t_img)
        s.n_imgs = self.imgs.shape[0]
        s.img_h, s.img_w = (self.imgs.shape[1], self.imgs.shape[2]) if self.imgs.shape[0] else (0, 0)
        s.imgs = bp(self.imgs)
        s.K_text = (C.c_double * 4)(*self.K_text)
        s.w_text = self.w_text
        s.huber_text = self.huber_text
        return s


def solve_options(max_iters=10, text_jac_mode=JAC_ANALYTIC, n_threads=1, dense_full=0,
                  function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, initial_radius=0.0):
    o = SolveOptionsC()
    o.max_iters, o.text_jac_mode, o.n_threads, o.dense_full = max_iters, text_jac_mode, n_threads, dense_full
    o.function_tolerance, o.gradient_tolerance = function_tolerance, gradient_tolerance
    o.parameter_tolerance, o.initial_radius = parameter_tolerance, initial_radius
    return o


class GateOptionsC(C.Structure):
    """tslam_gate_options (include/tslam_b200.h)."""
    _fields_ = [
        ("gate_points", C.c_int32), ("gate_text", C.c_int32), ("w_point", C.c_double * 2), ("chi2_mono", C.c_double),
        ("relax_below_text_blocks", C.c_int32), ("relax_amount", C.c_double), ("w_text", C.c_double), ("chi2_text", C.c_double),
        ("text_ratio", C.c_double),
    ]


def gate_options(w_point=(1.0 / 1.2, 1.0 / 1.2), chi2_mono=12.25, w_text=1.0 / 0.2, chi2_text=0.5, text_ratio=0.99,
                 relax_below_text_blocks=50, relax_amount=4.0, gate_points=True, gate_text=True):
    """Defaults = the constants of PyrPoseOptim / PyrBA (src/optimizer.cc:175-176, 1082-1087, 1240-1241)."""
    g = GateOptionsC()
    g.gate_points, g.gate_text = int(gate_points), int(gate_text)
    g.w_point = (C.c_double * 2)(*[float(v) for v in w_point])
    g.chi2_mono, g.relax_below_text_blocks, g.relax_amount = float(chi2_mono), int(relax_below_text_blocks), float(relax_amount)
    g.w_text, g.chi2_text, g.text_ratio = float(w_text), float(chi2_text), float(text_ratio)
    return g
