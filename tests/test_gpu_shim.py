"""Row a10 of the scope table as code: the class-surface shims (shim/optimizer_b200.cc, shim/ORBextractor_b200.cc) rebuild the
flat problem by walking a TextSLAM object graph exactly where the reference's Pyr* functions do (src/optimizer.cc:1106-1208,
1359-1588, 1716-1830, 1869-1972, 2175-2200). shim/test_shim.cc turns a synthetic flat problem back into such a graph (keyframes,
inverse-depth points hosted in keyframes — some outside the window —, text objects with reference features and box rays,
per-keyframe observation lists and Good flags), calls through the class surface and returns the mutated graph; it must agree with
the flat solve of the same problem through the Python mirror (same C-ABI, same level loop) and with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import textslam_b200 as T
from textslam_b200 import synth
from textslam_b200.api import PyramidLevel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim", "test_shim")


def build_shim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "shim")])


def test_shim_builds_against_stub_headers():
    """CPU: the shims compile against the stub TextSLAM / OpenCV / Eigen types and link the product library only."""
    build_shim()
    assert os.path.exists(SHIM)
    out = subprocess.run(["ldd", SHIM], capture_output=True, text=True).stdout
    assert "libtslam_b200.so" in out and "oracle" not in out


def plane_boxes(prob):
    """four box-corner rays per plane: the bounding rectangle of its pattern rays (vTextDeteRay)"""
    box = np.zeros((len(prob.theta), 4, 2))
    for t in range(len(prob.theta)):
        r = prob.t_rays[prob.t_plane == t].reshape(-1, 2)
        if len(r) == 0:
            continue
        x0, y0, x1, y1 = r[:, 0].min(), r[:, 1].min(), r[:, 0].max(), r[:, 1].max()
        box[t] = [[x0, y0], [x1, y0], [x1, y1], [x0, y1]]
    return box


def quat_rot(q):
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def refresh_musigma(ctx, sub, box):
    """tool::GetProjText + CalTextinfo per (keyframe, text object) with the current parameters (src/optimizer.cc:1179-1184, 1478-1492)"""
    if sub.n_tobs == 0:
        return
    key = np.stack([sub.t_cam, sub.t_plane], 1)
    first = np.r_[True, (key[1:] != key[:-1]).any(1)]
    starts = np.nonzero(first)[0]
    quads, qimg = [], []
    for s in starts:
        c, h, t = sub.t_cam[s], sub.t_host[s], sub.t_plane[s]
        Rc, Rh = quat_rot(sub.cams[c, :4]), quat_rot(sub.cams[h, :4])
        R = Rc @ Rh.T
        tt = sub.cams[c, 4:] - R @ sub.cams[h, 4:]
        q = []
        for rx, ry in box[t]:
            ray = np.array([rx, ry, 1.0])
            rho = -(ray @ sub.theta[t])
            p = R @ ray / rho + tt
            q += [sub.K_text[0] * p[0] / p[2] + sub.K_text[2], sub.K_text[1] * p[1] / p[2] + sub.K_text[3]]
        quads.append(q); qimg.append(sub.t_img[s])
    ok, mu, sg = T.text_info(ctx, sub.imgs, np.array(quads), np.array(qimg))
    obj = np.cumsum(first) - 1
    sub.t_musigma[:, 0] = mu[obj]; sub.t_musigma[:, 1] = sg[obj]


def run_shim(prob, mode, n_window, box, tmp_path, nlevels=4):
    build_shim()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    n_imgs = len(prob.imgs) if prob.n_tobs else 0
    h, w = (prob.imgs.shape[1:] if n_imgs else (1, 1))
    with open(fin, "wb") as f:
        np.array([len(prob.cams), len(prob.rho), len(prob.theta), prob.n_pobs, prob.n_tobs, n_imgs, w, h, mode, n_window, nlevels, 0], np.int32).tofile(f)
        np.asarray(prob.K_point, np.float64).tofile(f)
        for a in (prob.cams, prob.rho, prob.theta, prob.p_uv, prob.p_ray):
            np.ascontiguousarray(a, np.float64).tofile(f)
        for a in (prob.p_cam, prob.p_host, prob.p_lm):
            np.ascontiguousarray(a, np.int32).tofile(f)
        if prob.n_tobs:
            np.ascontiguousarray(prob.t_rays, np.float64).tofile(f); np.ascontiguousarray(prob.t_iref, np.float64).tofile(f)
            for a in (prob.t_cam, prob.t_host, prob.t_plane, prob.t_img):
                np.ascontiguousarray(a, np.int32).tofile(f)
        np.ascontiguousarray(box, np.float64).tofile(f)
        if n_imgs:
            np.ascontiguousarray(prob.imgs, np.uint8).tofile(f)
    p = subprocess.run([SHIM, fin, fout], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    with open(fout, "rb") as f:
        counts = np.fromfile(f, np.int32, 8)
        cams = np.fromfile(f, np.float64, 7 * len(prob.cams)).reshape(-1, 7)
        rho = np.fromfile(f, np.float64, len(prob.rho))
        theta = np.fromfile(f, np.float64, 3 * len(prob.theta)).reshape(-1, 3)
        cov = np.fromfile(f, np.float64, 9).reshape(3, 3)
        cost = np.fromfile(f, np.float64, 1)[0]
    return counts, cams, rho, theta, cov, cost


def rel(x, y):
    return float(np.abs(x - y).max() / (np.abs(y).max() + 1e-300)) if y.size else 0.0


def quat_same(a, b):
    """quaternions up to sign"""
    s = np.sign((a[:, :4] * b[:, :4]).sum(1))[:, None]
    return np.concatenate([a[:, :4] * s, a[:, 4:]], 1)


@pytest.mark.gpu
def test_global_ba_through_the_class_surface(ctx, oracle, tmp_path):
    """optimizer::GlobalBA -> PyrGlobalBA (src/optimizer.cc:334-453, 1701-1851): 20 iterations, unweighted points, KF 0/1 fixed"""
    prob = synth.c5_global_ba(seed=91, n_kf=40, n_lm=1500)
    counts, cams, rho, theta, _, cost = run_shim(prob, 0, len(prob.cams), np.zeros((0, 4, 2)), tmp_path)
    a, b = prob.copy(), prob.copy()
    sg, _, _ = ctx.solve(a, 20)
    so, _, _ = oracle.solve(b, 20)
    assert counts[3] == sg["iterations"] == so["iterations"]
    assert abs(cost - so["final_cost"]) <= 1e-7 * so["final_cost"]
    assert rel(quat_same(cams, b.cams), b.cams) < 1e-5 and rel(rho, b.rho) < 1e-5
    assert rel(quat_same(cams, a.cams), a.cams) < 1e-7 and rel(rho, a.rho) < 1e-7


def _objects(prob):
    key = np.stack([prob.t_cam, prob.t_plane], 1)
    first = np.r_[True, (key[1:] != key[:-1]).any(1)]
    t_obj = np.cumsum(first) - 1
    starts = np.nonzero(first)[0]
    t_feat = np.arange(prob.n_tobs) - starts[t_obj]
    return t_obj, t_feat


@pytest.mark.gpu
def test_local_ba_through_the_class_surface(ctx, tmp_path):
    """optimizer::LocalBundleAdjustment -> PyrBA x 3 levels (src/optimizer.cc:197-331, 1330-1698): landmarks hosted outside the
    window switch to the pose-only functors, mu / sigma are refreshed per level from the projected box, the chi^2 gates clear
    Good flags between the levels, mnId 0/1 and the first three participating keyframes are fixed (LOCAL)."""
    prob = synth.make_ba_problem(seed=92, n_kf=8, n_lm=300, obs_per_lm=3, band=8, fixed_cams=(0, 1, 2), n_ext=3, frac_ext_lm=0.3, n_planes=6)
    order = np.lexsort((np.arange(prob.n_pobs), prob.p_cam))   # the graph walk visits keyframe by keyframe
    for name in ("p_uv", "p_ray", "p_cam", "p_host", "p_lm"):
        setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[order]))
    box = plane_boxes(prob)
    counts, cams, rho, theta, _, _ = run_shim(prob, 1, 8, box, tmp_path)
    # the same three solves through the Python mirror of the level loop
    t_obj, t_feat = _objects(prob)
    n_obj = int(t_obj.max()) + 1
    pts_good = np.ones(prob.n_pobs, bool); texts_good = np.ones(n_obj, bool); feats_good = np.ones((n_obj, int(t_feat.max()) + 1), bool)
    q = prob.copy()
    levels = [PyramidLevel(q, np.arange(prob.n_pobs), t_obj, t_feat) for _ in range(3)]
    T.Optimizer(ctx).LocalBundleAdjustment(levels, 10, pts_good, texts_good, feats_good, on_level=lambda li, sub: refresh_musigma(ctx, sub, box))
    assert counts[0] == (~pts_good).sum() and counts[1] == (~feats_good).sum() and counts[2] == (~texts_good).sum(), (counts, (~pts_good).sum())
    assert rel(quat_same(cams[:8], q.cams[:8]), q.cams[:8]) < 1e-6
    free = prob.rho_fixed == 0
    assert rel(rho[free], q.rho[free]) < 1e-6 and rel(theta[prob.theta_fixed == 0], q.theta[prob.theta_fixed == 0]) < 1e-6
    assert np.array_equal(cams[8:], quat_same(cams[8:], prob.cams[8:])) or rel(quat_same(cams[8:], prob.cams[8:]), prob.cams[8:]) < 1e-12   # hosts outside the window untouched


@pytest.mark.gpu
def test_pose_optim_through_the_class_surface(ctx, tmp_path):
    """optimizer::PoseOptim -> PyrPoseOptim x 3 levels (src/optimizer.cc:135-195, 1060-1327)"""
    prob = synth.c3_pose_only(seed=93, n_pobs=800, n_planes=6)
    box = plane_boxes(prob)
    counts, cams, rho, theta, _, _ = run_shim(prob, 2, 1, box, tmp_path)
    t_obj, t_feat = _objects(prob)
    n_obj = int(t_obj.max()) + 1
    pts_good = np.ones(prob.n_pobs, bool); texts_good = np.ones(n_obj, bool); feats_good = np.ones((n_obj, int(t_feat.max()) + 1), bool)
    q = prob.copy()
    levels = [PyramidLevel(q, np.arange(prob.n_pobs), t_obj, t_feat) for _ in range(3)]
    T.Optimizer(ctx).PoseOptim(levels, 10, pts_good, texts_good, feats_good, on_level=lambda li, sub: refresh_musigma(ctx, sub, box))
    assert counts[0] == (~pts_good).sum() and counts[1] == (~feats_good).sum() and counts[2] == (~texts_good).sum()
    assert rel(quat_same(cams[:1], q.cams[:1]), q.cams[:1]) < 1e-7


@pytest.mark.gpu
def test_orb_extractor_through_the_class_surface(ctx, oracle, tmp_path):
    """ORBextractor::operator() (src/ORBextractor.cc:1054-1116) through the shim: cv::KeyPoint vector + descriptor Mat, bit-exact"""
    build_shim()
    img = synth.orb_images(seed=94, n=1)[0]
    fin, fout = str(tmp_path / "img.bin"), str(tmp_path / "kp.bin")
    with open(fin, "wb") as f:
        np.array([img.shape[1], img.shape[0], 1000], np.int32).tofile(f); img.tofile(f)
    p = subprocess.run([SHIM, "--orb", fin, fout], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    with open(fout, "rb") as f:
        n = int(np.fromfile(f, np.int32, 1)[0])
        kp = np.fromfile(f, T.KP_DTYPE, n)
        desc = np.fromfile(f, np.uint8, n * 32).reshape(n, 32)
    kpo, desco = oracle.orb_extract(img, 1000, 1.2, 8, 20, 7)
    assert n == len(kpo) and np.array_equal(desc, desco)
    for fld in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(kp[fld], kpo[fld]), fld


@pytest.mark.gpu
def test_optimize_landmarker_through_the_class_surface(ctx, tmp_path):
    """optimizer::OptimizeLandmarker -> PyrLandmarkers x 4 levels (src/optimizer.cc:456-562, 1853-2168): every pose constant, every
    inverse depth (auto_RhoScene, unweighted, Huber sqrt(5.991)) and plane (nume_thetaText, Huber 2.0) free, 50 iterations per level,
    only the point gate (chi^2 = 18) between the levels."""
    prob = synth.make_ba_problem(seed=95, n_kf=6, n_lm=200, obs_per_lm=3, band=6, fixed_cams=(0, 1), n_ext=2, frac_ext_lm=0.3, n_planes=5)
    order = np.lexsort((np.arange(prob.n_pobs), prob.p_cam))
    for name in ("p_uv", "p_ray", "p_cam", "p_host", "p_lm"):
        setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[order]))
    box = plane_boxes(prob)
    counts, cams, rho, theta, _, _ = run_shim(prob, 3, 6, box, tmp_path)
    q = prob.copy()
    q.cam_fixed[:] = 1; q.rho_fixed[:] = 0; q.theta_fixed[:] = 0
    q.w_point, q.huber_point, q.w_text, q.huber_text = (1.0, 1.0), float(np.sqrt(5.991)), 1.0, 2.0
    t_obj, _ = _objects(q)
    pts_good = np.ones(q.n_pobs, bool)
    g = T.gate_options(w_point=(1.0, 1.0), chi2_mono=18.0, w_text=1.0, chi2_text=1.5, gate_points=True, gate_text=False, relax_below_text_blocks=0)
    for _ in range(4):
        sub = q.subset(pts_good, np.ones(q.n_tobs, bool))
        refresh_musigma(ctx, sub, box)
        obj_size = np.bincount(t_obj, minlength=int(t_obj.max()) + 1).astype(np.int32)
        summ, fr, _, pb, tb, ob, cnt = ctx.solve_gated(sub, g, t_obj, obj_size, 50, want_trace=False)
        pts_good[np.nonzero(pts_good)[0][pb == 1]] = False
        q.set_params(*sub.params())
    # the reference keeps its vPtsGood copy local to OptimizeLandmarker (src/optimizer.cc:476-479, 532-540): the gate shapes the levels that
    # follow, the keyframes' flags stay as they were
    assert counts[0] == 0 and (~pts_good).sum() > 0
    seen = np.zeros(len(q.rho), bool); seen[q.p_lm] = True
    assert rel(rho[seen], q.rho[seen]) < 1e-6 and rel(theta, q.theta) < 1e-6
    assert rel(quat_same(cams[:6], prob.cams[:6]), prob.cams[:6]) < 1e-12       # poses untouched


@pytest.mark.gpu
def test_theta_optim_multi_fs_through_the_class_surface(ctx, tmp_path):
    """optimizer::ThetaOptimMultiFs -> PyrThetaOptim x 3 levels (src/optimizer.cc:565-624, 2170-2242): the plane of one text object from
    every keyframe that sees it plus the current frame, poses relative to the object's host keyframe held constant, 50 iterations per
    level (Ceres' default), no loss; then the 3x3 covariance of the plane."""
    prob = synth.c5_global_ba(seed=96, n_kf=12, n_lm=60, n_planes=4, text_kf_stride=1)
    # the synthetic planes are seen from one keyframe each: give the first object a second observer (keyframe 5, its own image), so that
    # the problem has two frames besides the current one's re-insertion logic (blocks stay grouped by keyframe, then object)
    g0 = np.nonzero(prob.t_plane == prob.t_plane.min())[0]
    assert prob.t_cam.max() < 5 and len(prob.imgs) > 5
    for name in ("t_rays", "t_iref", "t_musigma", "t_host", "t_plane"):
        setattr(prob, name, np.ascontiguousarray(np.concatenate([getattr(prob, name), getattr(prob, name)[g0]])))
    prob.t_cam = np.ascontiguousarray(np.concatenate([prob.t_cam, np.full(len(g0), 5)]).astype(np.int32))
    prob.t_img = np.ascontiguousarray(np.concatenate([prob.t_img, np.full(len(g0), 5)]).astype(np.int32))
    box = plane_boxes(prob)
    counts, cams, rho, theta, cov, _ = run_shim(prob, 4, len(prob.cams), box, tmp_path)
    assert counts[5] == 1
    t0 = int(prob.t_plane.min())                       # M.vTexts[0]: the first text object of the map
    cF = int(prob.t_cam[0])                            # F: the keyframe of the first text block
    grp = np.nonzero(prob.t_plane == t0)[0]
    host = int(prob.t_host[grp[0]])
    first = grp[(prob.t_cam[grp] == prob.t_cam[grp[0]])]          # reference features = the object's first (keyframe, object) group
    observers = [int(c) for c in np.unique(prob.t_cam[grp]) if c != cF and c != host] + [cF]
    img_of = {int(c): int(i) for c, i in zip(prob.t_cam, prob.t_img)}
    qh = prob.cams[host]
    cam_rows = [np.array([1.0, 0, 0, 0, 0, 0, 0])]
    for c in observers:                                 # T_cr = T_cw T_rw^-1
        qc = prob.cams[c]
        q_cr = synth.qmul(qc[:4], synth.qconj(qh[:4]))
        cam_rows.append(np.r_[q_cr, qc[4:] - synth.qrot(q_cr, qh[4:])])
    nb = len(first)
    P = T.BAProblem(np.array(cam_rows), np.ones(len(cam_rows), np.uint8), rho=None, theta=prob.theta[t0][None], theta_fixed=[0],
                    t_rays=np.tile(prob.t_rays[first], (len(observers), 1, 1)), t_iref=np.tile(prob.t_iref[first], (len(observers), 1)),
                    t_musigma=np.ones((nb * len(observers), 2)), t_cam=np.repeat(np.arange(1, len(observers) + 1), nb), t_host=np.zeros(nb * len(observers)),
                    t_plane=np.zeros(nb * len(observers)), t_img=np.repeat(np.arange(len(observers)), nb),
                    imgs=np.stack([prob.imgs[img_of[c]] for c in observers]), K_point=prob.K_point, K_text=prob.K_text, w_text=1.0, huber_text=0.0)
    bx = box[t0][None]
    for _ in range(3):
        refresh_musigma(ctx, P, bx)
        summ, fr, tr, cv, okv = T.Optimizer(ctx).ThetaOptimMultiFs(P, 50)
    assert rel(theta[t0], P.theta[0]) < 1e-6
    assert okv and rel(cov, cv[0]) < 1e-5
    others = [t for t in range(len(prob.theta)) if t != t0]
    assert np.array_equal(theta[others], prob.theta[others])
