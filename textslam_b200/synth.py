"""Seeded synthetic problems of the BASELINE.json shapes.

The reference ships no dataset and no fixtures; these generators create the inputs
that `optimizer.cc` would flatten out of its map (src/optimizer.cc:213-279) and that
`frame.cc:330` would hand to the ORB extractor. Intrinsics are yaml/GeneralMotion.yaml:12-15.
"""
import numpy as np
from ._abi import BAProblem

K0 = (384.396, 382.826, 315.636, 249.183)
IMG_W, IMG_H = 640, 480
HUBER_POINT = float(np.sqrt(5.991))  # src/optimizer.cc:1116,1369,1724
HUBER_TEXT = 3.0                     # src/optimizer.cc:1164,1454
W_POINT = 1.0 / 1.2                  # src/optimizer.cc:1087,1350
W_TEXT = 1.0 / 0.2                   # src/optimizer.cc:1088,1351
# INTERVAL8 pattern offsets (src/tool.cc:1550-1557)
PATTERN8 = np.array([[0, 0], [2, 0], [1, -1], [0, -2], [-1, -1], [-2, 0], [-1, 1], [0, 2]], dtype=np.float64)


# ------------------------------------------------------------------------------------------
# quaternion helpers (w,x,y,z), same conventions as include/rotation.h
# ------------------------------------------------------------------------------------------
def qmul(a, b):
    a, b = np.asarray(a), np.asarray(b)
    w = a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3]
    x = a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2]
    y = a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1]
    z = a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]
    return np.stack([w, x, y, z], -1)


def qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def qrot(q, p):
    """Rotate points p (...,3) by unit quaternions q (...,4)."""
    qv = q[..., 1:]
    t = 2.0 * np.cross(qv, p)
    return p + q[..., :1] * t + np.cross(qv, t)


def qexp(v):
    """Rotation vector (full angle) -> unit quaternion."""
    v = np.asarray(v, dtype=np.float64)
    ang = np.linalg.norm(v, axis=-1, keepdims=True)
    half = 0.5 * ang
    s = np.where(ang > 1e-12, np.sin(half) / np.maximum(ang, 1e-300), 0.5)
    return np.concatenate([np.cos(half), s * v], -1)


def trajectory(idx):
    """Smooth trajectory: T_cw = (q_cw, t_cw) and camera centre for (real-valued) keyframe index."""
    idx = np.asarray(idx, dtype=np.float64)
    c = np.stack([0.05 * idx, 0.02 * np.sin(0.05 * idx), 0.01 * idx], -1)
    yaw = 0.10 * np.sin(2.0 * np.pi * idx / 200.0)
    pitch = 0.03 * np.cos(2.0 * np.pi * idx / 90.0)
    q_wc = qmul(qexp(np.stack([0 * yaw, yaw, 0 * yaw], -1)), qexp(np.stack([pitch, 0 * pitch, 0 * pitch], -1)))
    q_cw = qconj(q_wc)
    t_cw = -qrot(q_cw, c)
    return q_cw, t_cw


def project(q_cw, t_cw, Xw, K):
    Xc = qrot(q_cw, Xw) + t_cw
    z = Xc[..., 2]
    u = K[0] * Xc[..., 0] / z + K[2]
    v = K[1] * Xc[..., 1] / z + K[3]
    return u, v, z


def noise_images(rng, n, w=IMG_W, h=IMG_H, sigma=2.0):
    """Seeded band-limited noise (uniform u8 blurred with a Gaussian, contrast re-stretched)."""
    from scipy.ndimage import gaussian_filter
    out = np.empty((n, h, w), dtype=np.uint8)
    for i in range(n):
        a = rng.integers(0, 256, size=(h, w)).astype(np.float32)
        b = gaussian_filter(a, sigma, mode="mirror")
        b = (b - b.mean()) / (b.std() + 1e-9)
        out[i] = np.clip(127.5 + 55.0 * b, 0, 255).astype(np.uint8)
    return out


def bilinear(img, u, v):
    """Reference bilinear sampler (include/nume_BAText.h:67-81); 0 outside."""
    uf, vf = np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)
    uc, vc = np.ceil(u).astype(np.int64), np.ceil(v).astype(np.int64)
    h, w = img.shape
    ok = (uf >= 0) & (vf >= 0) & (uc < w) & (vc < h)
    ufc, vfc = np.clip(uf, 0, w - 1), np.clip(vf, 0, h - 1)
    u1, v1 = np.clip(ufc + 1, 0, w - 1), np.clip(vfc + 1, 0, h - 1)
    su, sv = u - uf, v - vf
    val = ((1 - su) * (1 - sv) * img[vfc, ufc] + su * (1 - sv) * img[vfc, u1]
           + (1 - su) * sv * img[v1, ufc] + su * sv * img[v1, u1])
    return np.where(ok, val, 0.0)


def make_ba_problem(seed=0, n_kf=10, n_lm=1000, obs_per_lm=3, band=10, fixed_cams=(0, 1, 2),
                    n_ext=0, frac_ext_lm=0.0,
                    n_planes=0, feats_per_plane=25, text_kf_stride=1, level=0,
                    w_point=W_POINT, huber_point=HUBER_POINT, w_text=W_TEXT, huber_text=HUBER_TEXT,
                    pix_noise=1.0, outlier_frac=0.05, rot_noise=1e-2, trans_noise=1e-2, rho_noise=0.05,
                    theta_noise=0.02, inten_noise=0.02, perturb=True):
    """Generic generator. Keyframes 0..n_kf-1 are the window; `n_ext` extra constant keyframes
    (indices n_kf..n_kf+n_ext-1, trajectory positions -1, -2, ...) host a fraction `frac_ext_lm`
    of the landmarks / planes, which are then constant too (src/optimizer.cc:240-243,1394-1430)."""
    rng = np.random.default_rng(seed)
    K = np.array(K0)
    Kt = K / (2.0 ** level)  # src/optimizer.cc:43-52
    n_cams = n_kf + n_ext
    traj_idx = np.concatenate([np.arange(n_kf, dtype=np.float64), -(np.arange(n_ext, dtype=np.float64) + 1.0)])
    q_gt, t_gt = trajectory(traj_idx)
    cams_gt = np.concatenate([q_gt, t_gt], 1)
    cam_fixed = np.zeros(n_cams, dtype=np.uint8)
    for k in fixed_cams:
        if k < n_kf:
            cam_fixed[k] = 1
    cam_fixed[n_kf:] = 1
    m = min(obs_per_lm, n_kf - 1 if n_ext == 0 else n_kf)

    def sample_hosts(n, ext_frac):
        is_ext = (rng.random(n) < ext_frac) if n_ext > 0 else np.zeros(n, dtype=bool)
        host = np.where(is_ext, n_kf + rng.integers(0, max(n_ext, 1), n), rng.integers(0, n_kf, n))
        return host.astype(np.int64), is_ext

    def sample_observers(host, is_ext, m):
        """m distinct window keyframes != host within the band (ext hosts: the first `band` keyframes)."""
        n = len(host)
        lo = np.where(is_ext, 0, np.maximum(host - band, 0))
        hi = np.where(is_ext, min(band, n_kf) - 1, np.minimum(host + band, n_kf - 1))
        width = hi - lo + 1
        # random keys -> argsort gives a random permutation of the band; drop the host itself
        keys = rng.random((n, 2 * band + 1))
        offs = np.arange(2 * band + 1)[None, :]
        cand = lo[:, None] + offs
        keys = np.where((offs < width[:, None]) & (cand != host[:, None]), keys, 2.0)
        order = np.argsort(keys, axis=1)[:, :m]
        obs = np.take_along_axis(cand, order, 1)
        ok = np.take_along_axis(keys, order, 1) < 1.5
        return obs, ok.all(1)

    # ---------------- points ----------------
    uv_l, ray_l, cam_l, host_l, lm_l, rho_l, rho_fixed_l = [], [], [], [], [], [], []
    n_have = 0
    while n_have < n_lm:
        n_try = int((n_lm - n_have) * 1.6) + 64
        host, is_ext = sample_hosts(n_try, frac_ext_lm)
        px = np.stack([rng.uniform(8, IMG_W - 8, n_try), rng.uniform(8, IMG_H - 8, n_try)], 1)
        ray = np.stack([(px[:, 0] - K[2]) / K[0], (px[:, 1] - K[3]) / K[1], np.ones(n_try)], 1)
        rho = rng.uniform(0.1, 1.0, n_try)
        Xh = ray / rho[:, None]
        qh, th = cams_gt[host, :4], cams_gt[host, 4:]
        Xw = qrot(qconj(qh), Xh - th)
        obs, ok = sample_observers(host, is_ext, m)
        u, v, z = project(cams_gt[obs, :4], cams_gt[obs, 4:], Xw[:, None, :], K)
        ok &= ((u > 4) & (u < IMG_W - 4) & (v > 4) & (v < IMG_H - 4) & (z > 0.2)).all(1)
        idx = np.nonzero(ok)[0][: n_lm - n_have]
        k = len(idx)
        if k == 0:
            continue
        uv = np.stack([u[idx], v[idx]], -1).reshape(-1, 2)
        uv_l.append(uv)
        ray_l.append(np.repeat(ray[idx, :2], m, 0))
        cam_l.append(obs[idx].reshape(-1))
        host_l.append(np.repeat(host[idx], m))
        lm_l.append(np.repeat(np.arange(n_have, n_have + k), m))
        rho_l.append(rho[idx])
        rho_fixed_l.append(is_ext[idx].astype(np.uint8))
        n_have += k
    if n_lm > 0:
        p_uv = np.concatenate(uv_l); p_ray = np.concatenate(ray_l)
        p_cam = np.concatenate(cam_l); p_host = np.concatenate(host_l); p_lm = np.concatenate(lm_l)
        rho_gt = np.concatenate(rho_l); rho_fixed = np.concatenate(rho_fixed_l)
        # the reference inserts residual blocks keyframe-major (src/optimizer.cc:1365, 1720)
        order = np.lexsort((p_lm, p_cam))
        p_uv, p_ray, p_cam, p_host, p_lm = p_uv[order], p_ray[order], p_cam[order], p_host[order], p_lm[order]
        n_obs = len(p_uv)
        p_uv = p_uv + rng.normal(0.0, pix_noise, p_uv.shape)
        out = rng.random(n_obs) < outlier_frac
        p_uv[out] += rng.uniform(-20, 20, (int(out.sum()), 2))
    else:
        p_uv = np.zeros((0, 2)); p_ray = np.zeros((0, 2)); p_cam = p_host = p_lm = np.zeros(0, dtype=np.int32)
        rho_gt = np.zeros(0); rho_fixed = np.zeros(0, dtype=np.uint8)

    # ---------------- text planes ----------------
    t_rays, t_iref, t_ms, t_cam, t_host, t_plane, t_img = [], [], [], [], [], [], []
    theta_gt = np.zeros((n_planes, 3)); theta_fixed = np.zeros(n_planes, dtype=np.uint8)
    imgs = np.zeros((0, 1, 1), dtype=np.uint8)
    if n_planes > 0:
        lw, lh = IMG_W >> level, IMG_H >> level
        text_kfs = np.arange(0, n_kf, text_kf_stride)
        img_of_kf = -np.ones(n_cams, dtype=np.int64)
        img_of_kf[text_kfs] = np.arange(len(text_kfs))
        imgs = noise_images(rng, len(text_kfs), lw, lh)
        g = int(round(np.sqrt(feats_per_plane)))
        assert g * g == feats_per_plane, "feats_per_plane must be a square number"
        grid = (np.stack(np.meshgrid(np.arange(g), np.arange(g)), -1).reshape(-1, 2) - (g - 1) / 2.0) * 6.0
        ip = 0
        guard = 0
        while ip < n_planes:
            guard += 1
            assert guard < 200 * n_planes + 1000, "text plane generation failed"
            host, is_ext = sample_hosts(1, frac_ext_lm)
            host, is_ext = int(host[0]), bool(is_ext[0])
            ang = np.deg2rad(rng.uniform(0, 30)); az = rng.uniform(0, 2 * np.pi)
            nrm = np.array([np.sin(ang) * np.cos(az), np.sin(ang) * np.sin(az), np.cos(ang)])
            d = rng.uniform(1.0, 5.0)
            theta = -nrm / d
            ctr = np.array([rng.uniform(40, lw - 40), rng.uniform(40, lh - 40)])
            feats = ctr[None, :] + grid                               # (F,2) level pixels in the host image
            pix = feats[:, None, :] + PATTERN8[None, :, :]           # (F,8,2)
            rays = np.stack([(pix[..., 0] - Kt[2]) / Kt[0], (pix[..., 1] - Kt[3]) / Kt[1], np.ones(pix.shape[:2])], -1)
            rho_t = -(rays @ theta)
            if (rho_t <= 1e-3).any():
                continue
            Xh = rays / rho_t[..., None]
            Xw = qrot(qconj(cams_gt[host, :4]), Xh - cams_gt[host, 4:])
            # observer: a text keyframe within the band, != host
            if is_ext:
                cands = text_kfs[text_kfs < min(band, n_kf)]
            else:
                cands = text_kfs[(np.abs(text_kfs - host) <= band) & (text_kfs != host)]
            if len(cands) == 0:
                continue
            obs = int(rng.choice(cands))
            u, v, z = project(cams_gt[obs, :4], cams_gt[obs, 4:], Xw, Kt)
            if not ((u > 3) & (u < lw - 4) & (v > 3) & (v < lh - 4) & (z > 0.2)).all():
                continue
            img = imgs[img_of_kf[obs]].astype(np.float64)
            x0, x1 = int(np.floor(u.min())), int(np.ceil(u.max()))
            y0, y1 = int(np.floor(v.min())), int(np.ceil(v.max()))
            patch = img[y0:y1 + 1, x0:x1 + 1].ravel()
            mu = float(patch.mean()); sg = float(np.sqrt(((patch - mu) ** 2).sum() / (len(patch) - 1)))
            if sg == 0:
                continue
            inten = bilinear(img, u, v)
            iref = (inten - mu) / sg + rng.normal(0.0, inten_noise, inten.shape)
            theta_gt[ip] = theta; theta_fixed[ip] = 1 if is_ext else 0
            F = feats.shape[0]
            t_rays.append(rays[..., :2]); t_iref.append(iref)
            t_ms.append(np.tile([mu, sg], (F, 1)))
            t_cam.append(np.full(F, obs)); t_host.append(np.full(F, host)); t_plane.append(np.full(F, ip))
            t_img.append(np.full(F, img_of_kf[obs]))
            ip += 1
        t_rays = np.concatenate(t_rays); t_iref = np.concatenate(t_iref); t_ms = np.concatenate(t_ms)
        t_cam = np.concatenate(t_cam); t_host = np.concatenate(t_host); t_plane = np.concatenate(t_plane)
        t_img = np.concatenate(t_img)
        order = np.lexsort((t_plane, t_cam))  # keyframe-major like src/optimizer.cc:1447
        # keep features of one (cam, plane) together and in order: lexsort is stable
        t_rays, t_iref, t_ms = t_rays[order], t_iref[order], t_ms[order]
        t_cam, t_host, t_plane, t_img = t_cam[order], t_host[order], t_plane[order], t_img[order]
    else:
        t_rays = t_iref = t_ms = t_cam = t_host = t_plane = t_img = None

    # ---------------- initial estimate ----------------
    cams0 = cams_gt.copy(); rho0 = rho_gt.copy(); theta0 = theta_gt.copy()
    if perturb:
        free = cam_fixed == 0
        nf = int(free.sum())
        dq = qexp(rng.normal(0.0, rot_noise, (nf, 3)))
        cams0[free, :4] = qmul(dq, cams_gt[free, :4])
        cams0[free, 4:] += rng.normal(0.0, trans_noise, (nf, 3))
        fr = rho_fixed == 0
        rho0[fr] = rho_gt[fr] * (1.0 + rng.normal(0.0, rho_noise, int(fr.sum())))
        ft = theta_fixed == 0
        theta0[ft] = theta_gt[ft] * (1.0 + rng.normal(0.0, theta_noise, (int(ft.sum()), 3)))
    prob = BAProblem(cams0, cam_fixed, rho0, rho_fixed, theta0, theta_fixed,
                     p_uv, p_ray, p_cam, p_host, p_lm, K0, (w_point, w_point), huber_point,
                     t_rays, t_iref, t_ms, t_cam, t_host, t_plane, t_img, imgs, tuple(Kt), w_text, huber_text)
    prob.gt = (cams_gt, rho_gt, theta_gt)
    return prob


# ---- the named BASELINE.json configurations ---------------------------------------
def c3_pose_only(seed=0, level=0, n_pobs=2000, n_planes=10):
    """C3: 1 free camera, 2000 auto_PoseOptimScene blocks, 10 planes x 25 nume_PoseOptimText blocks."""
    return make_ba_problem(seed=seed, n_kf=1, n_lm=n_pobs, obs_per_lm=1, band=20, fixed_cams=(), n_ext=20,
                           frac_ext_lm=1.0, n_planes=n_planes, feats_per_plane=25, level=level)


def c4_local_ba(seed=0, level=0, n_kf=10, n_lm=1000, n_planes=30):
    """C4: 10 KF (first 3 fixed), 3000 auto_BAScene blocks, 30 planes x 25 nume_BAText blocks."""
    return make_ba_problem(seed=seed, n_kf=n_kf, n_lm=n_lm, obs_per_lm=3, band=n_kf, fixed_cams=(0, 1, 2),
                           n_planes=n_planes, feats_per_plane=25, level=level)


def c5_global_ba(seed=0, n_kf=500, n_lm=25000, obs_per_lm=4, n_planes=0, text_kf_stride=5):
    """C5: 500 KF (KF 0,1 fixed), 100k auto_BASceneNW blocks; text off by default as in the
    reference (src/optimizer.cc:1707); n_planes=1000 gives the text-on variant (w_T = 1, :1813)."""
    return make_ba_problem(seed=seed, n_kf=n_kf, n_lm=n_lm, obs_per_lm=obs_per_lm, band=10, fixed_cams=(0, 1),
                           n_planes=n_planes, feats_per_plane=25, text_kf_stride=text_kf_stride,
                           w_point=1.0, w_text=1.0)


def orb_images(seed=0, n=64, w=IMG_W, h=IMG_H):
    """C2: seeded band-limited-noise u8 images with corner-like structure for FAST."""
    rng = np.random.default_rng(seed)
    from scipy.ndimage import gaussian_filter
    out = np.empty((n, h, w), dtype=np.uint8)
    for i in range(n):
        a = rng.integers(0, 256, size=(h, w)).astype(np.float32)
        b = gaussian_filter(a, 1.2, mode="mirror")
        b = (b - b.mean()) / (b.std() + 1e-9)
        img = 127.5 + 60.0 * b
        # a few hundred random bright / dark rectangles give strong corners at several scales
        for _ in range(160):
            x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
            ww, hh = int(rng.integers(4, 60)), int(rng.integers(4, 60))
            img[y0:y0 + hh, x0:x0 + ww] += float(rng.uniform(-70, 70))
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out
