"""Synthetic SearchFrom3D inputs: a frame with ORB-like keypoints (clustered, several octaves, some off the grid) and descriptors,
map points hosted in a few key frames whose projections land near keypoints (plus misses, bad points and out-of-view points)."""
import numpy as np
from textslam_b200 import synth


def make_case(seed, n_kp=1500, n_pts=600, n_hosts=4, width=640, height=480):
    rng = np.random.default_rng(seed)
    K = np.array([520.0, 521.0, 319.5, 239.5])
    kp_xy = np.stack([rng.uniform(-3, width + 3, n_kp), rng.uniform(-3, height + 3, n_kp)], 1).astype(np.float32)
    kp_xy[: n_kp // 5] = (kp_xy[0] + rng.normal(0, 6, (n_kp // 5, 2))).astype(np.float32)      # a dense cluster: long candidate lists, ties
    kp_oct = rng.choice([0, 0, 0, 1, 1, 2, 3, 5], n_kp).astype(np.int32)
    train = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    train[n_kp // 5: n_kp // 5 + 40] = train[0]                                               # identical descriptors: first candidate wins
    poses = np.stack([np.r_[synth.qexp(rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, 3)] for _ in range(n_hosts)])
    Tcw = np.r_[synth.qexp(rng.normal(0, 0.02, 3)), rng.normal(0, 0.3, 3)]
    pt_host = rng.integers(0, n_hosts, n_pts).astype(np.int32)
    # points constructed from a target pixel in the frame and a depth, expressed in the host frame
    tgt_k = rng.integers(0, n_kp, n_pts)
    tgt = kp_xy[tgt_k].astype(np.float64) + rng.normal(0, 4.0, (n_pts, 2))
    tgt[: n_pts // 10] = rng.uniform(-200, 900, (n_pts // 10, 2))                             # some outside the image
    depth = rng.uniform(2.0, 12.0, n_pts)
    pt_ray = np.zeros((n_pts, 2)); pt_rho = np.zeros(n_pts)
    for i in range(n_pts):
        pc = depth[i] * np.array([(tgt[i, 0] - K[2]) / K[0], (tgt[i, 1] - K[3]) / K[1], 1.0])
        Xw = synth.qrot(synth.qconj(Tcw[:4]), pc - Tcw[4:])
        P = poses[pt_host[i]]
        pr = synth.qrot(P[:4], Xw) + P[4:]
        pt_ray[i] = pr[:2] / pr[2]; pt_rho[i] = 1.0 / pr[2]
    n_query = n_pts + 50
    query = rng.integers(0, 256, (n_query, 32), dtype=np.uint8)
    pt_query = rng.permutation(n_query)[:n_pts].astype(np.int32)
    pt_query[rng.choice(n_pts, n_pts // 12, replace=False)] = -1                              # FLAG_BAD / not observed in the last key frame
    pt_query[5] = pt_query[6] if pt_query[6] >= 0 else 0                                      # two points sharing an observation index
    near = rng.choice(n_pts, n_pts // 2, replace=False)                                       # make half of the queries resemble a keypoint
    for i in near:
        if pt_query[i] >= 0:
            k = tgt_k[i]                                                                      # ... the one its projection lands next to
            query[pt_query[i]] = train[k] ^ (rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8))
    return dict(Tcw=Tcw, K=K, pt_ray=pt_ray, pt_rho=pt_rho, poses=poses, pt_host=pt_host, pt_query=pt_query, query_desc=query, kp_xy=kp_xy,
                kp_octave=kp_oct, train_desc=train, width=width, height=height)
