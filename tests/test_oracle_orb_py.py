"""Second, independent restatement of the keypoint stage of ORBextractor (src/ORBextractor.cc:540-854) in plain Python on top
of the real cv2.FAST / cv2.resize — list manipulation exactly as the reference writes it (std::list push_front / erase), no code
shared with oracle/orb_oracle.cpp. The C++ oracle must reproduce its candidates (vToDistributeKeys order) and its quad-tree
winners (list order) level by level. The one documented freedom, the (size, node*) sort of the fine phase, is resolved the same
way in both: equal sizes are ordered by creation (a later node counts as the larger pointer)."""
import math
import numpy as np
import pytest
from textslam_b200 import synth

cv2 = pytest.importorskip("cv2")
F32 = np.float32


def pyramid(img, oracle):
    out, lvl = [img], img
    for l in range(1, 8):
        w, h = oracle.orb_level_size(img.shape[1], img.shape[0], 1.2, 8, l)
        lvl = cv2.resize(lvl, (w, h), interpolation=cv2.INTER_LINEAR)   # each level from the previous one (:1131)
        out.append(lvl)
    return out


def candidates(im, ini_th=20, min_th=7):
    """ComputeKeyPointsOctTree :772-830 for one level; coordinates relative to (minBorderX, minBorderY)."""
    minBX = minBY = 19 - 3
    maxBX, maxBY = im.shape[1] - 19 + 3, im.shape[0] - 19 + 3
    width, height = F32(maxBX - minBX), F32(maxBY - minBY)
    nCols, nRows = int(width / F32(30)), int(height / F32(30))
    wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
    det = {t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for t in (ini_th, min_th)}
    keys = []
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
            kps = det[ini_th].detect(roi)
            if not kps:
                kps = det[min_th].detect(roi)
            for k in kps:
                keys.append((int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)))
    return keys, (minBX, maxBX, minBY, maxBY)


class Node:
    seq_counter = 0

    def __init__(self, ul, ur, bl, br):
        self.UL, self.UR, self.BL, self.BR = ul, ur, bl, br
        self.keys, self.no_more = [], False
        Node.seq_counter += 1
        self.seq = Node.seq_counter

    def divide(self):
        halfX = int(math.ceil(F32(self.UR[0] - self.UL[0]) / 2))
        halfY = int(math.ceil(F32(self.BR[1] - self.UL[1]) / 2))
        ul = self.UL
        n1 = Node(ul, (ul[0] + halfX, ul[1]), (ul[0], ul[1] + halfY), (ul[0] + halfX, ul[1] + halfY))
        n2 = Node(n1.UR, self.UR, n1.BR, (self.UR[0], ul[1] + halfY))
        n3 = Node(n1.BL, n1.BR, self.BL, (n1.BR[0], self.BL[1]))
        n4 = Node(n3.UR, n2.BR, n3.BR, self.BR)
        for kp in self.keys:
            if kp[0] < n1.UR[0]:
                (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
            elif kp[1] < n1.BR[1]:
                n2.keys.append(kp)
            else:
                n4.keys.append(kp)
        for n in (n1, n2, n3, n4):
            if len(n.keys) == 1:
                n.no_more = True
        return n1, n2, n3, n4


def distribute(keys, box, N):
    """DistributeOctTree :540-764 with a Python list as the std::list (index 0 = front)."""
    minX, maxX, minY, maxY = box
    nIni = int(round(F32(maxX - minX) / F32(maxY - minY)))
    hX = F32(maxX - minX) / F32(nIni)
    nodes = []
    for i in range(nIni):
        ulx, urx = int(hX * F32(i)), int(hX * F32(i + 1))
        nodes.append(Node((ulx, 0), (urx, 0), (ulx, maxY - minY), (urx, maxY - minY)))
    ini = list(nodes)
    for kp in keys:
        ini[int(F32(kp[0]) / hX)].keys.append(kp)
    kept = []
    for n in nodes:
        if len(n.keys) == 1:
            n.no_more = True
            kept.append(n)
        elif n.keys:
            kept.append(n)
    nodes = kept
    finish = False
    while not finish:
        prev_size = len(nodes)
        n_expand, vsize = 0, []
        for n in list(nodes):                     # children go to the front: this pass never visits them
            if n.no_more:
                continue
            for c in n.divide():
                if c.keys:
                    nodes.insert(0, c)
                    if len(c.keys) > 1:
                        n_expand += 1
                        vsize.append(c)
            nodes.remove(n)
        if len(nodes) >= N or len(nodes) == prev_size:
            finish = True
        elif len(nodes) + n_expand * 3 > N:
            while not finish:
                prev_size = len(nodes)
                prev = sorted(vsize, key=lambda n: (len(n.keys), n.seq))
                vsize = []
                for n in reversed(prev):
                    for c in n.divide():
                        if c.keys:
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                vsize.append(c)
                    nodes.remove(n)
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev_size:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for kp in n.keys[1:]:
            if kp[2] > best[2]:
                best = kp
        out.append(best)
    return out


@pytest.mark.parametrize("seed", [32, 5])
def test_keypoint_stage_matches_python_restatement(oracle, seed):
    img = synth.orb_images(seed=seed, n=1)[0]
    per_level = oracle.orb_features_per_level(1000, 1.2, 8)
    pyr = pyramid(img, oracle)
    for level in range(8):
        keys, box = candidates(pyr[level])
        co = oracle.orb_debug(img, 1, level)
        assert len(keys) == len(co) and np.array_equal(np.array(keys, dtype=np.int32).reshape(-1, 3), co), f"candidates differ at level {level}"
        win = distribute(keys, box, int(per_level[level]))
        so = oracle.orb_debug(img, 2, level)
        assert np.array_equal(np.array(win, dtype=np.int32).reshape(-1, 3), so), f"quad-tree winners differ at level {level}"
        assert len(win) >= per_level[level] or len(win) == len(set(keys))


def umax_table():
    """ORBextractor ctor :453-470."""
    HP = 15
    umax = [0] * (HP + 1)
    vmax = int(math.floor(HP * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(HP * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(round(math.sqrt(HP * HP - v * v)))       # cvRound: no exact .5 occurs here
    v0 = 0
    for v in range(HP, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def ic_angle(im, x, y, umax):
    """IC_Angle :77-104 with cv2.fastAtan2."""
    c = im.astype(np.int64)
    m10 = sum(u * int(c[y, x + u]) for u in range(-15, 16))
    m01 = 0
    for v in range(1, 16):
        d = umax[v]
        plus, minus = c[y + v, x - d:x + d + 1], c[y - v, x - d:x + d + 1]
        u = np.arange(-d, d + 1)
        m01 += v * int((plus - minus).sum())
        m10 += int((u * (plus + minus)).sum())
    return cv2.fastAtan2(float(m01), float(m10))


def test_orientation_and_output_order_match_python_restatement(oracle):
    """operator() :1054-1116 up to the keypoint list: level-major order, coordinates = (winner + border) * level scale, size = int(31 * scale),
    angle = IC_Angle on the (unblurred) level image."""
    img = synth.orb_images(seed=32, n=1)[0]
    kp, _ = oracle.orb_extract(img)
    per_level = oracle.orb_features_per_level(1000, 1.2, 8)
    pyr, umax = pyramid(img, oracle), umax_table()
    assert umax == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    pos = 0
    scale = F32(1.0)
    for level in range(8):
        keys, box = candidates(pyr[level])
        win = distribute(keys, box, int(per_level[level]))
        seg = kp[pos:pos + len(win)]
        assert len(seg) == len(win) and np.all(seg["octave"] == level)
        for k, (x, y, resp) in zip(seg, win):
            xl, yl = x + 16, y + 16
            assert k["response"] == resp
            assert k["x"] == F32(xl) * scale and k["y"] == F32(yl) * scale if level else (k["x"] == xl and k["y"] == yl)
            assert k["size"] == int(F32(31.0) * scale)            # const int scaledPatchSize = PATCH_SIZE * mvScaleFactor[level] (:835)
            assert k["angle"] == F32(ic_angle(pyr[level], xl, yl, umax)), (level, xl, yl)
        pos += len(win)
        scale = scale * F32(1.2)          # mvScaleFactor chain in float32 (:416-420)
    assert pos == len(kp)


def _reference_pattern():
    import os
    import re
    path = "/root/reference/src/ORBextractor.cc"
    if not os.path.exists(path):
        pytest.skip("reference tree not present (the rBRIEF table is read from it to stay independent of the product's copy)")
    src = open(path).read()
    body = src[src.index("static int bit_pattern_31_[256*4]"):]
    body = re.sub(r"/\*.*?\*/", "", body[body.index("{") + 1: body.index("};")], flags=re.S)
    nums = np.array([int(v) for v in re.findall(r"-?\d+", body)], dtype=np.int32)
    assert nums.size == 1024
    return nums.reshape(512, 2)          # 512 points (x, y): 16 per descriptor byte


def test_descriptors_match_python_restatement(oracle):
    """computeOrbDescriptor :106-147 on cv2.GaussianBlur(7x7, sigma 2) of each level, pattern table parsed from the reference source,
    steering in float32 (cos / sin evaluated in double and rounded to float), cvRound = round half to even."""
    pat = _reference_pattern()
    img = synth.orb_images(seed=32, n=1)[0]
    kp, desc = oracle.orb_extract(img)
    pyr = pyramid(img, oracle)
    factor_pi = F32(np.pi / F32(180.0))
    px, py = pat[:, 0].astype(F32), pat[:, 1].astype(F32)
    scale, pos, n_bad = F32(1.0), 0, 0
    for level in range(8):
        seg = kp[kp["octave"] == level]
        blur = cv2.GaussianBlur(pyr[level], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        for k in seg:
            xl, yl = int(round(float(k["x"] / scale))), int(round(float(k["y"] / scale)))
            ang = F32(k["angle"]) * factor_pi
            a, b = F32(math.cos(float(ang))), F32(math.sin(float(ang)))
            rr = np.rint((px * b + py * a).astype(np.float64)).astype(np.int64)     # row offsets (x*b + y*a)
            cc = np.rint((px * a - py * b).astype(np.float64)).astype(np.int64)     # column offsets (x*a - y*b)
            vals = blur[yl + rr, xl + cc].astype(np.int32).reshape(32, 8, 2)
            bits = (vals[:, :, 0] < vals[:, :, 1]).astype(np.uint8)
            byte = (bits << np.arange(8, dtype=np.uint8)).sum(1).astype(np.uint8)
            n_bad += int(not np.array_equal(byte, desc[pos]))
            pos += 1
        scale = scale * F32(1.2)
    assert pos == len(kp)
    assert n_bad == 0, f"{n_bad} of {len(kp)} descriptors differ"
