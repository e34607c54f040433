"""Stage-by-stage GPU parity of the ORB extractor vs the oracle: FAST measure plane, per-level candidates
(vToDistributeKeys order) and quad-tree winners (list order)."""
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_stages_match_oracle(ctx, oracle):
    import textslam_b200 as T
    imgs = synth.orb_images(seed=32, n=2)
    orb = T.ORBextractor(ctx, 1000, 1.2, 8, 20, 7)
    orb.extract_batch(imgs)
    problems = []
    for i in range(2):
        for l in (0, 3, 7):
            mg, mo = orb.debug_get(0, i, l), oracle.orb_debug(imgs[i], 0, l)
            mg = np.where(mg > 7, mg, 0)          # the device plane keeps sub-threshold measures, the oracle plane zeroes them
            if not np.array_equal(mg, mo):
                bad = np.argwhere(mg != mo)
                problems.append(("measure", i, l, len(bad), bad[:3].tolist(), [(int(mg[y, x]), int(mo[y, x])) for y, x in bad[:3]]))
            cg, co = orb.debug_get(1, i, l), oracle.orb_debug(imgs[i], 1, l)
            if cg.shape != co.shape or not np.array_equal(cg, co):
                sg_, so_ = set(map(tuple, cg.tolist())), set(map(tuple, co.tolist()))
                problems.append(("candidates", i, l, cg.shape, co.shape, "same set" if sg_ == so_ else (len(sg_ - so_), len(so_ - sg_)),
                                 cg[:3].tolist(), co[:3].tolist()))
            sg, so = orb.debug_get(2, i, l), oracle.orb_debug(imgs[i], 2, l)
            if sg.shape != so.shape or not np.array_equal(sg, so):
                a, b = set(map(tuple, sg.tolist())), set(map(tuple, so.tolist()))
                problems.append(("winners", i, l, sg.shape, so.shape, "same set" if a == b else (len(a - b), len(b - a)), sg[:4].tolist(), so[:4].tolist()))
    orb.close()
    assert not problems, "\n".join(map(str, problems))
