"""Programmatic dependent launch must not change results: the same solves in a child process with TSLAM_PDL=0 (plain stream
order, the griddepcontrol instructions are no-ops) must reproduce this process' results bit for bit (all reductions are
order-deterministic)."""
import json
import os
import subprocess
import sys
import numpy as np
import pytest
from textslam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
out = {}
for name, prob, its in (("c4", synth.c4_local_ba(seed=31, n_lm=400, n_planes=6), 10),
                        ("gba", synth.c5_global_ba(seed=32, n_kf=200, n_lm=6000), 6)):
    summ, fr, _ = ctx.solve(prob, its)
    out[name] = {"iterations": summ["iterations"], "final_cost": summ["final_cost"].hex(),
                 "params": hashlib.sha256(prob.cams.tobytes() + prob.rho.tobytes() + prob.theta.tobytes()).hexdigest(),
                 "resid": hashlib.sha256(fr.tobytes()).hexdigest()}
print("RESULT " + json.dumps(out))
"""


def _run(pdl):
    env = dict(os.environ, TSLAM_PDL=pdl)
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_pdl_on_and_off_are_bit_identical():
    on, off = _run("1"), _run("0")
    assert on == off
    assert on["gba"]["iterations"] >= 3


def test_two_tile_panels_match_single_tile_schedule():
    """TSLAM_CHOL_PAIR=1 factors the two tiles of a layout node in one CTA (potrf2_trsm2_kernel): same iteration sequence,
    results equal to rounding (the trailing updates are summed in a different order)."""
    base = _run("1")
    env = dict(os.environ, TSLAM_CHOL_PAIR="1")
    p = subprocess.run([sys.executable, "-c", CHILD.replace('"final_cost": summ["final_cost"].hex()', '"final_cost": summ["final_cost"].hex(), "cams": prob.cams.tolist()') % ROOT],
                       env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    pair = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    p0 = subprocess.run([sys.executable, "-c", CHILD.replace('"final_cost": summ["final_cost"].hex()', '"final_cost": summ["final_cost"].hex(), "cams": prob.cams.tolist()') % ROOT],
                        env=dict(os.environ, TSLAM_CHOL_PAIR="0"), capture_output=True, text=True, timeout=300)
    single = json.loads([l for l in p0.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    for name in ("c4", "gba"):
        assert pair[name]["iterations"] == single[name]["iterations"] == base[name]["iterations"]
        a, b = np.array(pair[name]["cams"]), np.array(single[name]["cams"])
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
        fa, fb = float.fromhex(pair[name]["final_cost"]), float.fromhex(single[name]["final_cost"])
        assert abs(fa - fb) <= 1e-10 * abs(fb)
