"""GPU parity: residual + Jacobian kernels (through the C-ABI) vs the CPU oracle on the same inputs.
Tolerances: FP64 kernels with FMA contraction vs an un-contracted CPU restatement; closed-form
Jacobians vs dual numbers -> relative 1e-9 of the block scale."""
import numpy as np
import pytest
from textslam_b200 import synth
from textslam_b200._abi import (PT_BA, PT_BA_NW, PT_POSE, PT_RHO, TX_BA, TX_POSE, TX_THETA, JAC_ANALYTIC,
                                JAC_CENTRAL_DIFF, PT_NCOLS, TX_NCOLS)

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, name):
    scale = np.abs(b).reshape(len(b), -1).max(1) + 1.0
    err = np.abs(a - b).reshape(len(b), -1).max(1) / scale
    assert err.max() < rtol, f"{name}: max rel err {err.max():.3e} at block {err.argmax()}"


@pytest.mark.parametrize("kind", [PT_BA, PT_BA_NW, PT_POSE, PT_RHO])
def test_points_parity(ctx, oracle, kind):
    prob = synth.c4_local_ba(seed=11, n_planes=0)
    ro, Jo = oracle.eval_points(prob, kind)
    rg, Jg = ctx.eval_points(prob, kind)
    _close(rg, ro, 1e-11, "r")
    _close(Jg, Jo, 1e-9, "J")
    rg2, _ = ctx.eval_points(prob, kind, want_J=False)
    assert np.array_equal(rg2, rg)


def test_points_ragged_sizes(ctx, oracle):
    for n_lm in (1, 43, 129):   # not multiples of the CTA size, single observation
        prob = synth.make_ba_problem(seed=n_lm, n_kf=4, n_lm=n_lm, obs_per_lm=1, band=4, fixed_cams=(0,))
        ro, Jo = oracle.eval_points(prob, PT_BA)
        rg, Jg = ctx.eval_points(prob, PT_BA)
        _close(rg, ro, 1e-11, "r"); _close(Jg, Jo, 1e-9, "J")


def test_points_c5_shape(ctx, oracle):
    prob = synth.c5_global_ba(seed=0)
    assert prob.n_pobs == 100000 and len(prob.cams) == 500
    ro, Jo = oracle.eval_points(prob, PT_BA_NW, n_threads=8)
    rg, Jg = ctx.eval_points(prob, PT_BA_NW)
    _close(rg, ro, 1e-11, "r"); _close(Jg, Jo, 1e-9, "J")


@pytest.mark.parametrize("kind", [TX_BA, TX_POSE, TX_THETA])
@pytest.mark.parametrize("mode", [JAC_ANALYTIC, JAC_CENTRAL_DIFF])
def test_text_parity(ctx, oracle, kind, mode):
    prob = synth.c4_local_ba(seed=12, n_lm=30)
    ro, Jo = oracle.eval_text(prob, kind, mode)
    rg, Jg = ctx.eval_text(prob, kind, mode)
    _close(rg, ro, 1e-10, "r")
    if mode == JAC_ANALYTIC:
        _close(Jg, Jo, 1e-9, "J")
    else:
        # central differences amplify rounding by 1/h ~ 1e8 (h = sqrt(eps) floor): compare at that noise level
        scale = np.abs(Jo).reshape(len(Jo), -1).max(1) + 1.0
        err = np.abs(Jg - Jo).reshape(len(Jo), -1).max(1) / scale
        assert np.median(err) < 1e-6 and (err < 1e-4).mean() > 0.99, (np.median(err), err.max())


def test_text_edge_cases(ctx, oracle):
    prob = synth.c4_local_ba(seed=13, n_lm=20, n_planes=4)
    prob.t_musigma[:25, 1] = 0.0
    prob.t_rays[25:50] += 50.0
    ro, Jo = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC)
    rg, Jg = ctx.eval_text(prob, TX_BA, JAC_ANALYTIC)
    assert np.all(rg[:25] == 0) and np.all(Jg[:25] == 0) and np.all(Jg[25:50] == 0)
    _close(rg, ro, 1e-10, "r"); _close(Jg, Jo, 1e-9, "J")


def test_device_resident_eval_matches(ctx, oracle):
    prob = synth.c4_local_ba(seed=14)
    d = ctx.upload(prob)
    ms = d.eval_points(PT_BA, reps=3, flush_l2=True)
    assert ms > 0
    r, J = d.download_eval(0, 13)
    ro, Jo = oracle.eval_points(prob, PT_BA)
    _close(r, ro, 1e-11, "r"); _close(J, Jo, 1e-9, "J")
    d.eval_text(TX_BA, JAC_ANALYTIC, reps=2)
    r, J = d.download_eval(1, 15)
    ro, Jo = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC)
    _close(r, ro, 1e-10, "r"); _close(J, Jo, 1e-9, "J")
    d.free()


def test_bad_arguments_fail_loudly(ctx):
    import textslam_b200 as T
    prob = synth.c4_local_ba(seed=15, n_lm=10, n_planes=0)
    prob.p_cam[0] = 999
    with pytest.raises(T.TslamError):
        ctx.eval_points(prob, PT_BA)


def test_text_tma_staged_variant_is_identical(ctx):
    """TSLAM_JAC_ANALYTIC_TMA: the image window of every text object staged in shared memory by one cp.async.bulk.tensor.2d
    (csrc/ba_eval_tma.cu) — same residuals and Jacobians as the __ldg-tap kernel (to rounding) on the host-buffer entry point and
    on the device-resident one (C4 local BA, pose-only, a text-on global BA with objects near the image border)."""
    import textslam_b200 as T
    for prob, kind in ((synth.c4_local_ba(seed=81), T.TX_BA), (synth.c3_pose_only(seed=82), T.TX_POSE),
                       (synth.c5_global_ba(seed=83, n_kf=60, n_lm=500, n_planes=120, text_kf_stride=2), T.TX_BA),
                       (synth.c4_local_ba(seed=84), T.TX_THETA)):
        r0, J0 = ctx.eval_text(prob, kind, T.JAC_ANALYTIC)
        r1, J1 = ctx.eval_text(prob, kind, T.JAC_ANALYTIC_TMA)
        # same tap values, same formulas; the compiler shares sub-expressions with the window computation of the staged kernel, so
        # a few products are contracted differently: agreement to rounding, not bit for bit
        tol = lambda x: 1e-12 * (np.abs(x).max() + 1.0)
        assert np.abs(r0 - r1).max() <= tol(r0) and np.abs(J0 - J1).max() <= tol(J0), (np.abs(r0 - r1).max(), np.abs(J0 - J1).max())
        d = ctx.upload(prob)
        d.eval_text(kind, T.JAC_ANALYTIC_TMA, reps=2)
        r2, J2 = d.download_eval(1, J0.shape[-1])
        assert np.array_equal(r1, r2) and np.array_equal(J1, J2)
        d.free()
