"""textslam_b200 — B200-native (sm_100a) hot path of TextSLAM behind a C-ABI (include/tslam_b200.h)."""
from ._abi import (BAProblem, PT_BA, PT_BA_NW, PT_POSE, PT_RHO, TX_BA, TX_POSE, TX_THETA, JAC_ANALYTIC,  # noqa: F401
                   JAC_CENTRAL_DIFF, JAC_ANALYTIC_TMA, KP_DTYPE, gate_options)
from ._lib import TslamError  # noqa: F401
from .api import Context, DeviceProblem, Optimizer, PyramidLevel, run_pyramid, ORBextractor, FramePyramid, match_hamming, text_info, analyze_structure, FrameGrid, search_from_3d, search_in_area, resolve_matches, resolve_local_track  # noqa: F401
