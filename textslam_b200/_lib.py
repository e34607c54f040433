"""Loader of the in-tree CUDA library (textslam_b200/libtslam_b200.so).

There is NO CPU fallback: if the library is missing it must be built (`python __graft_entry__.py`
or `make -C textslam_b200/csrc`), and every compute call raises if no sm_100 device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtslam_b200.so")
_LIB = None

# every symbol include/tslam_b200.h declares
EXPORTS = [
    "tslam_last_error", "tslam_version", "tslam_launch_count", "tslam_ctx_create", "tslam_ctx_destroy", "tslam_nccl_unique_id",
    "tslam_ctx_init_comm", "tslam_shard_owner", "tslam_eval_points", "tslam_eval_text", "tslam_solve", "tslam_dev_upload", "tslam_dev_free",
    "tslam_dev_eval_points", "tslam_dev_eval_text", "tslam_dev_lm_iterations", "tslam_dev_download_eval",
    "tslam_dev_download_params", "tslam_orb_create", "tslam_orb_destroy", "tslam_orb_extract", "tslam_orb_level_size",
    "tslam_orb_get_level", "tslam_orb_dev_bench", "tslam_orb_debug_get", "tslam_frame_pyr_create", "tslam_frame_pyr_destroy", "tslam_frame_pyr_build",
    "tslam_frame_pyr_level_size", "tslam_frame_pyr_get", "tslam_match_hamming", "tslam_search_from_3d", "tslam_search_in_area", "tslam_theta_covariance", "tslam_text_info", "tslam_analyze_structure", "tslam_debug_compare_analysis",
    "tslam_gate_residuals", "tslam_solve_gated", "tslam_debug_chol_schedule", "tslam_dev_chol_solve",
]


class TslamError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise TslamError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                             "(nvcc, sm_100a). textslam_b200 has no CPU fallback.")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.tslam_last_error.restype = C.c_char_p
        for name in EXPORTS:
            if name not in ("tslam_last_error", "tslam_ctx_destroy", "tslam_dev_free", "tslam_orb_destroy", "tslam_frame_pyr_destroy"):
                getattr(_LIB, name).restype = C.c_int
        _LIB.tslam_launch_count.restype = C.c_longlong
        for name in ("tslam_ctx_destroy", "tslam_dev_free", "tslam_orb_destroy", "tslam_frame_pyr_destroy"):
            getattr(_LIB, name).restype = None
    return _LIB


def check(rc):
    if rc != 0:
        raise TslamError(f"libtslam_b200 error {rc}: {lib().tslam_last_error().decode(errors='replace')}")
