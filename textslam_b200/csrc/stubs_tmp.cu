#include "ctx.cuh"
namespace tsl { void free_solver(tslam_dev_problem* d) {} }
extern "C" {
void tslam_comm_destroy(tslam_ctx*) {}
int tslam_nccl_unique_id(uint8_t*) { return tsl::set_error(TSLAM_ERR_NCCL, "not built"); }
int tslam_ctx_init_comm(tslam_ctx*, int, int, const uint8_t*) { return tsl::set_error(TSLAM_ERR_NCCL, "not built"); }
int tslam_solve(tslam_ctx*, tslam_ba_problem*, const tslam_solve_options*, tslam_solve_summary*, double*, double*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_dev_lm_iterations(tslam_ctx*, tslam_dev_problem*, const tslam_solve_options*, int, float*, tslam_solve_summary*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_create(tslam_ctx*, int, float, int, int, int, int, tslam_orb**) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
void tslam_orb_destroy(tslam_orb*) {}
int tslam_orb_extract(tslam_orb*, const uint8_t* const*, int, int, int, int, int, tslam_keypoint*, uint8_t*, int32_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_level_size(tslam_orb*, int, int*, int*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_get_level(tslam_orb*, int, int, uint8_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_dev_bench(tslam_orb*, const uint8_t* const*, int, int, int, int, int, float*, int64_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
}
