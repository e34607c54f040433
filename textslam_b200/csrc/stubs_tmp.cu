#include "ctx.cuh"
extern "C" {
int tslam_orb_create(tslam_ctx*, int, float, int, int, int, int, tslam_orb**) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
void tslam_orb_destroy(tslam_orb*) {}
int tslam_orb_extract(tslam_orb*, const uint8_t* const*, int, int, int, int, int, tslam_keypoint*, uint8_t*, int32_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_level_size(tslam_orb*, int, int*, int*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_get_level(tslam_orb*, int, int, uint8_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
int tslam_orb_dev_bench(tslam_orb*, const uint8_t* const*, int, int, int, int, int, float*, int64_t*) { return tsl::set_error(TSLAM_ERR_ARG, "not built"); }
}
