// Internal interfaces between the solver translation units.
#pragma once
#include "ctx.cuh"

namespace tsl {

// chol.cu
int chol_workspace_dims(int n, int* ld, int* rows);
int chol_solve(tslam_ctx* ctx, double* A, int n, double* ywork, double* xout, int* d_fail);

// ba_eval.cu (robustified evaluation used inside the LM loop)
int launch_eval_points_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* rho, const uint8_t* active,
                              double* r, double* J, double* cost_part, int* n_parts);
int launch_eval_text_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* theta, const uint8_t* active,
                            const uint8_t* free_masks, int jac_mode, double* r, double* J, double* cost_part, int* n_parts);

// comm.cu
int comm_allreduce_sum(tslam_ctx* ctx, double* buf, size_t n);
int comm_allreduce_max(tslam_ctx* ctx, double* buf, size_t n);

}  // namespace tsl
