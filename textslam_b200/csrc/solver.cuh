// Internal interfaces between the solver translation units.
#pragma once
#include "ctx.cuh"
#include "analysis.hpp"

namespace tsl {

// chol.cu
struct CholSymbolic {  // tile-level symbolic factorisation + level schedule (per problem structure)
  int Tn = 0, n = 0, nwaves = 0;
  std::vector<int> item_ptr, target_ptr, panel_ptr;   // per wave ranges into the device lists
  DevBuf<int2> items, targets;                        // (panel, row tile | -1), (target tile i, k)
  DevBuf<int> src_ptr, src;                           // per target: source panels of its wave
  DevBuf<int> panels, below_ptr, below;               // backward solve: panels per wave and their non-zero tiles below
  mutable DevBuf<double> Ldiag;                       // Tn inverse diagonal factors L_jj^-1 (64x64, tight)
  mutable DevBuf<int> flags;                          // backward solve: flags[j] == epoch <=> x_j final in the current call
  mutable int epoch = 0;
  long long gemm_tiles = 0;                           // number of 64x64x64 tile updates (2*64^3 flop each)
};
int chol_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym);   // device copy of the host symbolic factorisation (analysis.cpp)
int chol_clear(tslam_ctx* ctx, const CholSymbolic& sym, double* A);
int chol_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* ywork, double* xout, int* d_fail);

// ba_eval.cu (robustified evaluation used inside the LM loop)
int launch_eval_points_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* rho, const uint8_t* active,
                              double* r, double* J, double* cost_part, int* n_parts);
int launch_eval_text_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* theta, const uint8_t* active,
                            const uint8_t* free_masks, int jac_mode, double* r, double* J, double* cost_part, int* n_parts);

// comm.cu
int comm_allreduce_sum(tslam_ctx* ctx, double* buf, size_t n);
int comm_allreduce_max(tslam_ctx* ctx, double* buf, size_t n);

}  // namespace tsl
