// Internal interfaces between the solver translation units.
#pragma once
#include "ctx.cuh"
#include "analysis.hpp"

namespace tsl {

// chol.cu
struct CholSymbolic {  // tile-level symbolic factorisation + level schedule (per problem structure)
  int Tn = 0, n = 0, nwaves = 0;
  std::vector<int> item_ptr, item2_ptr, target_ptr, panel_ptr;   // per wave ranges into the device lists
  DevBuf<int2> items, items2, targets, clear_items;   // single-tile panels (panel, row tile | -1), two-tile panels, (target tile i, k), all pattern tiles
  int n_clear = 0;
  DevBuf<int> src_ptr, src;                           // per target: source panels of its wave
  DevBuf<int> panels, below_ptr, below;               // backward solve: panels per wave and their non-zero tiles below
  DevBuf<int> pair_a; int n_pairs = 0;                // first tiles of the two-tile panels
  mutable DevBuf<double> Lpair;                       // per two-tile panel: L_ba parked until every CTA of its launch has read A_ba
  mutable DevBuf<double> Ldiag;                       // Tn inverse diagonal factors L_jj^-1 (64x64, tight)
  mutable DevBuf<int> flags;                          // backward solve: flags[j] == epoch <=> x_j final in the current call
  mutable int epoch = 0;
  long long gemm_tiles = 0;                           // number of 64x64x64 tile updates (2*64^3 flop each)
  // fused schedule (chol_fused.cu, chol_sched.hpp)
  int f_ntasks = 0, f_nsync = 0;
  DevBuf<int> f_tasks, f_srcs, f_below;
  DevBuf<int2> f_deps, f_dep_inl;
  DevBuf<int> f_idx_inl;
  mutable DevBuf<int> f_sync;                         // queue head + dependency counters, zeroed before every solve
};
int chol_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym);   // device copy of the host symbolic factorisation (analysis.cpp)
int chol_clear(tslam_ctx* ctx, const CholSymbolic& sym, double* A);
int chol_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* ywork, double* xout, int* d_fail);
int chol_solve_waves(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* xout, int* d_fail);
bool chol_fused_enabled();   // TSLAM_CHOL_FUSED != "0"
// chol_fused.cu
int chol_fused_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym);
int chol_fused_clear(tslam_ctx* ctx, const CholSymbolic& sym);
int chol_fused_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* xout, int* d_fail, unsigned long long* trace);

// Index structures of one problem on the device (what the structure analysis produces; analysis.cpp on the host or
// analysis_dev.cu on the GPU) — the LM kernels of ba_solve.cu only ever read these.
struct SolverIndex {
  int K = 0, nc = 0, nl = 0, npl = 0;         // cameras; free cameras / inverse depths / planes (global)
  int lp = 0, lt = 0;                         // local observations
  int nvp = 0, nvt = 0, nsp = 0, nst = 0;     // owned landmarks / (landmark, camera) slots
  int nblk = 0, noff = 0;                     // non-zero 6x6 blocks (a <= b), of which off-diagonal
  long long est_entries = 0;                  // gather-list entries over all blocks (an upper estimate on the device path)
  int n = 0, npad = 0, ld = 0, rows = 0, Tn = 0;   // reduced system: 6 nc unknowns, columns of the tile-aligned layout, workspace dims
  std::vector<int> camslot;                   // host copies (multi-GPU result merge only)
  std::vector<int> vp_gl_h, vt_gl_h, lmfree_p_h, lmfree_t_h;
  DevBuf<int> doff;                           // first column of each camera slot in the dense reduced matrix (nd_layout.h)
  DevBuf<int> camslot_d, p_cs, p_hs, p_ls, t_cs, t_hs, t_ls;
  DevBuf<uint8_t> p_active, t_active, t_fmask;
  DevBuf<int> vp_gl, vt_gl, vp_obs_ptr, vp_obs, vt_obs_ptr, vt_obs;
  DevBuf<int> sp_ptr, sp_cam, sp_lm, spe_ptr, spe, st_ptr, st_cam, st_lm, ste_ptr, ste;
  DevBuf<int> blk_a, blk_b, diag_blk, offdiag_blk, bdp_ptr, bdp, bdt_ptr, bdt, bsp_ptr, bst_ptr, gsel_p, gsel_t;
  DevBuf<int2> bsp, bst;
  CholSymbolic chol;
};
// analysis_dev.cu: the same analysis as analyze_structure(), for an unsharded problem, entirely on the device (radix sorts
// and scans instead of counting sorts); returns TSLAM_ERR_ARG + "unsupported" when the problem exceeds its packing limits.
bool device_analysis_supported(const tslam_ctx* ctx, const tslam_dev_problem* d);
int analyze_structure_device(tslam_ctx* ctx, tslam_dev_problem* d, SolverIndex& X, double* lap_ms /*[4]*/);

// ba_eval.cu (robustified evaluation used inside the LM loop)
int launch_eval_points_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* rho, const uint8_t* active,
                              double* r, double* J, double* cost_part, int* n_parts);
int launch_eval_text_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* theta, const uint8_t* active,
                            const uint8_t* free_masks, int jac_mode, double* r, double* J, double* cost_part, int* n_parts);

// ba_small.cu: one persistent kernel per solve for pose-only / local-window problems; *handled = false -> not eligible
int small_solve(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, tslam_solve_summary* summary, double* final_residuals,
                double* trace, const double** d_rp, const double** d_rt, bool* handled);
void small_workspace_free(tslam_ctx* ctx);

// comm.cu
int comm_allreduce_sum(tslam_ctx* ctx, double* buf, size_t n);
int comm_allreduce_max(tslam_ctx* ctx, double* buf, size_t n);
int comm_allreduce_sum_i32(tslam_ctx* ctx, int* buf, size_t n);

}  // namespace tsl
