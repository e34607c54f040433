/* tslam_b200.h — C-ABI of libtslam_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the numerical hot path of SJTU-ViSYS/TextSLAM. The reference has no FFI
 * layer; its boundary is two C++ classes whose bodies call Ceres / OpenCV:
 *     TextSLAM::optimizer      /root/reference/src/optimizer.h:52-135
 *     TextSLAM::ORBextractor   /root/reference/src/ORBextractor.h:45-114
 * A maintainer keeps those class surfaces and replaces the Ceres / OpenCV calls inside
 * optimizer.cc / ORBextractor.cc with the entry points below (see INTEGRATION.md for the shim).
 *
 * Conventions: plain C, caller-owned HOST memory for every pointer in the public structs,
 * SoA layouts, int return (0 = ok, <0 = error; text via tslam_last_error()), every call blocks
 * until its outputs are in host memory. One context per host thread. No CPU fallback: if no
 * CUDA device / kernel image is usable the call fails with TSLAM_ERR_CUDA.
 */
#ifndef TSLAM_B200_H_
#define TSLAM_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSLAM_OK 0
#define TSLAM_ERR_ARG (-1)
#define TSLAM_ERR_CUDA (-2)
#define TSLAM_ERR_NUMERIC (-3)
#define TSLAM_ERR_NCCL (-4)

/* ---- residual kinds ---------------------------------------------------------------------- */
/* point functors: include/auto_BAScene.h:89-92, auto_BASceneNW.h, auto_PoseOptimScene.h:90-93,
 * auto_RhoScene.h:66-69 */
#define TSLAM_PT_BA 0      /* J 2x13: [d_c(3) t_c(3) d_h(3) t_h(3) rho], weighted   */
#define TSLAM_PT_BA_NW 1   /* same, weights forced to 1                              */
#define TSLAM_PT_POSE 2    /* J 2x6 : [d_c t_c]                                      */
#define TSLAM_PT_RHO 3     /* J 2x1 : [rho], weights forced to 1                     */
/* text functors: include/nume_BAText.h:97-100, nume_PoseOptimText.h:81-84, nume_thetaText.h:77-80 */
#define TSLAM_TX_BA 0      /* J 8x15: [d_c t_c d_h t_h theta(3)]                     */
#define TSLAM_TX_POSE 1    /* J 8x6                                                   */
#define TSLAM_TX_THETA 2   /* J 8x3, weight forced to 1                               */
/* text Jacobian mode */
#define TSLAM_JAC_ANALYTIC 0      /* closed form                             */
#define TSLAM_JAC_CENTRAL_DIFF 1  /* replica of ceres::NumericDiffCostFunction<CENTRAL> step rule */
#define TSLAM_JAC_ANALYTIC_TMA 2  /* closed form, image window of every text object staged in shared memory by a TMA tensor load
                                   * (tslam_eval_text / tslam_dev_eval_text only; same results as TSLAM_JAC_ANALYTIC)           */

/* ---- problem description (shared by eval and solve) ---------------------------------------- */
/* Replaces the ceres::Problem built in src/optimizer.cc:1106-1208 (pose), :1359-1588 (local BA),
 * :1716-1830 (global BA), :1869-1972 (landmarks), :2175-2200 (theta).
 * Keyframe poses that the reference passes as constant matrices (Trw / Twr / Tcr of
 * auto_PoseOptimScene, auto_RhoScene, nume_PoseOptimText, nume_thetaText) are entries of `cams`
 * with cam_fixed = 1; constant landmarks are entries with rho_fixed / theta_fixed = 1. */
typedef struct tslam_ba_problem {
  int32_t n_cams;
  double* cams;             /* n_cams x 7 : qw qx qy qz tx ty tz of T_cw (in/out for solve) */
  const uint8_t* cam_fixed; /* n_cams, 1 = SetParameterBlockConstant                         */
  int32_t n_points;
  double* rho;              /* n_points inverse depths (in/out)                               */
  const uint8_t* rho_fixed;
  int32_t n_planes;
  double* theta;            /* n_planes x 3 plane parameters (in/out)                         */
  const uint8_t* theta_fixed;

  /* point observations; order = AddResidualBlock order */
  int32_t n_pobs;
  const double* p_uv;       /* n_pobs x 2  level-0 observation (u,v)            */
  const double* p_ray;      /* n_pobs x 2  host-frame ray (x,y), z == 1         */
  const int32_t* p_cam;     /* observing keyframe                               */
  const int32_t* p_host;    /* host keyframe of the landmark                    */
  const int32_t* p_lm;      /* index into rho                                   */
  double K_point[4];        /* fx fy cx cy (level 0, src/optimizer.cc:1403)     */
  double w_point[2];        /* 1/1.2 (local/pose) or 1 (global)                 */
  double huber_point;       /* delta; <= 0 disables the loss                    */

  /* text feature blocks (8 pixel residuals each) */
  int32_t n_tobs;
  const double* t_rays;     /* n_tobs x 8 x 2 pattern rays (x,y), z == 1        */
  const double* t_iref;     /* n_tobs x 8 normalised reference intensities      */
  const double* t_musigma;  /* n_tobs x 2 (mu, sigma) of the current-image quad */
  const int32_t* t_cam;
  const int32_t* t_host;
  const int32_t* t_plane;   /* index into theta                                 */
  const int32_t* t_img;     /* index into imgs                                  */
  int32_t n_imgs, img_w, img_h;
  const uint8_t* imgs;      /* n_imgs x img_h x img_w, stride == img_w          */
  double K_text[4];         /* intrinsics of the pyramid level of imgs          */
  double w_text;            /* 1/0.2 or 1                                       */
  double huber_text;        /* delta; <= 0 disables                             */
} tslam_ba_problem;

typedef struct tslam_solve_options {
  int32_t max_iters;        /* ceres max_num_iterations (10 / 20 / 50)                     */
  int32_t text_jac_mode;    /* TSLAM_JAC_*                                                 */
  int32_t n_threads;        /* oracle only: host threads for the evaluation pass (>=1)     */
  int32_t dense_full;       /* oracle only: 1 = solve the full (cams+landmarks) system     */
  double function_tolerance;   /* <= 0 -> Ceres default 1e-6  */
  double gradient_tolerance;   /* <= 0 -> 1e-10               */
  double parameter_tolerance;  /* <= 0 -> 1e-8                */
  double initial_radius;       /* <= 0 -> 1e4                 */
} tslam_solve_options;

#define TSLAM_TERM_NO_CONVERGENCE 0
#define TSLAM_TERM_FUNCTION_TOL 1
#define TSLAM_TERM_PARAMETER_TOL 2
#define TSLAM_TERM_GRADIENT_TOL 3
#define TSLAM_TERM_FAILURE (-1)

typedef struct tslam_solve_summary {
  int32_t iterations;        /* LM iterations performed (excluding iteration 0)             */
  int32_t successful_steps;
  int32_t unsuccessful_steps;
  int32_t termination;       /* TSLAM_TERM_*                                                */
  double initial_cost, final_cost, fixed_cost;
  double total_ms;           /* wall time of the call                                       */
  double solve_ms;           /* device/CPU time inside the LM loop (excl. problem upload)   */
  double setup_ms;           /* structure analysis + upload                                 */
  int32_t n_free_cams, n_free_points, n_free_planes, reduced_dim;
} tslam_solve_summary;

/* trace row (optional, per iteration incl. 0): cost, radius, relative_decrease, accepted(0/1/-1) */
#define TSLAM_TRACE_COLS 4

/* ---- context ------------------------------------------------------------------------------ */
typedef struct tslam_ctx tslam_ctx;
const char* tslam_last_error(void);
int tslam_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py: gpu_launches) */
long long tslam_launch_count(void);
int tslam_ctx_create(int device_id, tslam_ctx** out);
void tslam_ctx_destroy(tslam_ctx* ctx);
/* Multi-GPU global BA: rank/world of a one-process-per-GPU job. `nccl_unique_id` is the 128-byte
 * ncclUniqueId produced by tslam_nccl_unique_id() on rank 0 and broadcast by the host program
 * (torch.distributed / MPI / sockets). */
int tslam_nccl_unique_id(uint8_t id_out[128]);
/* The sharding rule: rank that owns an observation. An observation lives with its landmark when the
 * landmark is free; observations of constant landmarks are dealt round-robin. Pure function, no GPU needed. */
int tslam_shard_owner(int landmark_is_free, int landmark_index, int obs_index, int world);
int tslam_ctx_init_comm(tslam_ctx* ctx, int rank, int world, const uint8_t nccl_unique_id[128]);

/* ---- structure analysis (host only, no GPU needed) -------------------------------------------------- */
/* What tslam_solve derives from the index arrays before the first kernel: free parameter blocks, the reduced camera
 * system (6 x 6 blocks), the landmark -> camera incidence of this rank's shard and the tile schedule of the
 * factorisation. The counterpart of ceres::Problem / Solver::Summary's num_*_reduced fields. */
typedef struct tslam_structure_info {
  int32_t n_free_cams, n_free_points, n_free_planes, reduced_dim;
  int32_t n_blocks;             /* non-zero 6x6 blocks of the upper triangle of the reduced camera matrix            */
  int32_t n_local_pobs, n_local_tobs;   /* observations owned by `rank`                                             */
  int32_t n_owned_points, n_owned_planes, n_slots_point, n_slots_text;   /* landmarks / (landmark, camera) pairs owned */
  int32_t n_tiles, n_waves;     /* 64 x 64 tile rows of the factorisation; length of its level schedule              */
  int64_t n_tile_updates;       /* 64^3 tile multiply-adds of one factorisation                                      */
  int64_t n_schur_entries, n_direct_entries;   /* gather-list lengths on this rank                                   */
  double analysis_ms;
} tslam_structure_info;
int tslam_analyze_structure(const tslam_ba_problem* p, int rank, int world, tslam_structure_info* out);

/* Test hook: runs the structure analysis on the device (the path tslam_solve takes for an unsharded problem) and on the host
 * (the multi-GPU path) and writes the space-separated names of every index array that differs to `report` ("" = identical). */
int tslam_debug_compare_analysis(tslam_ctx* ctx, const tslam_ba_problem* p, char* report, int report_len);

/* Test hooks of the reduced-system solver (the linear solve inside ceres::Solve, src/optimizer.cc:1222,1602,1840).
 * tslam_debug_chol_schedule (host only): the task queue of the fused solve for an n x n system with the given Tn x Tn lower tile
 * pattern; counts_out = {tasks, deps, srcs, below, sync ints, Tn}; each task is 16 int32 (textslam_b200/csrc/chol_sched.hpp).
 * tslam_dev_chol_solve: S x = b for a dense SPD matrix (n x n row-major, lower triangle read) through the solver alone;
 * mode 0 = wave kernels, 1 = fused persistent kernel; ms_out = mean device time per solve; trace_out (fused, optional) = 16 uint64
 * per task (pop, inputs ready, done in ns, SM id, phase stamps); info_out = {tasks, waves, Tn, fail flag}. */
int tslam_debug_chol_schedule(int n, const uint8_t* tile_nz, int32_t counts_out[6], int32_t* tasks, int cap_tasks, int32_t* deps, int cap_deps,
                              int32_t* srcs, int cap_srcs, int32_t* below, int cap_below);
int tslam_dev_chol_solve(tslam_ctx* ctx, int n, const uint8_t* tile_nz, const double* S, const double* b, double* x_out, int mode, int reps,
                         float* ms_out, uint64_t* trace_out, int trace_cap, int32_t info_out[4]);

/* ---- residual + Jacobian evaluation (the metric kernel) --------------------------------------- */
/* Replaces ceres::AutoDiffCostFunction<...>::Evaluate + QuaternionParameterization projection for
 * every residual block of the given kind.  r: n_pobs x 2.  J: n_pobs x 2 x ncols (ncols 13/6/1),
 * observation-major, row-major inside a block; may be NULL. */
int tslam_eval_points(tslam_ctx* ctx, int kind, const tslam_ba_problem* p, double* r, double* J);
/* Replaces ceres::NumericDiffCostFunction<..., CENTRAL, 8, ...>::Evaluate (+ projection).
 * r: n_tobs x 8. J: n_tobs x 8 x ncols (15/6/3). */
int tslam_eval_text(tslam_ctx* ctx, int kind, int jac_mode, const tslam_ba_problem* p, double* r, double* J);

/* ---- Levenberg-Marquardt solve --------------------------------------------------------------- */
/* Replaces ceres::Solve (src/optimizer.cc:1222,1602,1840,1982,2209) followed by
 * Problem::Evaluate (:1228-1233,1609-1614): parameters are updated in place; final_residuals
 * (2*n_pobs + 8*n_tobs, loss-corrected, insertion order; may be NULL) feed the caller's chi^2 gates.
 * trace (may be NULL): (max_iters+1) x TSLAM_TRACE_COLS. */
int tslam_solve(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt,
                tslam_solve_summary* summary, double* final_residuals, double* trace);

/* ---- device-resident handles for benchmarking (inputs already in HBM) ------------------------- */
typedef struct tslam_dev_problem tslam_dev_problem;
int tslam_dev_upload(tslam_ctx* ctx, const tslam_ba_problem* p, tslam_dev_problem** out);
void tslam_dev_free(tslam_ctx* ctx, tslam_dev_problem* d);
/* Launch the point residual+Jacobian kernel `reps` times on the context stream into device-resident
 * outputs; returns the mean kernel time in ms measured with CUDA events on that stream. If
 * flush_l2 != 0 a >L2-sized buffer is rewritten between launches (outside the timed events). */
int tslam_dev_eval_points(tslam_ctx* ctx, tslam_dev_problem* d, int kind, int reps, int flush_l2, float* ms_mean);
int tslam_dev_eval_text(tslam_ctx* ctx, tslam_dev_problem* d, int kind, int jac_mode, int reps, int flush_l2, float* ms_mean);
/* Run `iters` LM iterations on the device-resident problem (parameters reset to the uploaded
 * values first); per-phase mean times in ms (CUDA events) are written to phase_ms[8]:
 * 0 eval+J, 1 landmark/Schur prep, 2 reduced-system build, 3 all-reduce, 4 Cholesky, 5 back-subst,
 * 6 candidate cost, 7 whole iteration. */
int tslam_dev_lm_iterations(tslam_ctx* ctx, tslam_dev_problem* d, const tslam_solve_options* opt, int iters,
                            float* phase_ms, tslam_solve_summary* summary);
/* Copy back device outputs of the last tslam_dev_eval_* (for parity checks of the timed path). */
int tslam_dev_download_eval(tslam_ctx* ctx, tslam_dev_problem* d, int which /*0 pts,1 text*/, double* r, double* J, int ncols);
int tslam_dev_download_params(tslam_ctx* ctx, tslam_dev_problem* d, double* cams, double* rho, double* theta);

/* ---- ORB extractor ------------------------------------------------------------------------ */
/* Replaces ORBextractor::ORBextractor (src/ORBextractor.cc:410-471) and operator()
 * (src/ORBextractor.cc:1054-1116). Keypoints use the 28-byte cv::KeyPoint layout
 * (pt.x pt.y size angle response : f32, octave class_id : i32). */
typedef struct tslam_orb tslam_orb;
typedef struct tslam_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} tslam_keypoint;
int tslam_orb_create(tslam_ctx* ctx, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                     int blur_variant /*0: OpenCV>=3.4/4.x taps, 1: OpenCV 3.3.1 taps*/, tslam_orb** out);
void tslam_orb_destroy(tslam_orb* h);
/* imgs: n_imgs pointers to CV_8UC1 images of w x h, row stride `stride` bytes.
 * kp_out: n_imgs x max_kp keypoints, desc_out: n_imgs x max_kp x 32 bytes, counts_out: n_imgs. */
int tslam_orb_extract(tslam_orb* h, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride,
                      int max_kp, tslam_keypoint* kp_out, uint8_t* desc_out, int32_t* counts_out);
/* Public pyramid (mvImagePyramid, src/ORBextractor.h:85): copy level `level` of image `img` of the
 * last extract call (without border) into out (level_w x level_h, tight). */
int tslam_orb_level_size(tslam_orb* h, int level, int* w, int* hgt);
int tslam_orb_get_level(tslam_orb* h, int img, int level, uint8_t* out);
/* Parity-test hook: read back an intermediate stage of the last extract call (0 FAST measure plane, 1 candidates
 * of a level in vToDistributeKeys order, 2 quad-tree winners of a level; int32 count followed by (x,y,response) triplets). */
int tslam_orb_debug_get(tslam_orb* h, int what, int img, int level, void* out, int out_bytes);
/* Benchmark hook: images already resident in HBM; runs the full extractor `reps` times and returns
 * the mean ms per batch (CUDA events) and the total keypoints of the last batch. */
int tslam_orb_dev_bench(tslam_orb* h, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride,
                        int reps, float* ms_mean, int64_t* n_kp);

/* ---- direct-method frame pyramid ------------------------------------------------ */
/* Replaces frame::GetPyrMat (src/frame.cc:178-202): cv::pyrDown chain, cv::Sobel (CV_8U) in x / y, addWeighted(0.5, 0.5).
 * what: 0 = vFrameImg[level], 1 = vFrameGrad, 2 = vFrameGradX, 3 = vFrameGradY (tight u8 planes). */
typedef struct tslam_frame_pyr tslam_frame_pyr;
int tslam_frame_pyr_create(tslam_ctx* ctx, int nlevels, tslam_frame_pyr** out);
void tslam_frame_pyr_destroy(tslam_frame_pyr* p);
int tslam_frame_pyr_build(tslam_frame_pyr* p, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride);
int tslam_frame_pyr_level_size(tslam_frame_pyr* p, int level, int* w, int* hgt);
int tslam_frame_pyr_get(tslam_frame_pyr* p, int img, int level, int what, uint8_t* out);

/* ---- mu / sigma of projected text quads ------------------------------------------- */
/* Replaces tool::CalTextinfo + CalStatistics (src/tool.cc:1178-1262) for a batch of quads: quads = n_quads x 8 doubles
 * (x0 y0 .. x3 y3, image pixels, the projected text box), quad_img = image index of each quad, imgs = n_imgs x h x w u8.
 * ok_out[i] = the function's bool result (0: empty mask or sigma == 0). */
int tslam_text_info(tslam_ctx* ctx, const uint8_t* imgs, int n_imgs, int w, int h, const double* quads, const int32_t* quad_img, int n_quads,
                    double* mu_out, double* sigma_out, int32_t* ok_out);

/* ---- plane covariance ------------------------------------------------------------- */
/* Replaces the ceres::Covariance block of PyrThetaOptim (src/optimizer.cc:2219-2238): cov_out[n_planes x 9] =
 * (J_theta' J_theta)^-1 per plane from its text blocks (loss-corrected Jacobian); singular blocks give zeros and are
 * counted in *n_singular (may be NULL). */
int tslam_theta_covariance(tslam_ctx* ctx, const tslam_ba_problem* p, int jac_mode, double* cov_out, int32_t* n_singular);

/* ---- post-solve chi^2 gates ------------------------------------------------------ */
/* Replaces the outlier loops that follow Problem::Evaluate in PyrPoseOptim (src/optimizer.cc:1236-1302) and PyrBA
 * (:1616-1684) on the loss-corrected final residuals (2*n_pobs point values followed by 8*n_tobs text values, insertion order):
 *   point observation i bad  <=>  (r[2i]/w_x)^2 > chi2  ||  (r[2i+1]/w_y)^2 > chi2, chi2 = chi2_mono (+ relax_amount when
 *                                 n_tobs < relax_below_text_blocks: ":1240-1241, :1620-1622");
 *   text block j bad         <=>  any of its 8 |r/w_text| > chi2_text (:1264-1281);
 *   text object o bad        <=>  (double)bad_blocks(o) / (double)obj_size[o] > text_ratio (:1286-1294).
 * t_obj[j] = object of block j (vIdx2vTextsGood / vIdx2IdxTexts), obj_size[o] = vSizeEachObj[o]; the reference adds the
 * blocks of an object contiguously, and the number of blocks naming an object must equal obj_size[o] (TSLAM_ERR_ARG
 * otherwise; objects of size 0 are never flagged). Outputs are 0/1 bytes (1 = the reference sets the Good flag to false);
 * the caller keeps applying them to its vObvGood* vectors through its own index maps. counts_out (may be NULL) =
 * {nBadS, bad text blocks, nBadT}. */
typedef struct tslam_gate_options {
  int32_t gate_points;              /* SCENEOutlier */
  int32_t gate_text;                /* TEXTOutlier  */
  double w_point[2];                /* weight_S_x, weight_S_y the residuals carry   */
  double chi2_mono;                 /* 12.25                                        */
  int32_t relax_below_text_blocks;  /* 50 (0 disables the relaxation)               */
  double relax_amount;              /* 4                                            */
  double w_text;                    /* weight_T                                     */
  double chi2_text;                 /* 0.5 / 0.95 (pose), 0.5..0.8 (BA)             */
  double text_ratio;                /* 0.99                                         */
} tslam_gate_options;
/* residuals in caller-owned host memory (any solve, any source) */
int tslam_gate_residuals(tslam_ctx* ctx, const double* final_residuals, int n_pobs, int n_tobs, const int32_t* t_obj,
                         const int32_t* obj_size, int n_obj, const tslam_gate_options* gate,
                         uint8_t* pt_bad, uint8_t* tf_bad, uint8_t* obj_bad, int32_t counts_out[3]);
/* tslam_solve followed by the gates on the residuals while they are still in HBM (no second upload): one call per
 * pyramid level of PoseOptim / LocalBundleAdjustment. final_residuals and trace may be NULL. */
int tslam_solve_gated(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, const tslam_gate_options* gate,
                      const int32_t* t_obj, const int32_t* obj_size, int n_obj,
                      tslam_solve_summary* summary, double* final_residuals, double* trace,
                      uint8_t* pt_bad, uint8_t* tf_bad, uint8_t* obj_bad, int32_t counts_out[3]);

/* ---- projection-guided descriptor matching core ------------------------------------ */
/* Inner loop of tracking::SearchFrom3D* (src/tracking.cc:1161-1175,1241-1256,1310-1325) with
 * tracking::DescriptorDistance (:2762-2778): per query the first candidate (list order) of minimum Hamming distance.
 * Descriptors are 32 bytes; candidates in CSR form (cand_ptr[n_query+1], cand_idx into train). Empty list -> idx -1,
 * dist INT_MAX. second_dist (may be NULL) = distance of the runner-up candidate. */
int tslam_match_hamming(tslam_ctx* ctx, const uint8_t* query_desc, int n_query, const uint8_t* train_desc, int n_train,
                        const int32_t* cand_ptr, const int32_t* cand_idx, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* tracking::SearchFrom3D / SearchFrom3DAdd / SearchFrom3DLocalTrack up to their uniqueness bookkeeping (src/tracking.cc:1124-1176,
 * 1206-1256, 1282-1327): per map point, projection into the frame (pose Tcw = [qw qx qy qz tx ty tz], host pose of the point from
 * `poses`, inverse depth, ray direction (x, y, 1)), the bounds test, frame::GetFeaturesInArea (src/frame.cc:415-468) on the frame's
 * keypoint grid and the first minimum-Hamming-distance candidate. pt_query[i] = row of query_desc (the descriptor of the point's
 * observation in the last key frame, F1->mDescr.row(IdxObserv)); < 0 skips the point (FLAG_BAD / not observed, :1126-1132).
 * Outputs per point: best_idx (-1: out of bounds or no candidate), best_dist (INT_MAX then), optionally the projection (u, v). */
typedef struct tslam_frame_grid {
  int32_t cols, rows;               /* FRAME_GRID_COLS = 64, FRAME_GRID_ROWS = 48 (src/frame.h:26-27)             */
  float min_x, min_y, max_x, max_y; /* mnMinX, mnMinY, mnMaxX, mnMaxY                                             */
  float inv_w, inv_h;               /* mfGridElementWidthInv, mfGridElementHeightInv (src/frame.cc:124-125)       */
  const int32_t* cell_ptr;          /* cols*rows + 1; cell id = ix * rows + iy, i.e. mGrid[ix][iy]                */
  const int32_t* cell_idx;          /* keypoint indices, insertion order inside a cell (src/frame.cc:386-389)     */
} tslam_frame_grid;
int tslam_search_from_3d(tslam_ctx* ctx, const double* Tcw, const double* K, int n_pts, const double* pt_ray, const double* pt_rho,
                         const double* poses, int n_poses, const int32_t* pt_host, const int32_t* pt_query,
                         const uint8_t* query_desc, int n_query, const float* kp_xy, const int32_t* kp_octave,
                         const uint8_t* train_desc, int n_kp, const tslam_frame_grid* grid, float radius, int min_level,
                         int max_level, int32_t* best_idx, int32_t* best_dist, double* uv_out);
/* tracking::SearchFrom3DLocalTrack (src/tracking.cc:1282-1345): the projections are given (mapPts::LocalTrackProj), no bounds test;
 * kp_skip[k] != 0 excludes key point k (already matched to a map point with more than two observations, :1311-1313; may be NULL);
 * second_dist receives the runner-up distance for the caller's ratio test bestDist <= 0.9 bestDist2 (:1331-1334; may be NULL). */
int tslam_search_in_area(tslam_ctx* ctx, int n_pts, const double* uv, const int32_t* pt_query, const uint8_t* query_desc, int n_query,
                         const float* kp_xy, const int32_t* kp_octave, const uint8_t* kp_skip, const uint8_t* train_desc, int n_kp,
                         const tslam_frame_grid* grid, float radius, int min_level, int max_level, int32_t* best_idx,
                         int32_t* best_dist, int32_t* second_dist);

#ifdef __cplusplus
}
#endif
#endif /* TSLAM_B200_H_ */
