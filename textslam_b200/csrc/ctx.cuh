// Context, error handling and the device-resident problem (HBM layout) of libtslam_b200.
#pragma once
#include "analysis.hpp"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>
#include "../../include/tslam_b200.h"

namespace tsl {

extern thread_local std::string g_last_error;
extern long long g_launches;
int set_error(int code, const char* fmt, ...);

#define TSL_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return tsl::set_error(TSLAM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define TSL_CHECK_LAUNCH() TSL_CUDA(cudaGetLastError())
// every kernel launch of this library goes through LAUNCH so that bench.py can report gpu_launches
#define LAUNCH(...) do { ++tsl::g_launches; __VA_ARGS__; } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// An LM iteration is ~50 short dependent kernels (29 of them the wave-scheduled Cholesky); between two of them the GPU
// otherwise drains, then fetches and schedules the next grid. Every solver kernel starts with PDL_PROLOGUE():
// `griddepcontrol.launch_dependents` lets the next kernel of the stream be scheduled as soon as all CTAs of this one have
// started, `griddepcontrol.wait` then blocks until the previous kernel has completed and its writes are visible — so
// the data dependencies are exactly those of plain stream order, only the launch latency overlaps the predecessor's tail.
// Both instructions are no-ops for a kernel launched without the attribute (TSLAM_PDL=0, or launch through <<<>>>).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#define PDL_PROLOGUE() tsl::pdl_prologue()
// Split form for kernels that can read launch-invariant index structures (uploaded once per problem, never written by a
// kernel) before they have to wait for their predecessor: PDL_TRIGGER(); <index loads>; PDL_WAIT(); <everything else>.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define PDL_TRIGGER() tsl::pdl_trigger()
#define PDL_WAIT() tsl::pdl_wait()
bool pdl_enabled();   // capi.cu: TSLAM_PDL != "0"
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through TSL_CHECK_LAUNCH (cudaGetLastError)
}
#endif

// Stream the calling entry point wants its device buffers allocated (and released) on. tslam_solve sets it to the context
// stream for the duration of the call: every use of those buffers is on that stream, so stream order alone makes the
// ~130 allocations of a solve safe and none of them needs a host synchronisation (0.6 ms per call on the global-BA shape).
// nullptr (default): allocate on the per-thread stream and synchronise, the buffer may then be used on any stream.
extern thread_local cudaStream_t g_alloc_stream;
struct AllocStreamScope {
  cudaStream_t prev;
  explicit AllocStreamScope(cudaStream_t s) : prev(g_alloc_stream) { g_alloc_stream = s; }
  ~AllocStreamScope() { g_alloc_stream = prev; }
};

template <typename T>
struct DevBuf {  // simple RAII device buffer (grow-only)
  T* p = nullptr;
  size_t cap = 0;
  // back to the (never trimmed) default pool: the next solve reuses it
  ~DevBuf() { if (p) cudaFreeAsync(p, g_alloc_stream ? g_alloc_stream : cudaStreamPerThread); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t reserve(size_t n) {
    if (p && n <= cap) return cudaSuccess;
    cudaStream_t as = g_alloc_stream;
    if (p) cudaFreeAsync(p, as ? as : cudaStreamPerThread);
    p = nullptr; cap = 0;
    // stream-ordered allocation from the device's default memory pool (release threshold raised in tslam_ctx_create):
    // repeated tslam_solve calls recycle their buffers instead of paying cudaMalloc/cudaFree every time
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(T), as ? as : cudaStreamPerThread);
    if (e == cudaSuccess && !as) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  cudaError_t upload(const T* h, size_t n, cudaStream_t s) {
    cudaError_t e = reserve(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

struct NcclApi;  // dlopen'ed NCCL entry points (comm.cu)

}  // namespace tsl

struct tslam_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // L2 flush buffer (> L2 size) used between timed launches
  tsl::DevBuf<uint8_t> flush;
  size_t l2_bytes = 0;
  // multi-GPU
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  // pinned staging for scalars
  double* h_scalars = nullptr;  // cudaHostAlloc (mapped), 64 doubles: [0..9] iteration scalars, [63] sequence number written last by the device
  double* h_scalars_dev = nullptr;   // device-side address of h_scalars
  double h_seq = 0.0;                // last sequence number handed to a publishing kernel
  // page-locked arena behind the host-side structure analysis (analysis.hpp); recycled by every solve on this context
  tsl::Arena* host_arena = nullptr;
  // per-context (= per-device) opt-ins to more than 48 KB of dynamic shared memory
  bool attr_chol_fused = false, attr_chol_waves = false, attr_orb = false;
  size_t attr_textinfo_smem = 0;
  void* small_ws = nullptr;   // tsl::SmallWorkspace (ba_small.cu): staging block + device workspace of the small-problem solve
};

// Device-resident problem: everything the kernels read, SoA, FP64 / int32 / u8.
struct tslam_dev_problem {
  int n_cams = 0, n_points = 0, n_planes = 0, n_pobs = 0, n_tobs = 0, n_imgs = 0, img_w = 0, img_h = 0;
  double K_point[4], w_point[2], huber_point, K_text[4], w_text, huber_text;
  tsl::DevBuf<double> cams, rho, theta;            // current parameters
  tsl::DevBuf<double> cams0, rho0, theta0;         // uploaded values (reset point for benchmarks)
  tsl::DevBuf<uint8_t> cam_fixed, rho_fixed, theta_fixed;
  tsl::DevBuf<double> p_uv, p_ray;                 // n_pobs x 2 each (16 B / obs each)
  tsl::DevBuf<int32_t> p_cam, p_host, p_lm;
  tsl::DevBuf<double> t_rays, t_iref, t_musigma;
  tsl::DevBuf<int32_t> t_cam, t_host, t_plane, t_img;
  tsl::DevBuf<uint8_t> imgs;
  tsl::DevBuf<int32_t> t_run_ptr; int n_truns = 0;   // runs of consecutive text blocks with the same (camera, host, plane, image), <= 32 blocks each (ba_eval_tma.cu)
  // evaluation outputs (observation-major): r, J
  tsl::DevBuf<double> pr, pJ, tr, tJ;
  int pJ_cols = 0, tJ_cols = 0;
  // host copies (GLOBAL problem) kept for the solver's structure analysis
  int g_pobs = 0, g_tobs = 0;                       // global observation counts (== n_pobs/n_tobs unless sharded)
  std::vector<int32_t> h_p_cam, h_p_host, h_p_lm, h_t_cam, h_t_host, h_t_plane;
  std::vector<uint8_t> h_cam_fixed, h_rho_fixed, h_theta_fixed;
  std::vector<int32_t> gsel_p, gsel_t;              // global index of each local observation (sharded upload)
  bool sharded = false;
  bool have_host_index = true;   // h_p_* / h_t_* filled (needed by the host-side structure analysis only)
  void* solver = nullptr;  // tsl::Solver*, owned (ba_solve.cu)
};

namespace tsl {
// shard = true: keep only the observations this rank owns (multi-GPU global BA, ctx->world > 1)
// persistent = false (one-shot tslam_solve, the caller is blocked until the call ends): no reset copies of the parameters, host
// copies of the observation index arrays only when the structure analysis will run on the host, and the uploads may still be in
// flight on return (everything that follows is ordered behind them on the context stream)
int upload_problem(tslam_ctx* ctx, const tslam_ba_problem* p, tslam_dev_problem* d, bool shard = false, bool persistent = true, bool validated = false);
int validate_problem(const tslam_ba_problem* p);   // argument checks of every entry point that takes a host problem
bool device_analysis_supported(const tslam_ctx* ctx, const tslam_dev_problem* d);
// observation ownership rule shared by upload and the solver's structure analysis
inline int obs_owner(bool lm_free, int lm_index, int obs_index, int world) { return (lm_free ? lm_index : obs_index) % world; }
int flush_l2(tslam_ctx* ctx);
void free_solver(tslam_dev_problem* d);
// kernels (ba_eval.cu)
int launch_eval_points(tslam_ctx* ctx, tslam_dev_problem* d, int kind, bool want_J);
int launch_eval_text(tslam_ctx* ctx, tslam_dev_problem* d, int kind, int jac_mode, bool want_J);
int launch_eval_text_tma(tslam_ctx* ctx, tslam_dev_problem* d, int kind);   // ba_eval_tma.cu (jac_mode TSLAM_JAC_ANALYTIC_TMA)
}  // namespace tsl
