// Structure analysis of an (unsharded) BA problem ON THE DEVICE: the same index structures analysis.cpp builds on the
// host — free-parameter layout, nested-dissection camera order, landmark -> camera-slot incidence, non-zero 6x6 blocks of
// the reduced camera system with their gather lists — produced with radix sorts and scans (CUB) over the index arrays
// that are already in HBM for the residual kernels. This is the per-call work ceres::Problem / the Schur ordering do
// inside ceres::Solve (src/optimizer.cc:1222,1602,1840,1982,2209); on the host it was 40 % of the end-to-end time of a
// global BA solve (4-7 ms on 8 threads, profiles/r1_notes.md), here it is a few hundred microseconds of launches.
// Every array is bit-identical to the host analysis (tests/test_gpu_analysis.py); the tile-level symbolic factorisation
// (47 x 47 tiles) stays on the host.
//
// Packing limits (device_analysis_supported): 2^24 free landmarks, 2^14 free cameras with nc^2 <= 2^24, 2^25 observations
// per type; larger problems take the host path.
//
// Landmark-sharded problems (multi-GPU global BA): a rank holds only the observations it owns, but the camera layout, the
// camera order and the block numbering of the reduced system must be the same on every rank (the packed block buffer is
// all-reduced element by element). Three small integer tables carry everything global — cameras in use (K), the
// co-visibility distance histogram (K) and the K x K block flags — and are summed across the ranks with NCCL right after
// the local passes that fill them; everything derived from them is then identical everywhere, the landmark side
// (numbering, slots, gather lists) stays local by construction (an observation lives with its landmark).
#include <cub/cub.cuh>
#include <chrono>
#include "ctx.cuh"
#include "solver.cuh"
#include "nd_layout.h"

namespace tsl {
namespace {

typedef unsigned long long u64;
constexpr int CODE_BITS = 26, CAM_BITS = 14;
constexpr u64 KEY_INVALID = ~0ull;

struct Counts {   // device-side scalars, copied to the host once
  int nc, nl, npl;
  int n_ent_p, n_ent_t, nsp, nst, nblk;
  int npairs_p, npairs_t;
  int overflow;
  int npad;   // columns of the tile-aligned camera layout (nd_layout.h)
};

static int bits_for(unsigned long long max_value) { int b = 1; while (b < 64 && (max_value >> b)) ++b; return b; }

// ---------------------------------------------------------------- layout -------------------------------------------
__global__ void flags_kernel(int n, const int* __restrict__ cam, const int* __restrict__ host, const int* __restrict__ lm,
                             const uint8_t* __restrict__ cf, const uint8_t* __restrict__ lf, uint8_t* __restrict__ act, int* cu, int* lu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cam[i], h = host[i], l = lm[i];
  const bool a = !cf[c] || !cf[h] || !lf[l];
  act[i] = a;
  if (a) { cu[c] = 1; cu[h] = 1; lu[l] = 1; }   // every writer stores the same value
}

__global__ void __launch_bounds__(1024) cam_layout_kernel(int K, const int* __restrict__ cu, const uint8_t* __restrict__ cf, int* __restrict__ camslot, Counts* cnt) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += 1024) {
    const int k = k0 + threadIdx.x;
    const int f = (k < K && cu[k] && !cf[k]) ? 1 : 0;
    int excl, total;
    Scan(tmp).ExclusiveSum(f, excl, total);
    if (k < K) camslot[k] = f ? base + excl : -1;
    __syncthreads();
    if (threadIdx.x == 0) base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt->nc = base;
}

__global__ void dist_hist_kernel(int n, const int* __restrict__ cam, const int* __restrict__ host, const uint8_t* __restrict__ act,
                                 const int* __restrict__ camslot, int* hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !act[i]) return;
  const int a = camslot[cam[i]], b = camslot[host[i]];
  if (a >= 0 && b >= 0) atomicAdd(&hist[abs(a - b)], 1);
}

// Nested-dissection order of the free cameras — the logic of analysis.cpp (analyze_structure) on one warp: the inputs are
// a histogram and the outputs a permutation of <= 2^14 cameras; the scans over it are lane-parallel, the (tiny) recursive
// bisection runs on lane 0.
__global__ void __launch_bounds__(32) nd_order_kernel(int K, int* camslot, const int* __restrict__ hist, Counts* cnt, int* new_of_old, int* doff) {
  const int lane = threadIdx.x;
  const int nc = cnt->nc;
  for (int s0 = lane; s0 < nc; s0 += 32) doff[s0] = 6 * s0;   // default: natural order, no padding
  if (nc < 128) { if (lane == 0) cnt->npad = 6 * nc; return; }
  unsigned long long nd = 0;
  for (int d = lane; d < nc; d += 32) nd += (unsigned long long)hist[d];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
  int bw = 0;
  if (nd) {   // 98th percentile of the |slot(cam) - slot(host)| distances: first d with cumulative count > q
    const unsigned long long q = (unsigned long long)(0.98 * (double)(nd - 1));
    unsigned long long acc = 0; int dq = 0;
    for (int d0 = 0; d0 < nc; d0 += 32) {
      const int d = d0 + lane;
      unsigned long long inc = d < nc ? (unsigned long long)hist[d] : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      const unsigned hit = __ballot_sync(0xffffffffu, d < nc && acc + inc > q);
      if (hit) { dq = d0 + __ffs(hit) - 1; break; }
      acc += __shfl_sync(0xffffffffu, inc, 31);
    }
    bw = 2 * dq;
  }
  // node table of the plan (nd_layout.h) on lane 0, then the lanes fill the permutation and the column offsets
  __shared__ int s_nat[128], s_size[128], s_slot[128], s_col[128];
  __shared__ int s_nodes, s_npad;
  if (lane == 0) {
    const NdPlan P = nd_plan(nc, bw);
    int nn = 0, slot = 0, d = 0;
    if (P.levels > 0) {
      nn = nd_node_count(P);
      for (int k = 0; k < nn; ++k) {
        int s0, sz;
        nd_node(P, k, &s0, &sz);
        d = (d + 63) / 64 * 64;
        s_nat[k] = s0; s_size[k] = sz; s_slot[k] = slot; s_col[k] = d;
        slot += sz; d += 6 * sz;
      }
    }
    s_nodes = nn; s_npad = nn ? d : 6 * nc;
  }
  __syncwarp();
  const int nn = s_nodes;
  if (lane == 0) cnt->npad = s_npad;
  if (nn == 0) return;   // natural order: doff = 6 s (filled by the caller's default), camslot unchanged
  for (int k = 0; k < nn; ++k)
    for (int c = lane; c < s_size[k]; c += 32) { new_of_old[s_nat[k] + c] = s_slot[k] + c; doff[s_slot[k] + c] = s_col[k] + 6 * c; }
  __threadfence_block();
  __syncwarp();
  for (int k = lane; k < K; k += 32) if (camslot[k] >= 0) camslot[k] = new_of_old[camslot[k]];
}

__global__ void lm_flag_kernel(int n, const int* __restrict__ lu, const uint8_t* __restrict__ lf, int* __restrict__ f) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) f[l] = (lu[l] && !lf[l]) ? 1 : 0;
}
__global__ void lm_free_kernel(int n, const int* __restrict__ f, const int* __restrict__ pre, int* __restrict__ lmfree, int* __restrict__ v_gl, int* n_free) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  lmfree[l] = f[l] ? pre[l] : -1;
  if (f[l]) v_gl[pre[l]] = l;
  if (l == n - 1) *n_free = pre[l] + f[l];
}

// per observation: camera slots, owned landmark, sort key by landmark, the two (landmark, camera, code) entry keys
__global__ void obs_kernel(int n, const int* __restrict__ cam, const int* __restrict__ host, const int* __restrict__ lm, const uint8_t* __restrict__ act,
                           const int* __restrict__ camslot, const int* __restrict__ lmfree, const int* n_free,
                           int* __restrict__ cs, int* __restrict__ hs, int* __restrict__ ls, uint8_t* __restrict__ fmask,
                           unsigned* __restrict__ key_lm, int* __restrict__ iota, u64* __restrict__ ent_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = camslot[cam[i]], h = camslot[host[i]], lf = lmfree[lm[i]];
  const int v = (act[i] && lf >= 0) ? lf : -1;
  cs[i] = c; hs[i] = h; ls[i] = v;
  if (fmask) fmask[i] = (uint8_t)((c >= 0 ? 1 : 0) | (h >= 0 ? 2 : 0) | (lf >= 0 ? 4 : 0));
  key_lm[i] = v >= 0 ? (unsigned)v : (unsigned)*n_free;
  iota[i] = i;
  const u64 hi = (u64)(unsigned)v << (CODE_BITS + CAM_BITS);
  ent_key[2 * (size_t)i] = (v >= 0 && c >= 0) ? (hi | ((u64)c << CODE_BITS) | (u64)((i << 1) | 0)) : KEY_INVALID;
  ent_key[2 * (size_t)i + 1] = (v >= 0 && h >= 0) ? (hi | ((u64)h << CODE_BITS) | (u64)((i << 1) | 1)) : KEY_INVALID;
}

// out[q] = first position p in sorted[0, *n) with sorted[p] >= q, for q = 0 .. nq-1 (CSR row pointers from sorted keys)
template <typename KeyT>
__global__ void lower_bound_kernel(const KeyT* __restrict__ sorted, const int* n_ptr, int n_fixed, int nq, int* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int n = n_ptr ? *n_ptr : n_fixed;
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] < (KeyT)q) lo = mid + 1; else hi = mid; }
  out[q] = lo;
}

__global__ void count_valid_kernel(const u64* __restrict__ sorted, int n, int* n_valid) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] != KEY_INVALID) lo = mid + 1; else hi = mid; }
  *n_valid = lo;
}

__global__ void slot_head_kernel(const u64* __restrict__ key, int n, const int* n_valid, int* __restrict__ head) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  head[p] = (p < *n_valid && (p == 0 || (key[p] >> CODE_BITS) != (key[p - 1] >> CODE_BITS))) ? 1 : 0;
}

__global__ void slot_emit_kernel(const u64* __restrict__ key, int n, const int* n_valid, const int* __restrict__ head, const int* __restrict__ excl,
                                 int* __restrict__ slot_cam, int* __restrict__ slot_lm, int* __restrict__ ent_ptr, int* __restrict__ ent, int* n_slots) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = *n_valid;
  if (p == 0 && nv == 0) { *n_slots = 0; ent_ptr[0] = 0; }
  if (p >= n || p >= nv) return;
  const u64 k = key[p];
  ent[p] = (int)(k & ((1ull << CODE_BITS) - 1));
  if (head[p]) {
    const int s = excl[p];
    slot_cam[s] = (int)((k >> CODE_BITS) & ((1ull << CAM_BITS) - 1));
    slot_lm[s] = (int)(k >> (CODE_BITS + CAM_BITS));
    ent_ptr[s] = p;
  }
  if (p == nv - 1) { const int ns = excl[p] + head[p]; *n_slots = ns; ent_ptr[ns] = nv; }
}

__global__ void pair_count_kernel(int nq, const int* n_lm, const int* __restrict__ slot_ptr, int* __restrict__ pc) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nq) return;
  int c = 0;
  if (v < *n_lm) { const int m = slot_ptr[v + 1] - slot_ptr[v]; c = m * (m + 1) / 2; }
  pc[v] = c;
}
__global__ void pair_total_kernel(const int* __restrict__ pair_off, const int* n_lm, int* total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *total = pair_off[*n_lm];
}

// ---------------------------------------------------------------- blocks -------------------------------------------
__global__ void mark_direct_kernel(int n, const int* __restrict__ cs, const int* __restrict__ hs, const uint8_t* __restrict__ act, const Counts* cnt, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !act[i]) return;
  const int nc = cnt->nc, c = cs[i], h = hs[i];
  if (c >= 0) flag[(size_t)c * nc + c] = 1;
  if (h >= 0) flag[(size_t)h * nc + h] = 1;
  if (c >= 0 && h >= 0 && c != h) flag[(size_t)min(c, h) * nc + max(c, h)] = 1;
}
__global__ void mark_schur_kernel(int nq, const int* n_lm, const int* __restrict__ slot_ptr, const int* __restrict__ slot_cam, const Counts* cnt, int* flag) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nq || v >= *n_lm) return;
  const int nc = cnt->nc;
  for (int x = slot_ptr[v]; x < slot_ptr[v + 1]; ++x)
    for (int y = x; y < slot_ptr[v + 1]; ++y) flag[(size_t)slot_cam[x] * nc + slot_cam[y]] = 1;
}
__global__ void block_count_kernel(int ncap2, const int* __restrict__ flag, const int* __restrict__ pre, Counts* cnt, const int* __restrict__ doff, uint8_t* __restrict__ tile_nz, int tn_cap) {
  // nblk, and the 64x64 tile pattern of the lower triangle (block (a,b), a <= b: rows doff[b].., cols doff[a]..)
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = cnt->nc;
  const size_t nc2 = (size_t)nc * nc;
  if (t == 0) cnt->nblk = nc2 ? pre[nc2 - 1] + flag[nc2 - 1] : 0;
  if (t >= nc2 || !flag[t]) return;
  const int a = (int)(t / nc), b = (int)(t % nc);
  const int Tn = (cnt->npad + 63) / 64;
  const int r0 = doff[b] / 64, r1 = (doff[b] + 5) / 64, c0 = doff[a] / 64, c1 = (doff[a] + 5) / 64;
  for (int r = r0; r <= r1; ++r) for (int c = c0; c <= c1; ++c) if (c <= r) tile_nz[(size_t)r * Tn + c] = 1;
  (void)ncap2; (void)tn_cap;
}
__global__ void block_emit_kernel(const int* __restrict__ flag, const int* __restrict__ pre, const Counts* cnt, int* __restrict__ blk_a, int* __restrict__ blk_b,
                                  int* __restrict__ diag_blk, int* __restrict__ offdiag_blk) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = cnt->nc;
  if (t >= (size_t)nc * nc || !flag[t]) return;
  const int a = (int)(t / nc), b = (int)(t % nc), id = pre[t];
  blk_a[id] = a; blk_b[id] = b;
  if (a == b) diag_blk[a] = id;
  else offdiag_blk[id - (a + 1)] = id;   // every free camera has its diagonal block: a + 1 of them precede (a, b > a)
}

// ---------------------------------------------------------------- gather lists --------------------------------------
__global__ void direct_keys_kernel(int n, const int* __restrict__ cs, const int* __restrict__ hs, const uint8_t* __restrict__ act, const int* __restrict__ pre,
                                   const int* __restrict__ diag_blk, const Counts* cnt, unsigned* __restrict__ key, int* __restrict__ val) {
  // four candidates per observation in generation order: (c,c) code 0, (h,h) code 1, off-diagonal code 2|3, and the second
  // entry of the degenerate cam == host case (never produced by the reference, src/optimizer.cc:1397)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nc = cnt->nc; const unsigned inv = (unsigned)cnt->nblk;
  unsigned k[4] = {inv, inv, inv, inv};
  int code2 = 2;
  if (act[i]) {
    const int c = cs[i], h = hs[i];
    if (c >= 0) k[0] = (unsigned)diag_blk[c];
    if (h >= 0) k[1] = (unsigned)diag_blk[h];
    if (c >= 0 && h >= 0) {
      if (c < h) k[2] = (unsigned)pre[(size_t)c * nc + h];
      else if (h < c) { k[2] = (unsigned)pre[(size_t)h * nc + c]; code2 = 3; }
      else { k[2] = (unsigned)diag_blk[c]; k[3] = (unsigned)diag_blk[c]; }
    }
  }
  key[4 * (size_t)i] = k[0]; val[4 * (size_t)i] = (i << 2) | 0;
  key[4 * (size_t)i + 1] = k[1]; val[4 * (size_t)i + 1] = (i << 2) | 1;
  key[4 * (size_t)i + 2] = k[2]; val[4 * (size_t)i + 2] = (i << 2) | code2;
  key[4 * (size_t)i + 3] = k[3]; val[4 * (size_t)i + 3] = (i << 2) | 3;
}
__global__ void schur_keys_kernel(int nq, const int* n_lm, const int* __restrict__ slot_ptr, const int* __restrict__ slot_cam, const int* __restrict__ pair_off,
                                  const int* __restrict__ pre, const Counts* cnt, unsigned* __restrict__ key, u64* __restrict__ val) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nq || v >= *n_lm) return;
  const int nc = cnt->nc;
  size_t e = (size_t)pair_off[v];
  for (int x = slot_ptr[v]; x < slot_ptr[v + 1]; ++x)
    for (int y = x; y < slot_ptr[v + 1]; ++y) {
      key[e] = (unsigned)pre[(size_t)slot_cam[x] * nc + slot_cam[y]];
      val[e] = ((u64)(unsigned)y << 32) | (u64)(unsigned)x;   // == int2{x, y} in memory
      ++e;
    }
}
__global__ void zero_int_kernel(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = 0; }
__global__ void clamp01_kernel(int* p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] != 0; }

inline int grid(size_t n, int b) { return n > 0 ? (int)((n + b - 1) / b) : 1; }

}  // namespace

bool device_analysis_supported(const tslam_ctx* ctx, const tslam_dev_problem* d) {
  static const bool force_host = getenv("TSLAM_HOST_ANALYSIS") != nullptr;
  if (force_host) return false;
  if ((size_t)d->n_cams * d->n_cams > ((size_t)1 << 24) || d->n_cams >= (1 << CAM_BITS)) return false;
  if (d->n_points >= (1 << 24) || d->n_planes >= (1 << 24)) return false;
  if (d->n_pobs >= (1 << 25) || d->n_tobs >= (1 << 25)) return false;
  return d->n_cams > 0;
}

// One landmark type (inverse depths or planes): everything between the observation arrays and the slot structure.
struct TypeIn { int n_obs, n_lm; const int *cam, *host, *lm; const uint8_t* lm_fixed; };
struct TypeBufs {   // scratch that must outlive the type pass (used again by the block / list stage)
  DevBuf<int> lu, f, pre, lmfree, iota, head, excl, pc, pair_off;
  DevBuf<unsigned> key_lm, key_lm_s;
  DevBuf<u64> ent_key, ent_key_s;
};

int analyze_structure_device(tslam_ctx* ctx, tslam_dev_problem* d, SolverIndex& X, double* lap_ms) {
  auto T0 = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  const int K = d->n_cams, np = d->n_pobs, nt = d->n_tobs, NP = d->n_points, NL = d->n_planes;
  X.K = K; X.lp = np; X.lt = nt;
  // ---- CUB temporary storage: one buffer sized for the largest call ----
  size_t temp_bytes = 0;
  {
    size_t b = 0;
    auto upd = [&](size_t x) { temp_bytes = std::max(temp_bytes, x); };
    const int nmax = std::max(std::max(np, nt), 1);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, 4 * nmax, 0, 32, st); upd(b);
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const u64*)nullptr, (u64*)nullptr, 2 * nmax, 0, 64, st); upd(b);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, std::max((size_t)K * K, (size_t)std::max(2 * nmax, std::max(NP, NL) + 1)), st); upd(b);
  }
  DevBuf<uint8_t> temp;
  TSL_CUDA(temp.reserve(temp_bytes + 256));
  DevBuf<Counts> cnt;
  TSL_CUDA(cnt.reserve(1));
  TSL_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(Counts), st));
  // ---- flags, camera layout, nested-dissection order ----
  DevBuf<int> cu, hist, new_of_old;
  TSL_CUDA(cu.reserve(K)); TSL_CUDA(hist.reserve(K)); TSL_CUDA(new_of_old.reserve(K)); TSL_CUDA(X.doff.reserve(K));
  TSL_CUDA(cudaMemsetAsync(cu.p, 0, sizeof(int) * K, st)); TSL_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(int) * K, st));
  TypeBufs BP, BT;
  TSL_CUDA(BP.lu.reserve(NP)); TSL_CUDA(BT.lu.reserve(NL));
  TSL_CUDA(cudaMemsetAsync(BP.lu.p, 0, sizeof(int) * (size_t)(NP ? NP : 1), st)); TSL_CUDA(cudaMemsetAsync(BT.lu.p, 0, sizeof(int) * (size_t)(NL ? NL : 1), st));
  TSL_CUDA(X.p_active.reserve(np)); TSL_CUDA(X.t_active.reserve(nt)); TSL_CUDA(X.t_fmask.reserve(nt));
  TSL_CUDA(X.camslot_d.reserve(K));
  if (np) LAUNCH(flags_kernel<<<grid(np, 256), 256, 0, st>>>(np, d->p_cam.p, d->p_host.p, d->p_lm.p, d->cam_fixed.p, d->rho_fixed.p, X.p_active.p, cu.p, BP.lu.p));
  if (nt) LAUNCH(flags_kernel<<<grid(nt, 256), 256, 0, st>>>(nt, d->t_cam.p, d->t_host.p, d->t_plane.p, d->cam_fixed.p, d->theta_fixed.p, X.t_active.p, cu.p, BT.lu.p));
  const bool multi = ctx->world > 1;
  int rc;
  if (multi && (rc = comm_allreduce_sum_i32(ctx, cu.p, K))) return rc;        // cameras used by anybody's active observations
  LAUNCH(cam_layout_kernel<<<1, 1024, 0, st>>>(K, cu.p, d->cam_fixed.p, X.camslot_d.p, cnt.p));
  if (np) LAUNCH(dist_hist_kernel<<<grid(np, 256), 256, 0, st>>>(np, d->p_cam.p, d->p_host.p, X.p_active.p, X.camslot_d.p, hist.p));
  if (nt) LAUNCH(dist_hist_kernel<<<grid(nt, 256), 256, 0, st>>>(nt, d->t_cam.p, d->t_host.p, X.t_active.p, X.camslot_d.p, hist.p));
  if (multi && (rc = comm_allreduce_sum_i32(ctx, hist.p, K))) return rc;      // global co-visibility distances -> same camera order
  LAUNCH(nd_order_kernel<<<1, 32, 0, st>>>(K, X.camslot_d.p, hist.p, cnt.p, new_of_old.p, X.doff.p));
  TSL_CHECK_LAUNCH();

  // ---- per landmark type: free numbering, observation CSR, slots ----
  auto type_pass = [&](const TypeIn& in, TypeBufs& B, const uint8_t* act, int* n_free_dev, int* n_ent_dev, int* n_slots_dev, int* n_pairs_dev,
                       DevBuf<int>& cs, DevBuf<int>& hs, DevBuf<int>& ls, uint8_t* fmask, DevBuf<int>& v_gl, DevBuf<int>& obs_ptr, DevBuf<int>& obs,
                       DevBuf<int>& slot_ptr, DevBuf<int>& slot_cam, DevBuf<int>& slot_lm, DevBuf<int>& ent_ptr, DevBuf<int>& ent) -> int {
    const int n = in.n_obs, nlm = in.n_lm;
    TSL_CUDA(B.f.reserve(nlm + 1)); TSL_CUDA(B.pre.reserve(nlm + 1)); TSL_CUDA(B.lmfree.reserve(nlm)); TSL_CUDA(v_gl.reserve(nlm));
    TSL_CUDA(cs.reserve(n)); TSL_CUDA(hs.reserve(n)); TSL_CUDA(ls.reserve(n));
    TSL_CUDA(obs_ptr.reserve((size_t)nlm + 1)); TSL_CUDA(obs.reserve(n));
    TSL_CUDA(slot_ptr.reserve((size_t)nlm + 1)); TSL_CUDA(slot_cam.reserve(2 * (size_t)n)); TSL_CUDA(slot_lm.reserve(2 * (size_t)n));
    TSL_CUDA(ent_ptr.reserve(2 * (size_t)n + 1)); TSL_CUDA(ent.reserve(2 * (size_t)n));
    TSL_CUDA(B.iota.reserve(n)); TSL_CUDA(B.key_lm.reserve(n)); TSL_CUDA(B.key_lm_s.reserve(n));
    TSL_CUDA(B.ent_key.reserve(2 * (size_t)n)); TSL_CUDA(B.ent_key_s.reserve(2 * (size_t)n));
    TSL_CUDA(B.head.reserve(2 * (size_t)n)); TSL_CUDA(B.excl.reserve(2 * (size_t)n));
    TSL_CUDA(B.pc.reserve((size_t)nlm + 1)); TSL_CUDA(B.pair_off.reserve((size_t)nlm + 1));
    if (nlm == 0 || n == 0) {   // nothing of this type: empty CSR structures the kernels can still index
      LAUNCH(zero_int_kernel<<<1, 32, 0, st>>>(obs_ptr.p, 1)); LAUNCH(zero_int_kernel<<<1, 32, 0, st>>>(slot_ptr.p, 1)); LAUNCH(zero_int_kernel<<<1, 32, 0, st>>>(ent_ptr.p, 1));
      if (nlm) { LAUNCH(zero_int_kernel<<<grid((size_t)nlm + 1, 256), 256, 0, st>>>(obs_ptr.p, nlm + 1)); LAUNCH(zero_int_kernel<<<grid((size_t)nlm + 1, 256), 256, 0, st>>>(slot_ptr.p, nlm + 1)); }
      return TSLAM_OK;   // counts stay 0 (cnt was cleared)
    }
    size_t tb = temp_bytes;
    LAUNCH(lm_flag_kernel<<<grid(nlm, 256), 256, 0, st>>>(nlm, B.lu.p, in.lm_fixed, B.f.p));
    TSL_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, B.f.p, B.pre.p, nlm, st)); ++g_launches;
    LAUNCH(lm_free_kernel<<<grid(nlm, 256), 256, 0, st>>>(nlm, B.f.p, B.pre.p, B.lmfree.p, v_gl.p, n_free_dev));
    LAUNCH(obs_kernel<<<grid(n, 256), 256, 0, st>>>(n, in.cam, in.host, in.lm, act, X.camslot_d.p, B.lmfree.p, n_free_dev, cs.p, hs.p, ls.p, fmask,
                                                    B.key_lm.p, B.iota.p, B.ent_key.p));
    // observations by landmark (stable: ascending observation index inside a landmark)
    tb = temp_bytes;
    TSL_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, B.key_lm.p, B.key_lm_s.p, B.iota.p, obs.p, n, 0, bits_for((unsigned long long)nlm), st)); ++g_launches;
    LAUNCH(lower_bound_kernel<unsigned><<<grid((size_t)nlm + 1, 256), 256, 0, st>>>(B.key_lm_s.p, nullptr, n, nlm + 1, obs_ptr.p));
    // (landmark, camera slot, obs << 1 | role) entries -> slots
    tb = temp_bytes;
    TSL_CUDA(cub::DeviceRadixSort::SortKeys(temp.p, tb, B.ent_key.p, B.ent_key_s.p, 2 * n, 0, CODE_BITS + CAM_BITS + bits_for((unsigned long long)nlm), st)); ++g_launches;
    LAUNCH(count_valid_kernel<<<1, 32, 0, st>>>(B.ent_key_s.p, 2 * n, n_ent_dev));
    LAUNCH(slot_head_kernel<<<grid(2 * (size_t)n, 256), 256, 0, st>>>(B.ent_key_s.p, 2 * n, n_ent_dev, B.head.p));
    tb = temp_bytes;
    TSL_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, B.head.p, B.excl.p, 2 * n, st)); ++g_launches;
    LAUNCH(slot_emit_kernel<<<grid(2 * (size_t)n, 256), 256, 0, st>>>(B.ent_key_s.p, 2 * n, n_ent_dev, B.head.p, B.excl.p, slot_cam.p, slot_lm.p, ent_ptr.p, ent.p, n_slots_dev));
    LAUNCH(lower_bound_kernel<int><<<grid((size_t)nlm + 1, 256), 256, 0, st>>>(slot_lm.p, n_slots_dev, 0, nlm + 1, slot_ptr.p));
    LAUNCH(pair_count_kernel<<<grid((size_t)nlm + 1, 256), 256, 0, st>>>(nlm + 1, n_free_dev, slot_ptr.p, B.pc.p));
    tb = temp_bytes;
    TSL_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, B.pc.p, B.pair_off.p, nlm + 1, st)); ++g_launches;
    LAUNCH(pair_total_kernel<<<1, 32, 0, st>>>(B.pair_off.p, n_free_dev, n_pairs_dev));
    TSL_CHECK_LAUNCH();
    return TSLAM_OK;
  };
  Counts* C = cnt.p;
  TypeIn inP{np, NP, d->p_cam.p, d->p_host.p, d->p_lm.p, d->rho_fixed.p}, inT{nt, NL, d->t_cam.p, d->t_host.p, d->t_plane.p, d->theta_fixed.p};
  if ((rc = type_pass(inP, BP, X.p_active.p, &C->nl, &C->n_ent_p, &C->nsp, &C->npairs_p, X.p_cs, X.p_hs, X.p_ls, nullptr, X.vp_gl, X.vp_obs_ptr, X.vp_obs,
                      X.sp_ptr, X.sp_cam, X.sp_lm, X.spe_ptr, X.spe))) return rc;
  if ((rc = type_pass(inT, BT, X.t_active.p, &C->npl, &C->n_ent_t, &C->nst, &C->npairs_t, X.t_cs, X.t_hs, X.t_ls, X.t_fmask.p, X.vt_gl, X.vt_obs_ptr, X.vt_obs,
                      X.st_ptr, X.st_cam, X.st_lm, X.ste_ptr, X.ste))) return rc;

  // ---- non-zero blocks: dense (a, b) flag table -> ids in (a, b) order ----
  const size_t K2 = (size_t)K * K;
  const int tn_cap = (6 * K + 63) / 64 + 128;   // + one padding tile per node of the layout at most
  DevBuf<int> flag, pre;
  DevBuf<uint8_t> tile_nz_d;
  TSL_CUDA(flag.reserve(K2)); TSL_CUDA(pre.reserve(K2)); TSL_CUDA(tile_nz_d.reserve((size_t)tn_cap * tn_cap));
  TSL_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int) * K2, st));
  TSL_CUDA(cudaMemsetAsync(tile_nz_d.p, 0, (size_t)tn_cap * tn_cap, st));
  if (np) LAUNCH(mark_direct_kernel<<<grid(np, 256), 256, 0, st>>>(np, X.p_cs.p, X.p_hs.p, X.p_active.p, C, flag.p));
  if (nt) LAUNCH(mark_direct_kernel<<<grid(nt, 256), 256, 0, st>>>(nt, X.t_cs.p, X.t_hs.p, X.t_active.p, C, flag.p));
  if (np && NP) LAUNCH(mark_schur_kernel<<<grid(NP, 128), 128, 0, st>>>(NP, &C->nl, X.sp_ptr.p, X.sp_cam.p, C, flag.p));
  if (nt && NL) LAUNCH(mark_schur_kernel<<<grid(NL, 128), 128, 0, st>>>(NL, &C->npl, X.st_ptr.p, X.st_cam.p, C, flag.p));
  if (multi) {   // global block pattern
    if ((rc = comm_allreduce_sum_i32(ctx, flag.p, K2))) return rc;
    LAUNCH(clamp01_kernel<<<grid(K2, 256), 256, 0, st>>>(flag.p, (int)K2));
  }
  {
    size_t tb = temp_bytes;
    TSL_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, flag.p, pre.p, (int)K2, st)); ++g_launches;
  }
  LAUNCH(block_count_kernel<<<grid(K2, 256), 256, 0, st>>>((int)K2, flag.p, pre.p, C, X.doff.p, tile_nz_d.p, tn_cap));
  TSL_CHECK_LAUNCH();
  // ---- the one round trip: counts + tile pattern ----
  DevBuf<int> gfree;   // multi-GPU: global number of free landmarks of each type (summary + termination test need the global count)
  int gfree_h[2] = {0, 0};
  if (multi) {
    TSL_CUDA(gfree.reserve(2));
    TSL_CUDA(cudaMemcpyAsync(gfree.p, &C->nl, sizeof(int), cudaMemcpyDeviceToDevice, st));
    TSL_CUDA(cudaMemcpyAsync(gfree.p + 1, &C->npl, sizeof(int), cudaMemcpyDeviceToDevice, st));
    if ((rc = comm_allreduce_sum_i32(ctx, gfree.p, 2))) return rc;
    TSL_CUDA(cudaMemcpyAsync(gfree_h, gfree.p, sizeof(gfree_h), cudaMemcpyDeviceToHost, st));
  }
  Counts hc;
  std::vector<uint8_t> tile_h((size_t)tn_cap * tn_cap);
  TSL_CUDA(cudaMemcpyAsync(&hc, C, sizeof(Counts), cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(tile_h.data(), tile_nz_d.p, tile_h.size(), cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  auto T1 = std::chrono::steady_clock::now();
  X.nc = hc.nc; X.nl = multi ? gfree_h[0] : hc.nl; X.npl = multi ? gfree_h[1] : hc.npl; X.nvp = hc.nl; X.nvt = hc.npl; X.nsp = hc.nsp; X.nst = hc.nst; X.nblk = hc.nblk; X.noff = hc.nblk - hc.nc;
  X.est_entries = (long long)hc.npairs_p + hc.npairs_t + 3LL * ((long long)np + nt);
  X.n = 6 * X.nc; X.npad = hc.npad;
  X.Tn = chol_workspace_dims(X.npad, &X.ld, &X.rows);
  {   // tile-level symbolic factorisation + level schedule on the host (Tn^2 flags), uploaded like on the host path
    if (!ctx->host_arena)
      ctx->host_arena = new Arena([](size_t n) -> void* { void* q = nullptr; return cudaHostAlloc(&q, n, cudaHostAllocDefault) == cudaSuccess ? q : nullptr; },
                                  [](void* q) { cudaFreeHost(q); });
    CholHost H;
    try { chol_symbolic_in_arena(X.npad, tile_h.data(), *ctx->host_arena, H); } catch (const std::exception& e) { return set_error(TSLAM_ERR_CUDA, "symbolic factorisation failed: %s", e.what()); }
    if ((rc = chol_upload(ctx, H, &X.chol))) return rc;
    TSL_CUDA(cudaStreamSynchronize(st));   // H lives in the arena only until the next analysis
  }
  auto T2 = std::chrono::steady_clock::now();
  // ---- block arrays and gather lists ----
  const int nblk = X.nblk;
  TSL_CUDA(X.blk_a.reserve(nblk)); TSL_CUDA(X.blk_b.reserve(nblk)); TSL_CUDA(X.diag_blk.reserve(X.nc)); TSL_CUDA(X.offdiag_blk.reserve(X.noff));
  LAUNCH(block_emit_kernel<<<grid(K2, 256), 256, 0, st>>>(flag.p, pre.p, C, X.blk_a.p, X.blk_b.p, X.diag_blk.p, X.offdiag_blk.p));
  const int kbits = bits_for((unsigned long long)nblk);
  DevBuf<unsigned> dkey, dkey_s, skey, skey_s;
  DevBuf<int> dval;
  DevBuf<u64> sval;
  auto direct_lists = [&](int n, const DevBuf<int>& cs, const DevBuf<int>& hs, const uint8_t* act, DevBuf<int>& ptr, DevBuf<int>& out) -> int {
    TSL_CUDA(ptr.reserve((size_t)nblk + 1)); TSL_CUDA(out.reserve(4 * (size_t)n));
    if (n == 0) { LAUNCH(zero_int_kernel<<<grid((size_t)nblk + 1, 256), 256, 0, st>>>(ptr.p, nblk + 1)); return TSLAM_OK; }
    TSL_CUDA(dkey.reserve(4 * (size_t)n)); TSL_CUDA(dkey_s.reserve(4 * (size_t)n)); TSL_CUDA(dval.reserve(4 * (size_t)n));
    LAUNCH(direct_keys_kernel<<<grid(n, 256), 256, 0, st>>>(n, cs.p, hs.p, act, pre.p, X.diag_blk.p, C, dkey.p, dval.p));
    size_t tb = temp_bytes;
    TSL_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dkey.p, dkey_s.p, dval.p, out.p, 4 * n, 0, kbits, st)); ++g_launches;
    LAUNCH(lower_bound_kernel<unsigned><<<grid((size_t)nblk + 1, 256), 256, 0, st>>>(dkey_s.p, nullptr, 4 * n, nblk + 1, ptr.p));
    return TSLAM_OK;
  };
  auto schur_lists = [&](int npairs, int nlm, int* n_free_dev, const DevBuf<int>& slot_ptr, const DevBuf<int>& slot_cam, const DevBuf<int>& pair_off,
                         DevBuf<int>& ptr, DevBuf<int2>& out) -> int {
    TSL_CUDA(ptr.reserve((size_t)nblk + 1)); TSL_CUDA(out.reserve(npairs));
    if (npairs == 0) { LAUNCH(zero_int_kernel<<<grid((size_t)nblk + 1, 256), 256, 0, st>>>(ptr.p, nblk + 1)); return TSLAM_OK; }
    TSL_CUDA(skey.reserve(npairs)); TSL_CUDA(skey_s.reserve(npairs)); TSL_CUDA(sval.reserve(npairs));
    LAUNCH(schur_keys_kernel<<<grid(nlm, 128), 128, 0, st>>>(nlm, n_free_dev, slot_ptr.p, slot_cam.p, pair_off.p, pre.p, C, skey.p, sval.p));
    size_t tb = 0;   // 64-bit values, data-dependent length: size this call on its own (the stream is idle after the round trip)
    TSL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, skey.p, skey_s.p, sval.p, reinterpret_cast<u64*>(out.p), npairs, 0, kbits, st));
    if (tb > temp_bytes) { TSL_CUDA(cudaStreamSynchronize(st)); TSL_CUDA(temp.reserve(tb + 256)); temp_bytes = tb; }
    tb = temp_bytes;
    TSL_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, skey.p, skey_s.p, sval.p, reinterpret_cast<u64*>(out.p), npairs, 0, kbits, st)); ++g_launches;
    LAUNCH(lower_bound_kernel<unsigned><<<grid((size_t)nblk + 1, 256), 256, 0, st>>>(skey_s.p, nullptr, npairs, nblk + 1, ptr.p));
    return TSLAM_OK;
  };
  if ((rc = direct_lists(np, X.p_cs, X.p_hs, X.p_active.p, X.bdp_ptr, X.bdp))) return rc;
  if ((rc = direct_lists(nt, X.t_cs, X.t_hs, X.t_active.p, X.bdt_ptr, X.bdt))) return rc;
  if ((rc = schur_lists(hc.npairs_p, NP, &C->nl, X.sp_ptr, X.sp_cam, BP.pair_off, X.bsp_ptr, X.bsp))) return rc;
  if ((rc = schur_lists(hc.npairs_t, NL, &C->npl, X.st_ptr, X.st_cam, BT.pair_off, X.bst_ptr, X.bst))) return rc;
  TSL_CHECK_LAUNCH();
  if (d->sharded) {   // global observation index of each local one (final residual scatter)
    TSL_CUDA(X.gsel_p.upload(d->gsel_p.data(), d->gsel_p.size(), st)); TSL_CUDA(X.gsel_t.upload(d->gsel_t.data(), d->gsel_t.size(), st));
  }
  TSL_CUDA(cudaStreamSynchronize(st));   // the scratch buffers of this function are released on return
  auto T3 = std::chrono::steady_clock::now();
  if (lap_ms) {
    lap_ms[0] = std::chrono::duration<double, std::milli>(T1 - T0).count();   // layout, slots, block table (+ round trip)
    lap_ms[1] = std::chrono::duration<double, std::milli>(T2 - T1).count();   // symbolic factorisation (host) + upload
    lap_ms[2] = std::chrono::duration<double, std::milli>(T3 - T2).count();   // block arrays + gather lists
    lap_ms[3] = std::chrono::duration<double, std::milli>(T3 - T0).count();
  }
  return TSLAM_OK;
}

}  // namespace tsl

// ---- test hook: device analysis vs host analysis, array by array ---------------------------------------------------
using namespace tsl;

namespace {
template <typename T, typename Vec>
bool same_as_host(const DevBuf<T>& dev, const Vec& host, size_t n, cudaStream_t st) {
  static_assert(sizeof(T) == sizeof(typename Vec::value_type), "element size");
  if (host.size() < n) return false;
  if (n == 0) return true;
  std::vector<T> h(n);
  if (cudaMemcpyAsync(h.data(), dev.p, n * sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return false;
  return memcmp(h.data(), host.data(), n * sizeof(T)) == 0;
}
}  // namespace

extern "C" int tslam_debug_compare_analysis(tslam_ctx* ctx, const tslam_ba_problem* p, char* report, int report_len) {
  if (!ctx || !p || !report || report_len < 1) return set_error(TSLAM_ERR_ARG, "null argument");
  report[0] = 0;
  TSL_CUDA(cudaSetDevice(ctx->device));
  tslam_dev_problem d;
  int rc = upload_problem(ctx, p, &d, false);
  if (rc) return rc;
  if (!device_analysis_supported(ctx, &d)) return set_error(TSLAM_ERR_ARG, "device analysis does not support this problem / context");
  cudaStream_t st = ctx->stream;
  SolverIndex X;
  double laps[4];
  if ((rc = analyze_structure_device(ctx, &d, X, laps))) return rc;
  // copy what the comparison needs from the device symbolic before the arena is recycled by the host analysis
  const CholSymbolic& sy = X.chol;
  const std::vector<int> item_ptr = sy.item_ptr, target_ptr = sy.target_ptr, panel_ptr = sy.panel_ptr;
  const int nwaves = sy.nwaves;
  IndexView V;
  V.n_cams = d.n_cams; V.n_points = d.n_points; V.n_planes = d.n_planes; V.g_pobs = d.g_pobs; V.g_tobs = d.g_tobs;
  V.cam_fixed = d.h_cam_fixed.data(); V.rho_fixed = d.h_rho_fixed.data(); V.theta_fixed = d.h_theta_fixed.data();
  V.p_cam = d.h_p_cam.data(); V.p_host = d.h_p_host.data(); V.p_lm = d.h_p_lm.data();
  V.t_cam = d.h_t_cam.data(); V.t_host = d.h_t_host.data(); V.t_plane = d.h_t_plane.data();
  V.lp = d.n_pobs; V.lt = d.n_tobs;
  Arena arena([](size_t n) { return malloc(n); }, [](void* q) { free(q); });
  Analysis A;
  try { analyze_structure(V, A, arena); } catch (const std::exception& e) { return set_error(TSLAM_ERR_ARG, "host analysis failed: %s", e.what()); }
  std::string bad;
  auto chk = [&](bool ok, const char* name) { if (!ok) { bad += name; bad += ' '; } };
  chk(X.K == A.K && X.nc == A.nc && X.nl == A.nl && X.npl == A.npl && X.lp == A.lp && X.lt == A.lt && X.nvp == A.nvp && X.nvt == A.nvt && X.nsp == A.nsp &&
      X.nst == A.nst && X.nblk == A.nblk && X.noff == (int)A.offdiag_blk.size() && X.n == A.n && X.npad == A.npad && X.ld == A.ld && X.rows == A.rows && X.Tn == A.Tn, "counts");
  chk(same_as_host(X.doff, A.doff, A.nc, st), "doff");
  chk(same_as_host(X.camslot_d, A.camslot, A.K, st), "camslot");
  chk(same_as_host(X.p_cs, A.p_cs, A.lp, st), "p_cs"); chk(same_as_host(X.p_hs, A.p_hs, A.lp, st), "p_hs"); chk(same_as_host(X.p_ls, A.LP.obs_ls, A.lp, st), "p_ls");
  chk(same_as_host(X.t_cs, A.t_cs, A.lt, st), "t_cs"); chk(same_as_host(X.t_hs, A.t_hs, A.lt, st), "t_hs"); chk(same_as_host(X.t_ls, A.LT.obs_ls, A.lt, st), "t_ls");
  chk(same_as_host(X.p_active, A.p_act, A.lp, st), "p_active"); chk(same_as_host(X.t_active, A.t_act, A.lt, st), "t_active"); chk(same_as_host(X.t_fmask, A.t_fm, A.lt, st), "t_fmask");
  auto side = [&](const char* tag, const LmSide& L, int nv, int ns, DevBuf<int>& v_gl, DevBuf<int>& obs_ptr, DevBuf<int>& obs, DevBuf<int>& slot_ptr, DevBuf<int>& slot_cam,
                  DevBuf<int>& slot_lm, DevBuf<int>& ent_ptr, DevBuf<int>& ent) {
    std::string t(tag);
    chk(same_as_host(v_gl, L.v_gl, nv, st), (t + ".v_gl").c_str());
    chk(same_as_host(obs_ptr, L.obs_ptr, (size_t)nv + 1, st), (t + ".obs_ptr").c_str());
    chk(same_as_host(obs, L.obs, L.obs.size(), st), (t + ".obs").c_str());
    chk(same_as_host(slot_ptr, L.slot_ptr, (size_t)nv + 1, st), (t + ".slot_ptr").c_str());
    chk(same_as_host(slot_cam, L.slot_cam, ns, st), (t + ".slot_cam").c_str());
    chk(same_as_host(slot_lm, L.slot_lm, ns, st), (t + ".slot_lm").c_str());
    chk(same_as_host(ent_ptr, L.ent_ptr, (size_t)ns + 1, st), (t + ".ent_ptr").c_str());
    chk(same_as_host(ent, L.ent, L.ent.size(), st), (t + ".ent").c_str());
  };
  if (bad.empty()) {   // lengths are only meaningful once the counts agree
    side("points", A.LP, A.nvp, A.nsp, X.vp_gl, X.vp_obs_ptr, X.vp_obs, X.sp_ptr, X.sp_cam, X.sp_lm, X.spe_ptr, X.spe);
    side("planes", A.LT, A.nvt, A.nst, X.vt_gl, X.vt_obs_ptr, X.vt_obs, X.st_ptr, X.st_cam, X.st_lm, X.ste_ptr, X.ste);
    chk(same_as_host(X.blk_a, A.blk_a, A.nblk, st), "blk_a"); chk(same_as_host(X.blk_b, A.blk_b, A.nblk, st), "blk_b");
    chk(same_as_host(X.diag_blk, A.diag_blk, A.nc, st), "diag_blk"); chk(same_as_host(X.offdiag_blk, A.offdiag_blk, A.offdiag_blk.size(), st), "offdiag_blk");
    chk(same_as_host(X.bdp_ptr, A.bdp_ptr, (size_t)A.nblk + 1, st), "bdp_ptr"); chk(same_as_host(X.bdp, A.bdp, A.bdp.size(), st), "bdp");
    chk(same_as_host(X.bdt_ptr, A.bdt_ptr, (size_t)A.nblk + 1, st), "bdt_ptr"); chk(same_as_host(X.bdt, A.bdt, A.bdt.size(), st), "bdt");
    chk(same_as_host(X.bsp_ptr, A.bsp_ptr, (size_t)A.nblk + 1, st), "bsp_ptr"); chk(same_as_host(X.bsp, A.bsp, A.bsp.size(), st), "bsp");
    chk(same_as_host(X.bst_ptr, A.bst_ptr, (size_t)A.nblk + 1, st), "bst_ptr"); chk(same_as_host(X.bst, A.bst, A.bst.size(), st), "bst");
    chk(nwaves == A.chol.nwaves && item_ptr.size() == A.chol.item_ptr.size() && std::equal(item_ptr.begin(), item_ptr.end(), A.chol.item_ptr.begin()) &&
        std::equal(target_ptr.begin(), target_ptr.end(), A.chol.target_ptr.begin()) && std::equal(panel_ptr.begin(), panel_ptr.end(), A.chol.panel_ptr.begin()), "chol.schedule");
    chk(same_as_host(X.chol.items, A.chol.items, A.chol.items.size(), st), "chol.items"); chk(same_as_host(X.chol.targets, A.chol.targets, A.chol.targets.size(), st), "chol.targets");
    chk(same_as_host(X.chol.src, A.chol.src, A.chol.src.size(), st), "chol.src"); chk(same_as_host(X.chol.below, A.chol.below, A.chol.below.size(), st), "chol.below");
  }
  snprintf(report, report_len, "%s", bad.c_str());
  TSL_CUDA(cudaStreamSynchronize(st));
  return TSLAM_OK;
}
