// Host-side structure analysis of a BA problem (pure C++, no CUDA): which parameter blocks are free, the camera
// ordering, the landmark -> (camera slot) incidence, the non-zero 6x6 blocks of the reduced camera system with their
// gather lists, and the tile-level symbolic Cholesky. This is the work ceres::Problem / Program / the Schur ordering
// do when the reference calls ceres::Solve on a freshly built problem (src/optimizer.cc:1222,1602,1840,1982,2209);
// it runs once per tslam_solve call and sits inside the end-to-end time, so it is written as counting sorts over
// flat arrays (no per-element allocation) rather than as containers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace tsl {

struct I2 { int x, y; };   // layout-compatible with CUDA's int2

// Bump allocator behind every vector of the analysis. The solver hands in an arena whose chunks are page-locked host
// memory that lives as long as the context: after the first call there is no malloc, no page fault and no staging copy
// (cudaMemcpyAsync reads the index arrays straight out of it). reset() recycles everything at the next analysis.
class Arena {
 public:
  typedef void* (*ChunkAlloc)(size_t);
  typedef void (*ChunkFree)(void*);
  Arena(ChunkAlloc a, ChunkFree f) : alloc_(a), free_(f) {}
  ~Arena() { for (auto& c : chunks_) free_(c.p); }
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  void reset() { cur_ = 0; off_ = 0; }
  void* allocate(size_t bytes);
  size_t reserved_bytes() const { size_t t = 0; for (auto& c : chunks_) t += c.cap; return t; }

 private:
  struct Chunk { char* p; size_t cap; };
  std::vector<Chunk> chunks_;
  size_t cur_ = 0, off_ = 0;
  ChunkAlloc alloc_; ChunkFree free_;
};
Arena* current_arena();            // arena the calling thread's AVec allocations go to (set by analyze_structure)

template <class T>
struct ArenaAlloc {
  typedef T value_type;
  ArenaAlloc() = default;
  template <class U> ArenaAlloc(const ArenaAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(current_arena()->allocate(n * sizeof(T))); }
  void deallocate(T*, size_t) {}   // recycled wholesale by Arena::reset()
  template <class U> bool operator==(const ArenaAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const ArenaAlloc<U>&) const { return false; }
};
template <class T> using AVec = std::vector<T, ArenaAlloc<T>>;

struct IndexView {          // index arrays of the GLOBAL problem + the locally owned observations
  int n_cams = 0, n_points = 0, n_planes = 0, g_pobs = 0, g_tobs = 0;
  const uint8_t *cam_fixed = nullptr, *rho_fixed = nullptr, *theta_fixed = nullptr;
  const int32_t *p_cam = nullptr, *p_host = nullptr, *p_lm = nullptr, *t_cam = nullptr, *t_host = nullptr, *t_plane = nullptr;
  int lp = 0, lt = 0;                                   // local observation counts
  const int32_t *gsel_p = nullptr, *gsel_t = nullptr;   // global index of each local observation; NULL = identity (not sharded)
};

struct LmSide {   // landmark-side structure of one landmark type (inverse depths or planes) for the locally owned landmarks
  AVec<int> v_gl;               // owned landmark -> global landmark index (ascending)
  AVec<int> obs_ptr, obs;       // CSR: local observations of each owned landmark (ascending)
  AVec<int> obs_ls;             // per local observation: owned landmark or -1
  AVec<int> slot_ptr, slot_cam, slot_lm;   // CSR: distinct camera slots touching each landmark (ascending)
  AVec<int> ent_ptr, ent;       // CSR per slot: (obs << 1 | role) entries, role 0 = observing camera, 1 = host camera
};

struct CholHost {  // tile-level symbolic factorisation + level schedule (see chol.cu)
  int Tn = 0, n = 0, nwaves = 0;
  long long gemm_tiles = 0;
  AVec<int> item_ptr, item2_ptr, target_ptr, panel_ptr;   // per wave ranges
  AVec<I2> items, items2, targets, clear_items;           // single-tile panels, two-tile panels (see analysis.cpp), update targets, all pattern tiles
  AVec<int> src_ptr, src, panels, below_ptr, below;
  // ---- fused schedule (chol_fused.cu): the whole factorisation + both triangular solves as ONE persistent launch whose
  // CTAs pop tasks from a topologically sorted queue and wait on counters in `sync` (zeroed before every solve) ----
  int f_ntasks = 0, f_nsync = 0, f_nunits = 0;
  AVec<int> f_tasks;   // F_TASK_INTS ints per task, see chol_sched.hpp
  AVec<I2> f_deps;     // (index into sync, minimum value) pairs a task waits for
  AVec<int> f_srcs;    // source tiles of the update tasks
  AVec<int> f_below;   // row tiles (< Tn) of the backward-solve tasks
};

struct Analysis {
  int K = 0, nc = 0, nl = 0, npl = 0, lp = 0, lt = 0, nvp = 0, nvt = 0, nsp = 0, nst = 0, nblk = 0;
  int n = 0, npad = 0, ld = 0, rows = 0, Tn = 0;   // n = 6 nc unknowns; npad = columns of the tile-aligned layout (>= n)
  AVec<int> camslot, lmfree_p, lmfree_t;
  AVec<int> doff;                               // first column of each camera slot in the dense reduced matrix (nd_layout.h)
  AVec<int> p_cs, p_hs, t_cs, t_hs;             // per local observation: camera slots (or -1)
  AVec<uint8_t> p_act, t_act, t_fm;
  LmSide LP, LT;
  AVec<int> blk_a, blk_b, diag_blk, offdiag_blk;   // diag_blk[c] = block (c,c); offdiag_blk = ids of the blocks with a < b
  AVec<int> bdp_ptr, bdp, bdt_ptr, bdt, bsp_ptr, bst_ptr;
  AVec<I2> bsp, bst;
  CholHost chol;
  double lap_ms[6] = {0, 0, 0, 0, 0, 0};   // layout+ordering, landmark side, block structure, symbolic, entry lists, total
};

int chol_workspace_dims(int n, int* ld, int* rows);
void chol_symbolic_host(int n, const AVec<uint8_t>& tile_nz, CholHost& H);   // needs a current arena (called by analyze_structure)
// tile-level symbolic factorisation alone (device analysis path): tile_nz = Tn x Tn row-major flags; resets `arena`.
void chol_symbolic_in_arena(int n, const uint8_t* tile_nz, Arena& arena, CholHost& H);
// All vectors of A live in `arena` (reset here first): they stay valid until the next analyze_structure on that arena.
void analyze_structure(const IndexView& V, Analysis& A, Arena& arena);

}  // namespace tsl
