// Structure analysis of a BA problem on the host — see analysis.hpp.
#include "analysis.hpp"
#include "nd_layout.h"
#include "chol_sched.hpp"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace tsl {

namespace {

constexpr int NB = 64;   // Cholesky tile (chol.cu)

// Persistent worker pool: nt - 1 threads created once per process. Between analyses they sleep on a condition variable;
// activate() wakes them and for the ~2 ms of an analysis they spin (then yield) on a generation counter, so a parallel
// region costs a few microseconds instead of a thread start.
class Pool {
 public:
  explicit Pool(int nt) : nt_(nt) {
    for (int t = 1; t < nt_; ++t) th_.emplace_back([this, t] { worker(t); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; active_ = true; }
    cv_.notify_all();
    gen_.fetch_add(1, std::memory_order_release);
    for (auto& x : th_) x.join();
  }
  int threads() const { return serial_ ? 1 : nt_; }
  void activate(bool serial) {
    serial_ = serial || nt_ <= 1;
    if (serial_) return;
    { std::lock_guard<std::mutex> lk(mu_); active_ = true; }
    cv_.notify_all();
  }
  void deactivate() {
    if (serial_) return;
    { std::lock_guard<std::mutex> lk(mu_); active_ = false; }
    gen_.fetch_add(1, std::memory_order_release);   // release the spinners into the sleep path (job_ is empty)
  }
  void set_serial() { serial_ = true; }   // rest of this analysis on the calling thread only (workers keep spinning idle)
  // f(tid, begin, end) over nt contiguous ranges of [0, n); the calling thread takes range 0. Ranges are ordered by tid,
  // which the counting sorts below rely on to keep the serial (generation) order inside every bucket.
  template <class F>
  void ranges(int n, F&& f) {
    if (serial_ || n < 2 * nt_) { f(0, 0, n); return; }
    job_ = [&f, n, this](int t) { f(t, (int)((long long)n * t / nt_), (int)((long long)n * (t + 1) / nt_)); };
    done_.store(0, std::memory_order_relaxed);
    has_job_.store(true, std::memory_order_relaxed);
    gen_.fetch_add(1, std::memory_order_release);
    job_(0);
    int spins = 0;
    while (done_.load(std::memory_order_acquire) != nt_ - 1) if (++spins > 256) std::this_thread::yield();
    has_job_.store(false, std::memory_order_relaxed);
  }

 private:
  void worker(int t) {
    int seen = 0;   // gen_'s value at construction: a worker that starts late must not miss the first job
    for (;;) {
      {   // sleep while no analysis is running
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return active_; });
        if (stop_) return;
      }
      for (;;) {   // spin phase
        int spins = 0, g;
        while ((g = gen_.load(std::memory_order_acquire)) == seen) if (++spins > 4096) std::this_thread::yield();
        seen = g;
        bool act;
        { std::lock_guard<std::mutex> lk(mu_); act = active_; if (stop_) return; }
        if (!act) break;
        if (has_job_.load(std::memory_order_relaxed)) { job_(t); done_.fetch_add(1, std::memory_order_release); }
      }
    }
  }
  int nt_;
  bool serial_ = false;
  std::vector<std::thread> th_;
  std::function<void(int)> job_;
  std::atomic<int> gen_{0}, done_{0};
  std::atomic<bool> has_job_{false};
  std::mutex mu_;
  std::condition_variable cv_;
  bool active_ = false, stop_ = false;
};

Pool& global_pool() {
  static Pool pool([] {
    if (const char* e = getenv("TSLAM_HOST_THREADS")) return std::max(1, atoi(e));
    const unsigned h = std::thread::hardware_concurrency();
    return (int)std::min(8u, std::max(1u, h));
  }());
  return pool;
}

thread_local Arena* g_arena = nullptr;

// One pass over the observations of one landmark type. lm_of_obs[i] = dense index (0..nv) of the landmark of local
// observation i when that landmark is free and the observation active, else -1.
void landmark_pass(Pool& pool, int n_obs, const int* cs, const int* hs, const int* lm_of_obs, int nv, LmSide& S) {
  // counting sort of the observations by landmark; thread t counts its (ordered) observation range into its own
  // histogram, so that observations keep their ascending order inside every landmark
  const int nt = pool.threads();
  AVec<int> cnt((size_t)nt * ((size_t)nv + 1), 0), entc((size_t)nt * ((size_t)nv + 1), 0);
  pool.ranges(n_obs, [&](int t, int i0, int i1) {
    int* c_ = &cnt[(size_t)t * (nv + 1)]; int* e_ = &entc[(size_t)t * (nv + 1)];
    for (int i = i0; i < i1; ++i) {
      const int v = lm_of_obs[i];
      if (v < 0) continue;
      c_[v]++;
      e_[v] += (cs[i] >= 0) + (hs[i] >= 0);
    }
  });
  S.obs_ptr.assign((size_t)nv + 1, 0);
  AVec<int> ent_off((size_t)nv + 1, 0);
  {
    int run = 0, erun = 0;
    for (int v = 0; v < nv; ++v) {
      S.obs_ptr[v] = run; ent_off[v] = erun;
      for (int t = 0; t < nt; ++t) { int& c = cnt[(size_t)t * (nv + 1) + v]; const int k = c; c = run; run += k; erun += entc[(size_t)t * (nv + 1) + v]; }
    }
    S.obs_ptr[nv] = run; ent_off[nv] = erun;
  }
  const size_t n_ent = (size_t)ent_off[nv];
  S.obs.resize((size_t)S.obs_ptr[nv]);
  pool.ranges(n_obs, [&](int t, int i0, int i1) {
    int* cur = &cnt[(size_t)t * (nv + 1)];
    for (int i = i0; i < i1; ++i) { const int v = lm_of_obs[i]; if (v >= 0) S.obs[cur[v]++] = i; }
  });
  // phase 1: per landmark, the (camera slot, obs << 1 | role) keys sorted; count the distinct camera slots
  AVec<uint64_t> keys(n_ent);
  S.slot_ptr.assign((size_t)nv + 1, 0);
  pool.ranges(nv, [&](int, int v0, int v1) {
    for (int v = v0; v < v1; ++v) {
      uint64_t* key = keys.data() + ent_off[v];
      int m = 0;
      for (int e = S.obs_ptr[v]; e < S.obs_ptr[v + 1]; ++e) {
        const int i = S.obs[e];
        if (cs[i] >= 0) key[m++] = ((uint64_t)cs[i] << 32) | (uint32_t)((i << 1) | 0);
        if (hs[i] >= 0) key[m++] = ((uint64_t)hs[i] << 32) | (uint32_t)((i << 1) | 1);
      }
      if (m <= 24) {
        for (int a = 1; a < m; ++a) { const uint64_t k = key[a]; int b = a - 1; while (b >= 0 && key[b] > k) { key[b + 1] = key[b]; --b; } key[b + 1] = k; }
      } else std::sort(key, key + m);
      int ns = 0;
      for (int k = 0; k < m; ++k) ns += (k == 0 || (key[k] >> 32) != (key[k - 1] >> 32));
      S.slot_ptr[v + 1] = ns;
    }
  });
  for (int v = 0; v < nv; ++v) S.slot_ptr[v + 1] += S.slot_ptr[v];
  const size_t n_slots = (size_t)S.slot_ptr[nv];
  S.slot_cam.resize(n_slots); S.slot_lm.resize(n_slots); S.ent.resize(n_ent); S.ent_ptr.resize(n_slots + 1);
  // phase 2: emit slots and their entry ranges (entries of a landmark are contiguous, slots ascending)
  pool.ranges(nv, [&](int, int v0, int v1) {
    for (int v = v0; v < v1; ++v) {
      const uint64_t* key = keys.data() + ent_off[v];
      const int m = ent_off[v + 1] - ent_off[v];
      int s = S.slot_ptr[v] - 1;
      for (int k = 0; k < m; ++k) {
        if (k == 0 || (key[k] >> 32) != (key[k - 1] >> 32)) { ++s; S.slot_cam[s] = (int)(key[k] >> 32); S.slot_lm[s] = v; S.ent_ptr[s] = ent_off[v] + k; }
        S.ent[(size_t)ent_off[v] + k] = (int)(uint32_t)key[k];
      }
    }
  });
  S.ent_ptr[n_slots] = (int)n_ent;
}

// Dense owned-landmark numbering: landmarks that are free (lmfree >= 0) and appear in an active local observation,
// in ascending global order. Fills S.v_gl, S.obs_ls.
int owned_landmarks(Pool& pool, int n_obs, const int32_t* lm, const uint8_t* act, const AVec<int>& lmfree, LmSide& S) {
  const int n_lm = (int)lmfree.size();
  AVec<int> local_of((size_t)n_lm, -1);
  pool.ranges(n_obs, [&](int, int i0, int i1) {   // every thread stores the same value: relaxed atomic stores
    for (int i = i0; i < i1; ++i) if (act[i] && lmfree[lm[i]] >= 0) __atomic_store_n(&local_of[lm[i]], 0, __ATOMIC_RELAXED);
  });
  S.v_gl.clear();
  for (int l = 0; l < n_lm; ++l) if (local_of[l] == 0) { local_of[l] = (int)S.v_gl.size(); S.v_gl.push_back(l); }
  S.obs_ls.resize((size_t)n_obs);
  pool.ranges(n_obs, [&](int, int i0, int i1) {
    for (int i = i0; i < i1; ++i) S.obs_ls[i] = (act[i] && lmfree[lm[i]] >= 0) ? local_of[lm[i]] : -1;
  });
  return (int)S.v_gl.size();
}

struct Laps {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), tl = t0;
  double lap() { auto now = std::chrono::steady_clock::now(); const double ms = std::chrono::duration<double, std::milli>(now - tl).count(); tl = now; return ms; }
  double total() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace

Arena* current_arena() { return g_arena; }

void* Arena::allocate(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;   // 256-byte granules: every vector starts on a DMA / cache-line friendly boundary
  if (bytes == 0) bytes = 256;
  while (cur_ < chunks_.size() && off_ + bytes > chunks_[cur_].cap) { ++cur_; off_ = 0; }
  if (cur_ == chunks_.size()) {
    const size_t cap = std::max(bytes, (size_t)32 << 20);
    void* p = alloc_(cap);
    if (!p) throw std::bad_alloc();
    chunks_.push_back(Chunk{static_cast<char*>(p), cap});
    off_ = 0;
  }
  void* r = chunks_[cur_].p + off_;
  off_ += bytes;
  return r;
}

int chol_workspace_dims(int n, int* ld, int* rows) {
  const int Tn = (n + NB - 1) / NB;
  *ld = Tn * NB;
  *rows = (Tn + 1) * NB;
  return Tn;
}

// tile_nz: Tn x Tn row-major flags of the lower-triangular tile pattern of S (diagonal always set).
// Besides the fill pattern this computes a LEVEL SCHEDULE of the tile elimination DAG: panel j can start once every
// panel k < j with L[j][k] != 0 is finished; panels of one wave are mutually independent, so a wave is three
// launches (factor+solve, update, and later the backward solve) however many panels it holds. With the nested-
// dissection camera order chosen below a banded problem needs ~15 waves instead of T = 47 panel steps.
static void chol_fused_schedule(int Tn, const AVec<uint8_t>& P, const AVec<AVec<int>>& below, CholHost& H);
void chol_symbolic_host(int n, const AVec<uint8_t>& tile_nz, CholHost& H) {
  int ld, rows;
  const int Tn = chol_workspace_dims(n, &ld, &rows);
  H = CholHost();
  H.Tn = Tn; H.n = n;
  const int T1 = Tn + 1;  // + the b tile row (dense)
  AVec<uint8_t> P((size_t)T1 * T1, 0);
  for (int i = 0; i < Tn; ++i)
    for (int k = 0; k <= i; ++k) P[(size_t)i * T1 + k] = (i == k) || tile_nz[(size_t)i * Tn + k];
  for (int k = 0; k < Tn; ++k) P[(size_t)Tn * T1 + k] = 1;
  AVec<AVec<int>> below(Tn);
  for (int j = 0; j < Tn; ++j) {   // symbolic factorisation (fill)
    AVec<int>& nz = below[j];
    for (int i = j + 1; i < T1; ++i) if (P[(size_t)i * T1 + j]) nz.push_back(i);
    for (size_t a = 0; a < nz.size(); ++a)
      for (size_t b = 0; b <= a; ++b) P[(size_t)nz[a] * T1 + nz[b]] = 1;
  }
  // ---- units of the schedule: single tiles, or PAIRS of consecutive coupled tiles (a, a+1) that one CTA factors together
  // (potrf2_trsm2_kernel: L_aa, L_ba, the update of A_bb and L_bb without leaving the SM). In the tile-aligned camera layout
  // (nd_layout.h) every node is two tiles wide, so a pair is a node and a tree level costs one launch pair instead of two.
  // A pair is formed when the second tile has no dependency that finishes later than the first tile's own dependencies.
  AVec<int> wave(Tn, 0);   // per-tile level of the plain (unpaired) schedule, used for the pairing test only
  for (int j = 0; j < Tn; ++j) {
    int w = 0;
    for (int k = 0; k < j; ++k) if (P[(size_t)j * T1 + k]) w = std::max(w, wave[k] + 1);
    wave[j] = w;
  }
  // Measured on the 500-keyframe global BA (profiles/r1_notes.md): 5 waves instead of 10, but a pair CTA runs its two 64-pivot
  // chains back to back (38 us per launch against 2 x 19 us), so the factorisation time is unchanged (283 vs 279 us) — the
  // schedule is kept behind TSLAM_CHOL_PAIR=1 until the pair kernel overlaps the row-tile work with the second chain.
  static const bool pairing = [] { const char* e = getenv("TSLAM_CHOL_PAIR"); return e && e[0] == '1'; }();
  AVec<int> unit_of(Tn, -1), unit_first, unit_size;
  for (int j = 0; j < Tn;) {
    bool pair = pairing && j + 1 < Tn && P[(size_t)(j + 1) * T1 + j];
    if (pair)
      for (int k = 0; k < j; ++k) if (P[(size_t)(j + 1) * T1 + k] && wave[k] >= wave[j]) { pair = false; break; }
    unit_of[j] = (int)unit_first.size();
    if (pair) unit_of[j + 1] = (int)unit_first.size();
    unit_first.push_back(j); unit_size.push_back(pair ? 2 : 1);
    j += pair ? 2 : 1;
  }
  const int nunits = (int)unit_first.size();
  AVec<int> uwave(nunits, 0);
  int nwaves = 0;
  for (int u = 0; u < nunits; ++u) {
    int w = 0;
    for (int t = 0; t < unit_size[u]; ++t) {
      const int j = unit_first[u] + t;
      for (int k = 0; k < unit_first[u]; ++k) if (P[(size_t)j * T1 + k]) w = std::max(w, uwave[unit_of[k]] + 1);
    }
    uwave[u] = w; nwaves = std::max(nwaves, w + 1);
  }
  H.nwaves = nwaves;
  AVec<AVec<int>> wave_units(nwaves);
  for (int u = 0; u < nunits; ++u) wave_units[uwave[u]].push_back(u);
  H.item_ptr.assign(nwaves + 1, 0); H.item2_ptr.assign(nwaves + 1, 0); H.target_ptr.assign(nwaves + 1, 0); H.panel_ptr.assign(nwaves + 1, 0);
  H.src_ptr.push_back(0); H.below_ptr.push_back(0);
  for (int j = 0; j < Tn; ++j) {   // every tile of the factor pattern (cleared before the reduced system is scattered)
    H.clear_items.push_back(I2{j, -1});
    for (int i : below[j]) H.clear_items.push_back(I2{j, i});
  }
  AVec<int> tgt_index((size_t)T1 * T1, -1);
  for (int w = 0; w < nwaves; ++w) {
    const size_t t_begin = H.targets.size();
    AVec<AVec<int>> tsrc;
    auto add_updates = [&](int j, int skip_tile) {   // trailing updates of panel j; targets in column skip_tile are done inside the pair kernel
      const AVec<int>& nz = below[j];
      for (size_t a = 0; a < nz.size(); ++a)
        for (size_t b = 0; b <= a; ++b) {
          const int i = nz[a], k = nz[b];
          if (i == Tn && k == Tn) continue;   // (b row, b row) is never read
          if (k == skip_tile) continue;
          int& ti = tgt_index[(size_t)i * T1 + k];
          if (ti < (int)t_begin) { ti = (int)H.targets.size(); H.targets.push_back(I2{i, k}); tsrc.emplace_back(); }
          tsrc[ti - t_begin].push_back(j);
          ++H.gemm_tiles;
        }
    };
    for (int u : wave_units[w]) {
      const int j = unit_first[u];
      if (unit_size[u] == 1) {
        H.items.push_back(I2{j, -1});
        for (int i : below[j]) H.items.push_back(I2{j, i});
        add_updates(j, -1);
      } else {
        // pair (j, j+1): one CTA per row tile of column j+1 (+ one for the diagonal); x carries bit 30 when the row tile is
        // structurally absent from column j
        H.items2.push_back(I2{j, -1});
        for (int i : below[j + 1]) H.items2.push_back(I2{P[(size_t)i * T1 + j] ? j : (j | (1 << 30)), i});
        add_updates(j, j + 1);
        add_updates(j + 1, -1);
        H.gemm_tiles += 2 * (long long)below[j + 1].size() + 1;   // the in-kernel updates of A_bb and of the row tiles of column j+1
      }
      for (int t = 0; t < unit_size[u]; ++t) {   // backward solve: tiles (i, j) of the factor below the diagonal, excluding the b row
        const int jj = j + t;
        H.panels.push_back(jj);
        for (int i : below[jj]) if (i < Tn) H.below.push_back(i);
        H.below_ptr.push_back((int)H.below.size());
      }
    }
    for (auto& v : tsrc) { for (int j : v) H.src.push_back(j); H.src_ptr.push_back((int)H.src.size()); }
    for (size_t t = t_begin; t < H.targets.size(); ++t) tgt_index[(size_t)H.targets[t].x * T1 + H.targets[t].y] = -1;
    H.item_ptr[w + 1] = (int)H.items.size(); H.item2_ptr[w + 1] = (int)H.items2.size();
    H.target_ptr[w + 1] = (int)H.targets.size(); H.panel_ptr[w + 1] = (int)H.panels.size();
  }
  chol_fused_schedule(Tn, P, below, H);
}

// Fused schedule (chol_sched.hpp, chol_fused.cu). P = T1 x T1 pattern of L incl. fill and the b row, below[j] = its column lists.
static void chol_fused_schedule(int Tn, const AVec<uint8_t>& P, const AVec<AVec<int>>& below, CholHost& H) {
  const int T1 = Tn + 1;
  // ---- compact ids of the pattern tiles + layout of the sync array ----
  AVec<int> tid((size_t)T1 * T1, -1);
  int ntile = 0;
  for (int j = 0; j < Tn; ++j) { tid[(size_t)j * T1 + j] = ntile++; for (int i : below[j]) tid[(size_t)i * T1 + j] = ntile++; }
  const int fin0 = FS_FIN0, bx0 = fin0 + Tn, xd0 = bx0 + Tn, uq0 = xd0 + ntile;
  H.f_nsync = uq0 + 4 * ntile;
  auto XD = [&](int i, int j) { return xd0 + tid[(size_t)i * T1 + j]; };
  auto UQ = [&](int i, int k, int q) { return uq0 + 4 * tid[(size_t)i * T1 + k] + q; };
  // ---- nodes: single tiles, or pairs (a, a+1) of coupled consecutive tiles whose second tile waits for nothing that
  // finishes after the first tile's own inputs (every node of the tile-aligned nested-dissection layout is such a pair) ----
  AVec<int> wave(Tn, 0);
  for (int j = 0; j < Tn; ++j) { int w = 0; for (int k = 0; k < j; ++k) if (P[(size_t)j * T1 + k]) w = std::max(w, wave[k] + 1); wave[j] = w; }
  AVec<int> unit_of(Tn, -1), unit_first, unit_size;
  for (int j = 0; j < Tn;) {
    bool pair = j + 1 < Tn && P[(size_t)(j + 1) * T1 + j];
    if (pair) for (int k = 0; k < j; ++k) if (P[(size_t)(j + 1) * T1 + k] && wave[k] >= wave[j]) { pair = false; break; }
    unit_of[j] = (int)unit_first.size();
    if (pair) unit_of[j + 1] = (int)unit_first.size();
    unit_first.push_back(j); unit_size.push_back(pair ? 2 : 1);
    j += pair ? 2 : 1;
  }
  const int nunits = (int)unit_first.size();
  H.f_nunits = nunits;
  AVec<int> uwave(nunits, 0);
  int nwaves = 0;
  for (int u = 0; u < nunits; ++u) {
    int w = 0;
    for (int t = 0; t < unit_size[u]; ++t) { const int j = unit_first[u] + t; for (int k = 0; k < unit_first[u]; ++k) if (P[(size_t)j * T1 + k]) w = std::max(w, uwave[unit_of[k]] + 1); }
    uwave[u] = w; nwaves = std::max(nwaves, w + 1);
  }
  auto twave = [&](int tile) { return tile >= Tn ? nwaves : uwave[unit_of[tile]]; };   // the b row is "needed" after everything else
  // ---- tasks with sort keys ----
  // Priority class of a task = 4 n + c, n = the wave whose factorisation (c = 0: the update lands in a diagonal block of a node
  // of wave n; c = 1: that node's F task itself) or whose row solves (c = 2) consume its result; a row solve S inherits the
  // most urgent class among the updates it feeds. The queue is sorted by (class, stage inside the chain S(a) -> U(col b from a)
  // -> S(b) -> U -> F, source wave): every dependency of a task has a class <= its own and, if equal, an earlier stage, so
  // the order is topological, and what the next diagonal factorisation waits for is popped before everything that only
  // later row solves need (an F task that sits behind hundreds of such updates in the queue is popped ~10 us late).
  struct Tk { long long key; int rec[F_TASK_INTS]; int s0, s1; };
  AVec<Tk> T;
  AVec<int> usrc;
  auto new_task = [&](int type) {
    Tk t; t.key = 0; for (int& v : t.rec) v = 0; t.rec[FK_TYPE] = type; t.rec[FK_SIG] = -1; t.s0 = t.s1 = 0; T.push_back(t); return (int)T.size() - 1;
  };
  auto KEY = [](int cls, int stage, long long sec) { return ((long long)cls << 44) | ((long long)stage << 40) | sec; };
  const int BIG = 4 * (nwaves + 2);
  AVec<int> ps((size_t)ntile, BIG);   // class of the row solve of tile (i, j)
  AVec<int> rows_in_wave(nwaves, 0);
  for (int j = 0; j < Tn; ++j) rows_in_wave[uwave[unit_of[j]]] += (int)below[j].size();
  static const int srows_crit = [] { const char* e = getenv("TSLAM_CHOL_SROWS"); const int v = e ? atoi(e) : 16; return (v == 16 || v == 32 || v == 64) ? v : 16; }();
  AVec<int> uq_total((size_t)4 * ntile, 0);
  struct Upd { int i, k, wv, cls; int s0, s1; bool mid, early; };   // target tile, source wave, class, sources in usrc
  AVec<Upd> upds;
  AVec<int> tgt_index((size_t)T1 * T1, -1);
  auto upd_class = [&](int i, int k) { return 4 * twave(k) + ((i < Tn && unit_of[i] == unit_of[k]) ? 0 : 2); };
  for (int w = 0; w < nwaves; ++w) {
    const size_t u_begin = upds.size();
    AVec<AVec<int>> tsrc;
    for (int u = 0; u < nunits; ++u) {
      if (uwave[u] != w) continue;
      const int a = unit_first[u], nt = unit_size[u], b = a + nt - 1;
      for (int t2 = nt - 1; t2 >= 0; --t2) {   // second tile first: the class of S(b, i) is what the updates of column b from tile a inherit
        const int j = a + t2;
        const AVec<int>& nz = below[j];
        for (size_t x = 0; x < nz.size(); ++x)
          for (size_t y = 0; y <= x; ++y) {
            const int i = nz[x], k = nz[y];
            if (i == Tn && k == Tn) continue;                 // (b row, b row) is never read
            if (nt == 2 && j == a && k == b) {
              if (i == b) continue;                           // A_bb -= L_ba L_ba^T happens inside F
              const int cls = ps[tid[(size_t)i * T1 + b]];    // consumer: S(b, i)
              upds.push_back(Upd{i, k, w, cls, (int)usrc.size(), (int)usrc.size() + 1, true, false}); usrc.push_back(a);
              int& pa = ps[tid[(size_t)i * T1 + a]]; pa = std::min(pa, cls);
              continue;
            }
            const int cls = upd_class(i, k);
            int& pi = ps[tid[(size_t)i * T1 + j]]; pi = std::min(pi, cls);
            int& pk = ps[tid[(size_t)k * T1 + j]]; pk = std::min(pk, cls);
            // sources that are SECOND tiles of the wave's pairs carry bit 30: the task takes the first tiles' contributions
            // (complete while the nodes still factor their second tiles) before it waits for those
            int& ti = tgt_index[(size_t)i * T1 + k];
            if (ti < 0) { ti = (int)tsrc.size(); tsrc.emplace_back(); upds.push_back(Upd{i, k, w, cls, 0, 0, false, false}); }
            tsrc[ti].push_back((nt == 2 && j == b) ? (j | (1 << 30)) : j);
          }
      }
    }
    {   // flatten the per-target source lists of this wave (ascending tile order)
      size_t x = 0;
      for (size_t q = u_begin; q < upds.size(); ++q) {
        Upd& U = upds[q];
        if (U.mid) continue;
        std::sort(tsrc[x].begin(), tsrc[x].end());   // bit 30 sorts the second tiles behind the first ones
        U.s0 = (int)usrc.size(); for (int j : tsrc[x]) usrc.push_back(j); U.s1 = (int)usrc.size(); ++x;
        tgt_index[(size_t)U.i * T1 + U.k] = -1;
      }
    }
    // F and S tasks of this wave (the classes of its row solves are final now)
    const int nr = 4 * rows_in_wave[w] <= 160 ? 16 : (2 * rows_in_wave[w] <= 160 ? 32 : 64);
    for (int u = 0; u < nunits; ++u) {
      if (uwave[u] != w) continue;
      const int a = unit_first[u], nt = unit_size[u], b = a + nt - 1;
      {
        const int t = new_task(FT_F);
        Tk& k = T[t];
        k.key = KEY(4 * w + 1, 0, a);
        k.rec[FK_F_TILE] = a; k.rec[FK_F_NT] = nt; k.rec[FK_F_XBA] = nt == 2 ? XD(b, a) : -1; k.rec[FK_F_FIN] = fin0 + a;
      }
      for (int t2 = 0; t2 < nt; ++t2) {
        const int j = a + t2;
        for (int i : below[j]) {
          if (nt == 2 && j == a && i == b) continue;   // L_ba is produced inside F
          const int cls = ps[tid[(size_t)i * T1 + j]];
          // what the next diagonal factorisation waits for goes in 16-row pieces; the rest as coarse as the machine allows
          const int step = i == Tn ? 64 : ((cls & 3) == 0 ? std::min(nr, srows_crit) : nr);
          for (int r0 = 0; r0 < 64; r0 += step) {
            const int t = new_task(FT_S);
            Tk& k = T[t];
            k.key = KEY(cls, (nt == 2 && t2 == 1) ? 2 : 0, ((long long)w << 24) | ((long long)i << 12) | ((long long)j << 3) | (r0 >> 4));
            k.rec[FK_S_J] = j; k.rec[FK_S_I] = i; k.rec[FK_S_ROW0] = r0; k.rec[FK_S_NROWS] = i == Tn ? 8 : step;   // b row: only row 0 carries data
            k.rec[FK_SIG] = XD(i, j); k.rec[FK_SIGINC] = step;
          }
        }
      }
    }
  }
  for (const Upd& U : upds) {
    // what a diagonal factorisation waits for goes in 32x32 quadrants (four CTAs per tile, short tasks); everything else — the
    // bulk of the flops, needed only by later row solves — as whole tiles: a quarter of the task overheads and half the loads
    const bool whole = U.i < Tn && U.i != U.k && (U.cls & 3) != 0 && !U.mid;
    const int nq0 = whole ? 1 : (U.i == Tn ? 2 : 4);   // b row: only the upper quadrants (rows 0..31) carry data
    for (int q = 0; q < nq0; ++q) {
      if (U.i == U.k && q == 1) continue;   // upper-right quadrant of a diagonal tile is never read
      const int t = new_task(FT_U);
      Tk& k = T[t];
      k.key = KEY(U.cls, U.mid ? 1 : 3, ((long long)U.wv << 24) | ((long long)U.i << 12) | ((long long)U.k << 2) | q);
      k.rec[FK_U_I] = U.i; k.rec[FK_U_K] = U.k; k.rec[FK_U_Q] = whole ? 4 : q; k.s0 = U.s0; k.s1 = U.s1;   // sources resolved below (usrc)
      k.rec[FK_SIG] = UQ(U.i, U.k, whole ? 0 : q); k.rec[FK_SIGINC] = 1;
      for (int qq = (whole ? 0 : q); qq < (whole ? 4 : q + 1); ++qq) uq_total[(size_t)4 * tid[(size_t)U.i * T1 + U.k] + qq]++;
    }
  }
  for (int j = Tn - 1; j >= 0; --j) {   // backward solve, last tile first
    const int t = new_task(FT_B);
    T[t].key = KEY(BIG + 1, 0, Tn - 1 - j);
    T[t].rec[FK_B_J] = j; T[t].rec[FK_SIG] = bx0 + j; T[t].rec[FK_SIGINC] = 1;
  }
  // ---- final order, then the dependencies (the order of the updates on one quadrant is their queue order) ----
  AVec<int> ord(T.size());
  for (size_t q = 0; q < ord.size(); ++q) ord[q] = (int)q;
  std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return T[x].key < T[y].key; });
  AVec<int> uq_seen((size_t)4 * ntile, 0);
  H.f_tasks.clear(); H.f_deps.clear(); H.f_srcs.clear(); H.f_below.clear();
  H.f_tasks.reserve(T.size() * F_TASK_INTS);
  auto need_uq = [&](int i, int k, int q) { const int n = uq_total[(size_t)4 * tid[(size_t)i * T1 + k] + q]; if (n > 0) H.f_deps.push_back(I2{UQ(i, k, q), n}); };
  for (int o : ord) {
    Tk& k = T[o];
    k.rec[FK_DEP0] = (int)H.f_deps.size();
    switch (k.rec[FK_TYPE]) {
      case FT_F: {
        const int a = k.rec[FK_F_TILE], nt = k.rec[FK_F_NT];
        for (int t2 = 0; t2 < nt; ++t2) { need_uq(a + t2, a + t2, 0); need_uq(a + t2, a + t2, 2); need_uq(a + t2, a + t2, 3); }
        if (nt == 2) for (int q = 0; q < 4; ++q) need_uq(a + 1, a, q);
      } break;
      case FT_S: {
        const int j = k.rec[FK_S_J], i = k.rec[FK_S_I], r0 = k.rec[FK_S_ROW0], nr = k.rec[FK_S_NROWS];
        H.f_deps.push_back(I2{fin0 + j, 1});
        for (int h = r0 / 32; h <= (r0 + nr - 1) / 32; ++h) { need_uq(i, j, 2 * h); need_uq(i, j, 2 * h + 1); }
      } break;
      case FT_U: {
        const int i = k.rec[FK_U_I], kk = k.rec[FK_U_K], q = k.rec[FK_U_Q];
        k.rec[FK_U_SRC0] = (int)H.f_srcs.size();
        for (int e = k.s0; e < k.s1; ++e) {   // per source, in source order: the task waits for them chunk by chunk
          const int j = usrc[e] & 0x3fffffff;
          H.f_srcs.push_back(usrc[e]);
          H.f_deps.push_back(I2{XD(i, j), 64});
          if (kk != i) H.f_deps.push_back(I2{XD(kk, j), 64});
        }
        k.rec[FK_U_SRC1] = (int)H.f_srcs.size();
        // then the earlier updates of the same quadrant(s): only the final read-modify-write waits for them
        for (int qq = (q == 4 ? 0 : q); qq < (q == 4 ? 4 : q + 1); ++qq) {
          int& seen = uq_seen[(size_t)4 * tid[(size_t)i * T1 + kk] + qq];
          if (seen > 0) H.f_deps.push_back(I2{UQ(i, kk, qq), seen});
          ++seen;
        }
      } break;
      case FT_B: {
        const int j = k.rec[FK_B_J];
        // first what is final before any x_i is (the task stages it while it waits): L_jj^-1, y_j, the L_ij tiles; then the x_i
        H.f_deps.push_back(I2{fin0 + j, 1});
        H.f_deps.push_back(I2{XD(Tn, j), 64});
        k.rec[FK_B_BEL0] = (int)H.f_below.size();
        for (int i : below[j]) if (i < Tn) { H.f_below.push_back(i); H.f_deps.push_back(I2{XD(i, j), 64}); }
        k.rec[FK_B_BEL1] = (int)H.f_below.size();
        const bool partner = j + 1 < Tn && unit_of[j + 1] == unit_of[j];   // first tile of a pair: below[j] starts with its second tile
        for (int i : below[j]) if (i < Tn && !(partner && i == j + 1)) H.f_deps.push_back(I2{bx0 + i, 1});
        if (partner) H.f_deps.push_back(I2{bx0 + j + 1, 1});
        k.rec[FK_B_PARTNER] = partner ? 1 : 0;
      } break;
    }
    k.rec[FK_DEP1] = (int)H.f_deps.size();
    for (int v : k.rec) H.f_tasks.push_back(v);
  }
  H.f_ntasks = (int)T.size();
}

namespace {
struct ArenaScope {   // makes `a` the calling thread's arena for the lifetime of the scope
  Arena* prev;
  explicit ArenaScope(Arena* a) : prev(g_arena) { g_arena = a; }
  ~ArenaScope() { g_arena = prev; }
};
struct PoolScope {
  Pool& p;
  PoolScope(Pool& pool, bool serial) : p(pool) { p.activate(serial); }
  ~PoolScope() { p.deactivate(); }
};
std::mutex g_analysis_mutex;   // one analysis at a time per process (the pool and its job slot are shared)
}  // namespace

void chol_symbolic_in_arena(int n, const uint8_t* tile_nz, Arena& arena, CholHost& H) {
  std::lock_guard<std::mutex> serialise(g_analysis_mutex);
  arena.reset();
  ArenaScope arena_scope(&arena);
  int ld, rows;
  const int Tn = chol_workspace_dims(n, &ld, &rows);
  AVec<uint8_t> tz(tile_nz, tile_nz + (size_t)Tn * Tn);
  chol_symbolic_host(n, tz, H);
}

void analyze_structure(const IndexView& V, Analysis& A, Arena& arena) {
  std::lock_guard<std::mutex> serialise(g_analysis_mutex);
  Laps T;
  arena.reset();
  ArenaScope arena_scope(&arena);
  A = Analysis();
  const int K = V.n_cams, GP = V.g_pobs, GT = V.g_tobs;
  Pool& pool = global_pool();
  PoolScope pool_scope(pool, (size_t)GP + 8 * (size_t)GT < 20000);   // waking the workers only pays on the large problems
  A.K = K;
  auto cf = [&](int k) { return V.cam_fixed[k] != 0; };
  // ---- global layout (same on every rank) ----
  AVec<uint8_t> cu(K, 0), lu(V.n_points, 0), pu(V.n_planes, 0);
  AVec<uint8_t> gp_active(GP), gt_active(GT);
  auto mark1 = [](uint8_t& f) { __atomic_store_n(&f, (uint8_t)1, __ATOMIC_RELAXED); };   // same value from every thread
  pool.ranges(GP, [&](int, int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      const int c = V.p_cam[i], h = V.p_host[i], l = V.p_lm[i];
      const bool act = !cf(c) || !cf(h) || !V.rho_fixed[l];
      gp_active[i] = act;
      if (act) { mark1(cu[c]); mark1(cu[h]); mark1(lu[l]); }
    }
  });
  pool.ranges(GT, [&](int, int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      const int c = V.t_cam[i], h = V.t_host[i], l = V.t_plane[i];
      const bool act = !cf(c) || !cf(h) || !V.theta_fixed[l];
      gt_active[i] = act;
      if (act) { mark1(cu[c]); mark1(cu[h]); mark1(pu[l]); }
    }
  });
  A.camslot.assign(K, -1); A.nc = 0;
  for (int k = 0; k < K; ++k) if (cu[k] && !cf(k)) A.camslot[k] = A.nc++;
  A.lmfree_p.assign(V.n_points, -1); A.nl = 0;
  for (int k = 0; k < V.n_points; ++k) if (lu[k] && !V.rho_fixed[k]) A.lmfree_p[k] = A.nl++;
  A.lmfree_t.assign(V.n_planes, -1); A.npl = 0;
  for (int k = 0; k < V.n_planes; ++k) if (pu[k] && !V.theta_fixed[k]) A.lmfree_t[k] = A.npl++;
  const int nc = A.nc;
  A.n = 6 * nc;
  // ---- camera order + tile-aligned layout (nd_layout.h): leaves first, then the separators by height; every node starts on
  // a tile boundary of the dense reduced matrix. Any order is valid, this one only shortens the critical path of the
  // reduced-system factorisation (keyframes are temporally ordered, co-visibility is banded).
  A.doff.assign(nc, 0);
  for (int s = 0; s < nc; ++s) A.doff[s] = 6 * s;
  A.npad = A.n;
  if (nc >= 128) {
    AVec<int> hist((size_t)nc, 0);   // histogram of |slot(cam) - slot(host)| over the active observations
    size_t nd = 0;
    for (int i = 0; i < GP; ++i) if (gp_active[i]) { const int a = A.camslot[V.p_cam[i]], b = A.camslot[V.p_host[i]]; if (a >= 0 && b >= 0) { hist[std::abs(a - b)]++; ++nd; } }
    for (int i = 0; i < GT; ++i) if (gt_active[i]) { const int a = A.camslot[V.t_cam[i]], b = A.camslot[V.t_host[i]]; if (a >= 0 && b >= 0) { hist[std::abs(a - b)]++; ++nd; } }
    int bw = 0;
    if (nd) {   // 98th percentile of the distances (index floor(0.98 (nd-1)) of the sorted list)
      const size_t q = (size_t)(0.98 * (double)(nd - 1));
      size_t acc = 0; int dq = 0;
      for (int dd = 0; dd < nc; ++dd) { acc += hist[dd]; if (acc > q) { dq = dd; break; } }
      bw = 2 * dq;
    }
    const NdPlan P = nd_plan(nc, bw);
    if (P.levels > 0) {
      AVec<int> new_of_old(nc, -1);
      int slot = 0, d = 0;
      for (int k = 0; k < nd_node_count(P); ++k) {
        int s0, sz;
        nd_node(P, k, &s0, &sz);
        d = (d + NB - 1) / NB * NB;
        for (int c = s0; c < s0 + sz; ++c) { new_of_old[c] = slot; A.doff[slot] = d; ++slot; d += 6; }
      }
      A.npad = d;
      for (int k = 0; k < K; ++k) if (A.camslot[k] >= 0) A.camslot[k] = new_of_old[A.camslot[k]];
    }
  }
  A.Tn = chol_workspace_dims(A.npad, &A.ld, &A.rows);
  // camera slots of every global observation
  AVec<int> gp_cs(GP), gp_hs(GP), gt_cs(GT), gt_hs(GT);
  pool.ranges(GP, [&](int, int i0, int i1) { for (int i = i0; i < i1; ++i) { gp_cs[i] = A.camslot[V.p_cam[i]]; gp_hs[i] = A.camslot[V.p_host[i]]; } });
  pool.ranges(GT, [&](int, int i0, int i1) { for (int i = i0; i < i1; ++i) { gt_cs[i] = A.camslot[V.t_cam[i]]; gt_hs[i] = A.camslot[V.t_host[i]]; } });
  A.lap_ms[0] = T.lap();

  // ---- local observations + landmark side ----
  const bool sharded = V.gsel_p != nullptr || V.gsel_t != nullptr;
  const int lp = V.lp, lt = V.lt; A.lp = lp; A.lt = lt;
  AVec<int32_t> lp_lm, lt_lm;
  const int32_t *l_p_lm = V.p_lm, *l_t_lm = V.t_plane;
  if (!sharded) {
    A.p_cs = gp_cs; A.p_hs = gp_hs; A.p_act = gp_active; A.t_cs = gt_cs; A.t_hs = gt_hs; A.t_act = gt_active;
  } else {
    A.p_cs.resize(lp); A.p_hs.resize(lp); A.p_act.resize(lp); lp_lm.resize(lp);
    for (int i = 0; i < lp; ++i) { const int g = V.gsel_p[i]; A.p_cs[i] = gp_cs[g]; A.p_hs[i] = gp_hs[g]; A.p_act[i] = gp_active[g]; lp_lm[i] = V.p_lm[g]; }
    A.t_cs.resize(lt); A.t_hs.resize(lt); A.t_act.resize(lt); lt_lm.resize(lt);
    for (int i = 0; i < lt; ++i) { const int g = V.gsel_t[i]; A.t_cs[i] = gt_cs[g]; A.t_hs[i] = gt_hs[g]; A.t_act[i] = gt_active[g]; lt_lm[i] = V.t_plane[g]; }
    l_p_lm = lp_lm.data(); l_t_lm = lt_lm.data();
  }
  A.t_fm.resize(lt);
  for (int i = 0; i < lt; ++i) A.t_fm[i] = (uint8_t)((A.t_cs[i] >= 0 ? 1 : 0) | (A.t_hs[i] >= 0 ? 2 : 0) | (A.lmfree_t[l_t_lm[i]] >= 0 ? 4 : 0));
  A.nvp = owned_landmarks(pool, lp, l_p_lm, A.p_act.data(), A.lmfree_p, A.LP);
  A.nvt = owned_landmarks(pool, lt, l_t_lm, A.t_act.data(), A.lmfree_t, A.LT);
  landmark_pass(pool, lp, A.p_cs.data(), A.p_hs.data(), A.LP.obs_ls.data(), A.nvp, A.LP);
  landmark_pass(pool, lt, A.t_cs.data(), A.t_hs.data(), A.LT.obs_ls.data(), A.nvt, A.LT);
  A.nsp = (int)A.LP.slot_cam.size(); A.nst = (int)A.LT.slot_cam.size();
  // the block structure needs the slot sets of ALL free landmarks (every rank builds the same reduced matrix layout)
  LmSide GLP, GLT;
  const LmSide *SP = &A.LP, *ST = &A.LT;
  if (sharded) {
    owned_landmarks(pool, GP, V.p_lm, gp_active.data(), A.lmfree_p, GLP);
    owned_landmarks(pool, GT, V.t_plane, gt_active.data(), A.lmfree_t, GLT);
    landmark_pass(pool, GP, gp_cs.data(), gp_hs.data(), GLP.obs_ls.data(), (int)GLP.v_gl.size(), GLP);
    landmark_pass(pool, GT, gt_cs.data(), gt_hs.data(), GLT.obs_ls.data(), (int)GLT.v_gl.size(), GLT);
    SP = &GLP; ST = &GLT;
  }
  A.lap_ms[1] = T.lap();

  // ---- global block structure: unique (a <= b) camera-slot pairs from every active observation / landmark ----
  // dense (a,b) -> block id table when nc^2 is small enough (O(1) insert / lookup); sorted-key fallback otherwise
  const bool dense_tab = (size_t)nc * (size_t)nc <= ((size_t)1 << 24);
  AVec<int> btab;
  AVec<uint64_t> bkeys;
  if (dense_tab) btab.assign((size_t)nc * nc, -1);
  if (dense_tab) {
    // every thread stores the same value (0) into the table: relaxed atomic stores, no ordering needed before the join
    auto mark = [&](int a, int b) { __atomic_store_n(&btab[(size_t)a * nc + b], 0, __ATOMIC_RELAXED); };
    auto direct_keys = [&](int n_obs, const int* cs, const int* hs, const uint8_t* act) {
      pool.ranges(n_obs, [&](int, int i0, int i1) {
        for (int i = i0; i < i1; ++i) if (act[i]) {
          const int c = cs[i], h = hs[i];
          if (c >= 0) mark(c, c);
          if (h >= 0) mark(h, h);
          if (c >= 0 && h >= 0 && c != h) mark(std::min(c, h), std::max(c, h));
        }
      });
    };
    direct_keys(GP, gp_cs.data(), gp_hs.data(), gp_active.data());
    direct_keys(GT, gt_cs.data(), gt_hs.data(), gt_active.data());
    auto schur_keys = [&](const LmSide& L) {
      pool.ranges((int)L.slot_ptr.size() - 1, [&](int, int v0, int v1) {
        for (int v = v0; v < v1; ++v)
          for (int x = L.slot_ptr[v]; x < L.slot_ptr[v + 1]; ++x)
            for (int y = x; y < L.slot_ptr[v + 1]; ++y) mark(L.slot_cam[x], L.slot_cam[y]);
      });
    };
    schur_keys(*SP); schur_keys(*ST);
  } else {
    auto add_key = [&](int a, int b) { bkeys.push_back((uint64_t)a * (uint64_t)nc + (uint64_t)b); };   // a <= b
    auto direct_keys = [&](int n_obs, const int* cs, const int* hs, const uint8_t* act) {
      for (int i = 0; i < n_obs; ++i) if (act[i]) {
        const int c = cs[i], h = hs[i];
        if (c >= 0) add_key(c, c);
        if (h >= 0) add_key(h, h);
        if (c >= 0 && h >= 0 && c != h) add_key(std::min(c, h), std::max(c, h));
      }
    };
    direct_keys(GP, gp_cs.data(), gp_hs.data(), gp_active.data());
    direct_keys(GT, gt_cs.data(), gt_hs.data(), gt_active.data());
    auto schur_keys = [&](const LmSide& L) {
      const int nv = (int)L.slot_ptr.size() - 1;
      for (int v = 0; v < nv; ++v)
        for (int x = L.slot_ptr[v]; x < L.slot_ptr[v + 1]; ++x)
          for (int y = x; y < L.slot_ptr[v + 1]; ++y) add_key(L.slot_cam[x], L.slot_cam[y]);
    };
    schur_keys(*SP); schur_keys(*ST);
  }
  A.blk_a.clear(); A.blk_b.clear(); A.diag_blk.assign(nc, -1);
  if (dense_tab) {
    int nb = 0;
    for (int a = 0; a < nc; ++a)
      for (int b = a; b < nc; ++b)
        if (btab[(size_t)a * nc + b] == 0) { btab[(size_t)a * nc + b] = nb++; A.blk_a.push_back(a); A.blk_b.push_back(b); }
    A.nblk = nb;
  } else {
    std::sort(bkeys.begin(), bkeys.end());
    bkeys.erase(std::unique(bkeys.begin(), bkeys.end()), bkeys.end());
    A.nblk = (int)bkeys.size();
    A.blk_a.resize(A.nblk); A.blk_b.resize(A.nblk);
    for (int b = 0; b < A.nblk; ++b) { A.blk_a[b] = (int)(bkeys[b] / (uint64_t)nc); A.blk_b[b] = (int)(bkeys[b] % (uint64_t)nc); }
  }
  A.offdiag_blk.clear();
  for (int b = 0; b < A.nblk; ++b) { if (A.blk_a[b] == A.blk_b[b]) A.diag_blk[A.blk_a[b]] = b; else A.offdiag_blk.push_back(b); }
  A.lap_ms[2] = T.lap();
  {  // tile pattern of the reduced matrix -> symbolic tile Cholesky
    AVec<uint8_t> tile_nz((size_t)A.Tn * A.Tn, 0);
    for (int b = 0; b < A.nblk; ++b) {
      // block (a,b'), a <= b' lands in rows doff[b']..+5, cols doff[a]..+5 of the lower triangle
      const int r0 = A.doff[A.blk_b[b]] / NB, r1 = (A.doff[A.blk_b[b]] + 5) / NB, c0 = A.doff[A.blk_a[b]] / NB, c1 = (A.doff[A.blk_a[b]] + 5) / NB;
      for (int r = r0; r <= r1; ++r) for (int c = c0; c <= c1; ++c) if (c <= r) tile_nz[(size_t)r * A.Tn + c] = 1;
    }
    chol_symbolic_host(A.npad, tile_nz, A.chol);
  }
  A.lap_ms[3] = T.lap();
  auto blk_of = [&](int a, int b) {   // a <= b
    if (dense_tab) return btab[(size_t)a * nc + b];
    const uint64_t key = (uint64_t)a * (uint64_t)nc + (uint64_t)b;
    return (int)(std::lower_bound(bkeys.begin(), bkeys.end(), key) - bkeys.begin());
  };

  // ---- local gather lists per block: counting sort by block, order inside a block = generation order. Parallel form:
  // thread t counts its (ordered) range into its own histogram, bucket b then starts at ptr[b] + sum_{t' < t} hist[t'][b].
  if ((size_t)pool.threads() * (size_t)A.nblk > ((size_t)1 << 22)) pool.set_serial();
  const int nt = pool.threads();
  AVec<int> hist((size_t)nt * ((size_t)A.nblk + 1));
  auto offsets_from_hist = [&](AVec<int>& ptr) {   // hist[t][b] (counts) -> hist[t][b] (start offsets), ptr = CSR
    ptr.assign((size_t)A.nblk + 1, 0);
    int run = 0;
    for (int b = 0; b < A.nblk; ++b) {
      ptr[b] = run;
      for (int t = 0; t < nt; ++t) { int& h = hist[(size_t)t * (A.nblk + 1) + b]; const int c = h; h = run; run += c; }
    }
    ptr[A.nblk] = run;
    return run;
  };
  auto direct_lists = [&](int n_obs, const int* cs, const int* hs, const uint8_t* act, AVec<int>& ptr, AVec<int>& out) {
    // codes: 0 = J_cam^T J_cam on the diagonal block of the observing camera, 1 = same for the host, 2 / 3 = the
    // off-diagonal block (2: observing slot < host slot, 3: host slot < observing slot)
    AVec<int> key(3 * (size_t)n_obs);
    std::fill(hist.begin(), hist.end(), 0);
    pool.ranges(n_obs, [&](int t, int i0, int i1) {
      int* h_ = &hist[(size_t)t * (A.nblk + 1)];
      for (int i = i0; i < i1; ++i) {
        int* k = &key[3 * (size_t)i];
        k[0] = k[1] = k[2] = -1;
        if (!act[i]) continue;
        const int c = cs[i], h = hs[i];
        if (c >= 0) { k[0] = A.diag_blk[c]; h_[k[0]]++; }
        if (h >= 0) { k[1] = A.diag_blk[h]; h_[k[1]]++; }
        if (c >= 0 && h >= 0) { k[2] = c < h ? blk_of(c, h) : (h < c ? blk_of(h, c) : A.diag_blk[c]); h_[k[2]] += (c == h) ? 2 : 1; }
      }
    });
    out.resize((size_t)offsets_from_hist(ptr));
    pool.ranges(n_obs, [&](int t, int i0, int i1) {
      int* cur = &hist[(size_t)t * (A.nblk + 1)];
      for (int i = i0; i < i1; ++i) {
        const int* k = &key[3 * (size_t)i];
        if (k[0] >= 0) out[cur[k[0]]++] = (i << 2) | 0;
        if (k[1] >= 0) out[cur[k[1]]++] = (i << 2) | 1;
        if (k[2] >= 0) {
          const int c = cs[i], h = hs[i];
          if (c < h) out[cur[k[2]]++] = (i << 2) | 2;
          else if (h < c) out[cur[k[2]]++] = (i << 2) | 3;
          else { out[cur[k[2]]++] = (i << 2) | 2; out[cur[k[2]]++] = (i << 2) | 3; }   // cam == host never happens in the reference (src/optimizer.cc:1397)
        }
      }
    });
  };
  direct_lists(lp, A.p_cs.data(), A.p_hs.data(), A.p_act.data(), A.bdp_ptr, A.bdp);
  direct_lists(lt, A.t_cs.data(), A.t_hs.data(), A.t_act.data(), A.bdt_ptr, A.bdt);
  auto schur_lists = [&](const LmSide& L, AVec<int>& ptr, AVec<I2>& out) {
    const int nv = (int)L.slot_ptr.size() - 1;
    AVec<size_t> pair_off((size_t)nv + 1, 0);
    for (int v = 0; v < nv; ++v) { const size_t m = (size_t)(L.slot_ptr[v + 1] - L.slot_ptr[v]); pair_off[v + 1] = pair_off[v] + m * (m + 1) / 2; }
    AVec<int> key(pair_off[nv]);
    std::fill(hist.begin(), hist.end(), 0);
    pool.ranges(nv, [&](int t, int v0, int v1) {
      int* h_ = &hist[(size_t)t * (A.nblk + 1)];
      size_t e = pair_off[v0];
      for (int v = v0; v < v1; ++v)
        for (int x = L.slot_ptr[v]; x < L.slot_ptr[v + 1]; ++x)
          for (int y = x; y < L.slot_ptr[v + 1]; ++y) { const int b = blk_of(L.slot_cam[x], L.slot_cam[y]); key[e++] = b; h_[b]++; }
    });
    out.resize((size_t)offsets_from_hist(ptr));
    pool.ranges(nv, [&](int t, int v0, int v1) {
      int* cur = &hist[(size_t)t * (A.nblk + 1)];
      size_t e = pair_off[v0];
      for (int v = v0; v < v1; ++v)
        for (int x = L.slot_ptr[v]; x < L.slot_ptr[v + 1]; ++x)
          for (int y = x; y < L.slot_ptr[v + 1]; ++y) out[cur[key[e++]]++] = I2{x, y};
    });
  };
  schur_lists(A.LP, A.bsp_ptr, A.bsp);
  schur_lists(A.LT, A.bst_ptr, A.bst);
  A.lap_ms[4] = T.lap();
  A.lap_ms[5] = T.total();
}

}  // namespace tsl
