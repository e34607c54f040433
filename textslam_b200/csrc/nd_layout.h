// Camera order and tile-aligned layout of the reduced camera system — ONE definition shared by the host analysis
// (analysis.cpp) and the device analysis (analysis_dev.cu), so that both produce the same permutation by construction.
//
// Keyframes are temporally ordered and co-visibility is banded (bandwidth bw cameras), so the free cameras form a chain.
// The chain is cut by 2^L - 1 separators of sep_c >= bw + 1 cameras into 2^L leaves (a complete binary tree of nested
// dissection). Elimination order: all leaves, then the separators by height. Every node starts on a 64-column tile
// boundary of the dense reduced matrix (the columns between a node's last camera and the next boundary are identity
// padding): nodes of one height never share a tile, so the tile elimination DAG of chol.cu has depth
//   L * tiles(separator) + tiles(leaf)
// instead of T. Separators use whole tiles (2 tiles = 21 cameras for bw = 20), leaves take what is left; L is chosen to
// minimise that depth. (Round 1 used units of 32 cameras = 3 tiles for leaves and separators alike and put the remainder
// last: 14 waves on the 500-keyframe global BA; this layout needs 10.)
#pragma once
#if defined(__CUDACC__)
#define TSL_HD __host__ __device__
#else
#define TSL_HD
#endif

namespace tsl {

struct NdPlan {
  int levels;      // L (0 = natural order)
  int sep_c;       // cameras per separator
  int n_sep, n_leaf;
  int leaf_base, leaf_extra;   // leaf j has leaf_base + (j < leaf_extra) cameras
};

TSL_HD inline int nd_tiles(int cams) { return (6 * cams + 63) / 64; }

// Chooses the plan for nc free cameras and co-visibility bandwidth bw. levels == 0 -> keep the natural order.
TSL_HD inline NdPlan nd_plan(int nc, int bw) {
  NdPlan best; best.levels = 0; best.sep_c = 0; best.n_sep = 0; best.n_leaf = 1; best.leaf_base = nc; best.leaf_extra = 0;
  if (nc < 128) return best;
  const int sep_tiles = nd_tiles(bw + 1);
  const int sep_c = (64 * sep_tiles) / 6;
  int best_depth = nd_tiles(nc);
  for (int L = 1; L <= 6; ++L) {
    const int n_leaf = 1 << L, n_sep = n_leaf - 1;
    const long long rem = (long long)nc - (long long)n_sep * sep_c;
    if (rem < (long long)n_leaf * 4) break;            // leaves of fewer than 4 cameras: not worth another level
    const int leaf_max = (int)((rem + n_leaf - 1) / n_leaf);
    const int depth = L * sep_tiles + nd_tiles(leaf_max);
    if (depth < best_depth) {
      best_depth = depth;
      best.levels = L; best.sep_c = sep_c; best.n_sep = n_sep; best.n_leaf = n_leaf;
      best.leaf_base = (int)(rem / n_leaf); best.leaf_extra = (int)(rem % n_leaf);
    }
  }
  return best;
}

TSL_HD inline int nd_node_count(const NdPlan& P) { return P.n_leaf + P.n_sep; }
TSL_HD inline int nd_leaf_size(const NdPlan& P, int j) { return P.leaf_base + (j < P.leaf_extra ? 1 : 0); }
// cameras in leaves 0..j-1
TSL_HD inline int nd_leaves_before(const NdPlan& P, int j) { return j * P.leaf_base + (j < P.leaf_extra ? j : P.leaf_extra); }

// Node k of the ELIMINATION order -> its camera range [nat_start, nat_start + size) in the natural (temporal) order.
// k < n_leaf: leaf k. Then the separators by height h = 0 .. L-1; the separators of height h are those whose in-order
// index i (chain position: leaf 0, sep 0, leaf 1, sep 1, ...) has exactly h trailing one bits: i = 2^h - 1 + m 2^(h+1).
TSL_HD inline void nd_node(const NdPlan& P, int k, int* nat_start, int* size) {
  if (k < P.n_leaf) {
    *nat_start = nd_leaves_before(P, k) + k * P.sep_c;
    *size = nd_leaf_size(P, k);
    return;
  }
  int r = k - P.n_leaf, h = 0;
  while (true) {
    const int cnt = 1 << (P.levels - 1 - h);
    if (r < cnt) break;
    r -= cnt; ++h;
  }
  const int i = (1 << h) - 1 + r * (1 << (h + 1));
  *nat_start = nd_leaves_before(P, i + 1) + i * P.sep_c;
  *size = P.sep_c;
}

}  // namespace tsl
