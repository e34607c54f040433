"""CPU property tests of the camera layout shared by the host and the device structure analysis (csrc/nd_layout.h): the node
table must be a partition of the camera chain into tile-aligned nodes whose separators really separate (no two cameras closer
than the bandwidth sit in different subtrees), and the predicted depth must not exceed the natural order's."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "nd_layout.h"
using namespace tsl;
// in-order index of a node of the elimination order: leaves are even positions, separators odd positions of the chain
static int chain_pos(const NdPlan& P, int k, int nat_start) { (void)P; (void)k; return nat_start; }
int main(int argc, char** argv) {
  int bad = 0;
  for (int nc = 1; nc <= 3300; nc += (nc < 300 ? 1 : 37))
    for (int bw : {0, 2, 10, 20, 21, 40, 64, 200}) {
      const NdPlan P = nd_plan(nc, bw);
      if (P.levels == 0) { if (P.n_leaf != 1 || P.leaf_base != nc) { printf("natural-order plan broken nc=%d bw=%d\n", nc, bw); ++bad; } continue; }
      const int nn = nd_node_count(P);
      std::vector<int> owner(nc, -1), start(nn), size(nn);
      int cams = 0, depth_tiles_leaf = 0;
      for (int k = 0; k < nn; ++k) {
        nd_node(P, k, &start[k], &size[k]);
        if (size[k] <= 0 || start[k] < 0 || start[k] + size[k] > nc) { printf("node out of range nc=%d bw=%d k=%d\n", nc, bw, k); ++bad; break; }
        for (int c = start[k]; c < start[k] + size[k]; ++c) { if (owner[c] != -1) { printf("overlap nc=%d bw=%d cam=%d\n", nc, bw, c); ++bad; } owner[c] = k; }
        cams += size[k];
        if (k < P.n_leaf && nd_tiles(size[k]) > depth_tiles_leaf) depth_tiles_leaf = nd_tiles(size[k]);
        if (k >= P.n_leaf && size[k] < bw + 1) { printf("separator narrower than the bandwidth nc=%d bw=%d\n", nc, bw); ++bad; }
      }
      if (cams != nc) { printf("not a partition nc=%d bw=%d (%d)\n", nc, bw, cams); ++bad; }
      for (int c = 0; c < nc; ++c) if (owner[c] < 0) { printf("camera without a node nc=%d bw=%d cam=%d\n", nc, bw, c); ++bad; break; }
      // leaves alternate with separators along the chain: two different leaves are always at least one separator apart
      for (int c = 0; c + 1 < nc; ++c) {
        const int a = owner[c], b = owner[c + 1];
        if (a != b && a < P.n_leaf && b < P.n_leaf) { printf("adjacent leaves nc=%d bw=%d cam=%d\n", nc, bw, c); ++bad; break; }
      }
      const int depth = P.levels * nd_tiles(P.sep_c) + depth_tiles_leaf;
      if (depth >= nd_tiles(nc)) { printf("plan no shallower than the natural order nc=%d bw=%d depth=%d\n", nc, bw, depth); ++bad; }
      if (nn > 127) { printf("node table exceeds the device kernel's 128 entries nc=%d bw=%d\n", nc, bw); ++bad; }
    }
  const NdPlan C5 = nd_plan(498, 20);
  printf("C5 levels=%d sep_c=%d leaves=%d\n", C5.levels, C5.sep_c, C5.n_leaf);
  printf("bad=%d\n", bad);
  return bad != 0;
}
'''


def test_layout_properties(tmp_path):
    src = tmp_path / "nd_check.cpp"
    src.write_text(SRC)
    exe = tmp_path / "nd_check"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "textslam_b200", "csrc"), str(src), "-o", str(exe)])
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "C5 levels=4 sep_c=21 leaves=16" in p.stdout and "bad=0" in p.stdout
