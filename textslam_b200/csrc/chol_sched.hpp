// Task encoding of the fused reduced-system solve (host builder: analysis.cpp chol_fused_schedule; device: chol_fused.cu).
//
// The tile-level elimination DAG of the reduced camera matrix (64x64 tiles, pattern incl. fill from the symbolic
// factorisation) is flattened into ONE queue of tasks, sorted so that every task comes after everything it depends on and,
// among those orders, by the wave that needs its result (small slack first). CTAs of a single persistent launch pop the
// queue with an atomic counter and wait for their inputs on monotonic counters in a small `sync` array:
//   F  node (1 or 2 consecutive tiles a[, a+1]): factor the diagonal block, publish L_jj^-1 per tile (and L_ba of a pair)
//   S  rows [row0, row0+nrows) of tile (i, j):  X = A_ij L_jj^-T                       (forward solve rides on the b row, i = Tn)
//   U  32x32 quadrant q of tile (i, k):          A_ik -= sum_j X_ij X_kj^T over the source tiles j of one wave
//   B  tile j of the backward solve:             x_j = L_jj^-T (y_j - sum_i L_ij^T x_i)
// Progress never depends on how many CTAs are resident: a task only waits for tasks that were popped before it.
#pragma once

namespace tsl {

constexpr int F_TASK_INTS = 16;
enum { FT_F = 0, FT_S = 1, FT_U = 2, FT_B = 3 };
// common header
constexpr int FK_TYPE = 0, FK_DEP0 = 1, FK_DEP1 = 2, FK_SIG = 3, FK_SIGINC = 4;
// F: first tile, tiles in the node (1 | 2), sync index of xdone(b, a) for a pair (else -1), sync index of fin[first tile]
constexpr int FK_F_TILE = 5, FK_F_NT = 6, FK_F_XBA = 7, FK_F_FIN = 8;
// S: column tile j, row tile i, first row, rows
constexpr int FK_S_J = 5, FK_S_I = 6, FK_S_ROW0 = 7, FK_S_NROWS = 8;
// U: target tile (i, k), quadrant (2 qi + qk; 4 = the whole tile, signals all four quadrant counters), source range in f_srcs
// (bit 30 of a source = second tile of a pair: waited for after the sources without it). Dependencies: one or two per source
// (X_ij, and X_kj when k != i) in source order, then the earlier updates of the target
constexpr int FK_U_I = 5, FK_U_K = 6, FK_U_Q = 7, FK_U_SRC0 = 8, FK_U_SRC1 = 9;
// B: tile j, range of its row tiles in f_below, 1 if the first of them is the partner tile of a pair (its x is waited for last)
constexpr int FK_B_J = 5, FK_B_BEL0 = 6, FK_B_BEL1 = 7, FK_B_PARTNER = 8;

// sync array layout: [0] queue head, [1] abort flag, [FS_FIN0 + j] L_jj^-1 published, [bx0 + j] x_j final,
// [xd0 + tile id] rows of X_ij finished (64 = complete), [uq0 + 4 tile id + q] update tasks finished on that quadrant
constexpr int FS_HEAD = 0, FS_ABORT = 1, FS_FIN0 = 8;

}  // namespace tsl
