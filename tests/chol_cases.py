"""Shared helpers of the reduced-system solver tests: schedule export, numpy replay of the task queue, test matrices."""
import ctypes as C

import numpy as np

from textslam_b200._lib import lib, check

NB, HB = 64, 32
FT_F, FT_S, FT_U, FT_B = 0, 1, 2, 3


def get_schedule(n, tile_nz):
    L = lib()
    counts = (C.c_int32 * 6)()
    tz = np.ascontiguousarray(tile_nz, dtype=np.uint8)
    p8 = tz.ctypes.data_as(C.POINTER(C.c_uint8))
    check(L.tslam_debug_chol_schedule(n, p8, counts, None, 0, None, 0, None, 0, None, 0))
    nt, nd, ns, nb, nsync, Tn = list(counts)
    tasks = np.zeros((nt, 16), np.int32); deps = np.zeros((max(nd, 1), 2), np.int32)
    srcs = np.zeros(max(ns, 1), np.int32); below = np.zeros(max(nb, 1), np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    check(L.tslam_debug_chol_schedule(n, p8, counts, ip(tasks), nt, ip(deps), nd, ip(srcs), ns, ip(below), nb))
    return tasks, deps, srcs, below, nsync, Tn


def replay(S, b, tile_nz):
    n = S.shape[0]
    tasks, deps, srcs, below, nsync, Tn = get_schedule(n, tile_nz)
    ld = Tn * NB
    A = np.zeros(((Tn + 1) * NB, ld))
    A[:n, :n] = np.tril(S)
    for r in range(n, ld):
        A[r, r] = 1.0
    A[Tn * NB, :n] = b
    sync = np.zeros(nsync, np.int64)
    Linv = np.zeros((Tn, NB, NB))
    x = np.zeros(ld)
    T = lambda i, j: A[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]
    n_by_type = [0, 0, 0, 0]
    for t, tk in enumerate(tasks):
        typ = tk[0]
        n_by_type[typ] += 1
        for e in range(tk[1], tk[2]):
            assert sync[deps[e, 0]] >= deps[e, 1], f"task {t} (type {typ}, {tk[5:10]}) waits for sync[{deps[e, 0]}] >= {deps[e, 1]}, has {sync[deps[e, 0]]}"
        if typ == FT_F:
            a, nt = tk[5], tk[6]
            for tt in range(nt):
                j = a + tt
                D = np.tril(T(j, j)); D = D + np.tril(D, -1).T
                Lj = np.linalg.cholesky(D)
                Linv[j] = np.linalg.inv(Lj)
                if nt == 2 and tt == 0:
                    X = T(a + 1, a) @ Linv[a].T
                    T(a + 1, a)[:] = X
                    Tb = T(a + 1, a + 1); Tb -= X @ X.T
                    sync[tk[7]] += 64
                sync[tk[8] + tt] += 1
        elif typ == FT_S:
            j, i, r0, nr = tk[5], tk[6], tk[7], tk[8]
            blk = A[i * NB + r0:i * NB + r0 + nr, j * NB:(j + 1) * NB]
            blk[:] = blk @ Linv[j].T
        elif typ == FT_U:
            i, k, q = tk[5], tk[6], tk[7]
            if q == 4:   # whole tile
                ri, rk = slice(i * NB, (i + 1) * NB), slice(k * NB, (k + 1) * NB)
            else:
                qi, qk = q >> 1, q & 1
                ri, rk = slice(i * NB + HB * qi, i * NB + HB * (qi + 1)), slice(k * NB + HB * qk, k * NB + HB * (qk + 1))
            Cq = A[ri, k * NB + (0 if q == 4 else HB * (q & 1)):k * NB + (NB if q == 4 else HB * ((q & 1) + 1))]
            phases = [srcs[e] >> 30 for e in range(tk[8], tk[9])]
            assert phases == sorted(phases), "second tiles of pairs must come after the first tiles"
            for e in range(tk[8], tk[9]):
                j = srcs[e] & 0x3fffffff
                Cq -= A[ri, j * NB:(j + 1) * NB] @ A[rk, j * NB:(j + 1) * NB].T
        else:
            j = tk[5]
            tt = A[Tn * NB, j * NB:(j + 1) * NB].copy()
            for e in range(tk[6], tk[7]):
                i = below[e]
                tt -= T(i, j).T @ x[i * NB:(i + 1) * NB]
            x[j * NB:(j + 1) * NB] = Linv[j].T @ tt
        if tk[3] >= 0:
            for sq in range(4 if (typ == FT_U and tk[7] == 4) else 1):
                sync[tk[3] + sq] += tk[4]
    return x[:n], n_by_type, Tn


def spd_with_pattern(rng, n, pat):
    """dense SPD matrix whose 64x64 tile pattern is `pat` (lower, Tn x Tn bool)"""
    Tn = pat.shape[0]
    M = np.zeros((n, n))
    for i in range(Tn):
        for k in range(i + 1):
            if pat[i, k] or i == k:
                r0, r1, c0, c1 = i * NB, min(n, (i + 1) * NB), k * NB, min(n, (k + 1) * NB)
                M[r0:r1, c0:c1] = rng.standard_normal((r1 - r0, c1 - c0)) * 0.3
    S = np.tril(M) + np.tril(M, -1).T
    S += np.diag(np.abs(S).sum(axis=1) + 1.0)
    return S


def nd_pattern(levels, leaf_tiles=2, sep_tiles=2):
    """tile pattern of a banded chain cut by nested dissection (leaves first, separators by height), like nd_layout.h"""
    n_leaf = 1 << levels
    order = []   # (kind, in-order index) in elimination order
    for j in range(n_leaf):
        order.append(("leaf", j))
    for h in range(levels):
        for m in range(1 << (levels - 1 - h)):
            order.append(("sep", (1 << h) - 1 + m * (1 << (h + 1))))
    first = {}
    t = 0
    for node in order:
        first[node] = t
        t += leaf_tiles if node[0] == "leaf" else sep_tiles
    Tn = t
    pat = np.zeros((Tn, Tn), bool)

    def couple(a, na, b, nb):
        for x in range(a, a + na):
            for y in range(b, b + nb):
                pat[max(x, y), min(x, y)] = True

    for node in order:
        couple(first[node], leaf_tiles if node[0] == "leaf" else sep_tiles, first[node], leaf_tiles if node[0] == "leaf" else sep_tiles)
    for j in range(n_leaf):   # leaf j sits between separators j-1 and j of the chain
        for sidx in (j - 1, j):
            if 0 <= sidx < n_leaf - 1:
                couple(first[("leaf", j)], leaf_tiles, first[("sep", sidx)], sep_tiles)
    return pat


def dev_chol_solve(ctx, S, b, pat, mode, reps=1, want_trace=False):
    n = S.shape[0]
    S = np.ascontiguousarray(S, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n); ms = C.c_float(0); info = (C.c_int32 * 4)()
    tz = None if pat is None else np.ascontiguousarray(pat, np.uint8)
    cap = 200000 if want_trace else 0
    trace = np.zeros((max(cap, 1), 16), np.uint64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    check(lib().tslam_dev_chol_solve(ctx._h, n, None if tz is None else tz.ctypes.data_as(C.POINTER(C.c_uint8)), dp(S), dp(b), dp(x), mode, reps,
                                     C.byref(ms), trace.ctypes.data_as(C.POINTER(C.c_uint64)) if want_trace else None, cap, info))
    return x, ms.value, list(info), trace[:info[0]]
