"""Executable specification (CPU, numpy) of the IVF candidate plan: the sequential walk of the reference
(src/rii.h:286-322, restated as make_plan in rii_b200/csrc/kernels.cuh) against the prefix-scan form that ONE WARP
computes inside the fused kernel (plan_warp, rii_b200/csrc/scan_stream.cuh).  Both are restated here line by line and
compared on random list-length profiles, including sharded ones (pre / loc) and the flagged regimes (walk beyond w,
empty result).  The CUDA versions are compared with the oracle on the GPU; this pins the derivation on any machine."""
import numpy as np
import pytest


def plan_sequential(f, pre, loc, L, topk, w, nlist):
    """make_plan: returns (J, take_last, flag, local takes per rank j < J)."""
    W = len(f)
    P, J, flag, done, take_last = 0, 0, 0, False, 0
    lts = []
    for j in range(W):
        take = int(f[j])
        if P + f[j] >= L:                       # src/rii.h:302-304
            take, done = L - P, True
        P += take
        lt = min(max(take - int(pre[j]), 0), int(loc[j]))
        lts.append(lt)
        take_last, J = take, j + 1
        if done:
            break
        if j == w - 1 and P >= topk:            # src/rii.h:309
            done = True
            break
    if not done:
        flag = 2 if W >= nlist else 1
    return J, take_last, flag, lts


def plan_parallel(f, pre, loc, L, topk, w, nlist):
    """plan_warp: inclusive prefix F of the list lengths, jL = first rank with F >= L, stop rule, local takes."""
    W = len(f)
    F = np.cumsum(np.asarray(f, np.int64))
    hit = np.nonzero(F >= L)[0]
    jL = int(hit[0]) if len(hit) else W
    F_before = int(F[jL] - f[jL]) if jL < W else 0
    F_w = int(F[w - 1]) if w - 1 < W else -1
    jstop, by_L = -1, False
    if jL < W and jL <= w - 1:
        jstop, by_L = jL, True
    elif w - 1 < W and F_w >= topk:
        jstop = w - 1
    elif jL < W:
        jstop, by_L = jL, True
    J = jstop + 1 if jstop >= 0 else W
    flag = 0 if jstop >= 0 else (2 if W >= nlist else 1)
    lts, take_last = [], 0
    for j in range(J):
        take = L - F_before if (by_L and j == jstop) else int(f[j])
        lts.append(min(max(take - int(pre[j]), 0), int(loc[j])))
        if j == J - 1:
            take_last = take
    return J, take_last, flag, lts


@pytest.mark.parametrize("seed", range(8))
def test_plan_warp_equals_the_sequential_walk(seed):
    rng = np.random.default_rng(seed)
    for _ in range(1500):
        nlist = int(rng.integers(1, 300))
        full = rng.random() < 0.2
        w = int(rng.integers(1, nlist + 1))
        W = nlist if full else w                                   # ranked lists available: w, or all on the re-run
        kind = rng.integers(0, 3)
        f = rng.integers(0, [4, 60, 2000][kind], W)                # sparse (filtered subsets) ... dense lists
        G = int(rng.integers(1, 5))                                # shards: split every list
        cut = np.sort(rng.integers(0, f + 1, (G - 1, W)), axis=0) if G > 1 else np.zeros((0, W), np.int64)
        bounds = np.vstack([np.zeros((1, W), np.int64), cut, f[None, :]])
        rank = int(rng.integers(0, G))
        pre, loc = bounds[rank], bounds[rank + 1] - bounds[rank]
        total = int(f.sum())
        L = int(rng.integers(1, max(2, total + 50)))
        topk = int(rng.integers(1, max(2, min(L, 40) + 1)))
        a = plan_sequential(f, pre, loc, L, topk, w, nlist)
        b = plan_parallel(f, pre, loc, L, topk, w, nlist)
        assert a[2] == b[2], (f, L, topk, w, nlist, a, b)
        if a[2] == 0:                                              # (flagged plans scan nothing: J = 0 downstream)
            assert a == b, (f, L, topk, w, nlist, a, b)


def test_plan_examples():
    f = np.array([10, 10, 10, 10])
    z = np.zeros(4, np.int64)
    # L reached inside the 3rd list
    assert plan_parallel(f, z, f, 25, 1, 4, 100) == (3, 5, 0, [10, 10, 5])
    # w = 2 lists hold >= topk candidates: stop there although L is not reached
    assert plan_parallel(f, z, f, 1000, 5, 2, 100) == (2, 10, 0, [10, 10])
    # fewer than topk candidates in the first w lists, more lists exist: flag 1 (host re-runs with the full ranking)
    assert plan_parallel(np.array([1, 1]), z[:2], np.array([1, 1]), 50, 5, 2, 100)[2] == 1
    # all lists ranked and L never reached with < topk at the w-th: the reference returns nothing (src/rii.h:325)
    assert plan_parallel(np.array([1, 1]), z[:2], np.array([1, 1]), 50, 5, 1, 2)[2] == 2
