"""Executable specification (CPU, numpy) of the v4 scan engine's data layout and arithmetic, checked against the oracle:

  * the skew64 layout (rii_b200/csrc/scan_stream.cuh, k_skew64_build):
        window[b][s][i] = stream_s[32 b - (s & 31) + i],   stream_s = rows s, 64 + s, 128 + s, ... back to back
  * the table arrangement: table h (h < M / 32), column c < 64 holds sub-space (32 h + c - 32) mod M, and lane l looks
    up column t + 32 - l at step t, i.e. 32 different sub-spaces (= 32 different banks) per warp instruction
  * the predicate-free accumulation  acc = acc * keep_t + v,  out = acc * sel_t + out  (keep_t = 0 and sel_t = 1 only at
    the lane's row boundary t == l), which must reproduce the reference's sequential fp32 sum (src/rii.h:386-394) BIT
    FOR BIT.

The CUDA kernels are tested against the oracle on the GPU (tests/test_gpu_parity.py); this file pins the *design* on
any machine and documents it in 60 lines of numpy."""
import numpy as np
import pytest

from _util import O, bits, synth


def skew64_build(codes):
    """numpy restatement of k_skew64_build for ONE segment: (n, M) codes -> (blocks, 64 streams, 32) windows."""
    n, M = codes.shape
    H = M // 32
    G = (n + 63) // 64
    nb = (G + 1) * H
    out = np.zeros((nb, 64, 32), np.uint8)
    for s in range(64):
        rows = codes[s::64]                                    # rows s, 64 + s, ...
        stream = rows.reshape(-1)
        lag = s & 31
        padded = np.concatenate([np.zeros(lag, np.uint8), stream, np.zeros(nb * 32 - lag - len(stream), np.uint8)])
        out[:, s, :] = padded.reshape(nb, 32)
    return out


def walk(windows, T, n, M):
    """One warp walks every block of the segment the way k_scan_stream32 does; returns the n distances (float32)."""
    H = M // 32
    Ks = T.shape[1]
    f32 = np.float32
    # tables: table h, column c holds sub-space (32 h + c - 32) mod M
    tab = np.zeros((H, Ks, 64), f32)
    for h in range(H):
        for c in range(64):
            tab[h, :, c] = T[(32 * h + c - 32) % M]
    lanes = np.arange(32)
    dist = np.full(n, np.nan, f32)
    acc = np.zeros((2, 32), f32)                                # [x / y stream][lane]
    out = np.zeros((2, 32), f32)
    prev_group = -1
    banks_ok = True
    for b in range(windows.shape[0]):
        h = b % H
        for t in range(32):
            col = t + 32 - lanes                                # the lane's table column at this step
            banks_ok &= len(set((col % 32).tolist())) == 32     # 32 lanes -> 32 different banks: conflict free
            for y in range(2):
                ks = windows[b, 32 * y + lanes, t]
                v = tab[h, ks, col]
                if h == 0:
                    keep = (lanes != t).astype(f32)
                    sel = (lanes == t).astype(f32)
                    out[y] = (acc[y] * sel + out[y]).astype(f32)    # x * {0, 1} is exact: one rounding, like an FMA
                    acc[y] = (acc[y] * keep + v).astype(f32)
                else:
                    acc[y] = (acc[y] + v).astype(f32)
        if h == 0:                                              # finished: the rows of the PREVIOUS group
            if prev_group >= 0:
                for y in range(2):
                    r = 64 * prev_group + 32 * y + lanes
                    ok = r < n
                    dist[r[ok]] = out[y][ok]
            out[:] = 0
            prev_group = b // H if b // H < (n + 63) // 64 else -1
    assert banks_ok
    return dist


@pytest.mark.parametrize("M,n", [(32, 1), (32, 64), (32, 65), (32, 1000), (64, 63), (64, 130), (64, 700)])
def test_skew64_walk_reproduces_the_sequential_sum(M, n):
    D, Ks = 4 * M, 256
    cw, codes, Q = synth(D, M, Ks, n, 2, seed=7 * M + n)
    windows = skew64_build(codes)
    for q in Q:
        T = O.dtable(q, cw, 16).reshape(M, Ks)
        got = walk(windows, T, n, M)
        exp = O.adist_all(T, codes)
        assert not np.isnan(got).any()
        assert np.array_equal(bits(got), bits(exp)), "skew64 walk differs from the oracle's sequential ADC sum"


def test_skew64_is_a_permutation_of_the_codes_plus_padding():
    cw, codes, _ = synth(128, 32, 256, 333, 1, seed=1)
    w = skew64_build(codes)
    assert w.shape == ((333 + 63) // 64 + 1, 64, 32)
    # un-skew stream 37: bytes lag .. lag + 32 * rows
    s, lag = 37, 37 & 31
    rows = codes[s::64]
    stream = w[:, s, :].reshape(-1)[lag:lag + rows.size]
    assert np.array_equal(stream, rows.reshape(-1))
    assert int(w.astype(np.int64).sum()) == int(codes.astype(np.int64).sum())   # everything else is zero padding
