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


# ---------------------------------------------------------------------------------------------------------------------
# The walk over SEVERAL segments by several warps (the ST_ISSUE / ST_STAGE logic of k_scan_stream32): flattened groups
# split into contiguous warp ranges, H blocks per group, H drain blocks at the end of a segment or of a range, the
# group's descriptor (flattened group, row-valid bits) travelling with its first block and used when the rows complete.
def issue_blocks(segs_groups, takes, f0, f_end, H):
    """Block sequence of one warp: list of (segment, block in segment, descriptor or None)."""
    gcum = np.cumsum(segs_groups)
    seg = int(np.searchsorted(gcum, f0, side="right"))
    out, cur_f, hb, drain_left = [], f0, 0, 0
    g0 = lambda j: int(gcum[j - 1]) if j else 0
    while True:
        g = cur_f - g0(seg)
        if drain_left == 0:
            if cur_f >= f_end:
                break
            dsc = (cur_f, seg, g) if hb == 0 else None
            out.append((seg, g * H + hb, dsc))
            hb += 1
            if hb == H:
                hb = 0
                cur_f += 1
                if cur_f == f_end or cur_f == gcum[seg]:
                    drain_left = H
        else:
            out.append((seg, g * H + (H - drain_left), None))
            drain_left -= 1
            if drain_left == 0:
                if cur_f < f_end:
                    seg += 1
                else:
                    break
    return out


def walk_segments(seg_windows, takes, T, M, nwarps):
    H, Ks, f32 = M // 32, T.shape[1], np.float32
    tab = np.zeros((H, Ks, 64), f32)
    for h in range(H):
        for c in range(64):
            tab[h, :, c] = T[(32 * h + c - 32) % M]
    lanes = np.arange(32)
    groups = [(t + 63) // 64 for t in takes]
    Gtot = sum(groups)
    per = (Gtot + nwarps - 1) // nwarps
    got = {}
    for w in range(nwarps):
        f0, f_end = min(w * per, Gtot), min((w + 1) * per, Gtot)
        if f_end <= f0:
            continue
        blocks = issue_blocks(groups, takes, f0, f_end, H)
        assert len(blocks) % H == 0
        acc, out = np.zeros((2, 32), f32), np.zeros((2, 32), f32)
        d_last = None
        for pos, (seg, blk, dsc) in enumerate(blocks):
            h = pos % H
            assert h == blk % H, "a block's half-row index must equal its pipeline-stage parity"
            win = seg_windows[seg][blk]
            for t in range(32):
                col = t + 32 - lanes
                for y in range(2):
                    v = tab[h, win[32 * y + lanes, t], col]
                    if h == 0:
                        out[y] = (acc[y] * (lanes == t) + out[y]).astype(f32)
                        acc[y] = (acc[y] * (lanes != t) + v).astype(f32)
                    else:
                        acc[y] = (acc[y] + v).astype(f32)
            if h == 0:
                if d_last is not None:
                    _, s2, g2 = d_last
                    for y in range(2):
                        for l in range(32):
                            r = 64 * g2 + 32 * y + l
                            if r < takes[s2]:
                                assert (s2, r) not in got, "a candidate was emitted twice"
                                got[(s2, r)] = out[y][l]
                out[:] = 0
                d_last = dsc
    return got


@pytest.mark.parametrize("M,nwarps", [(32, 1), (32, 6), (32, 12), (64, 3), (64, 8)])
def test_segment_walk_emits_every_candidate_once_with_the_exact_distance(M, nwarps):
    D, Ks = 4 * M, 256
    rng = np.random.default_rng(M + nwarps)
    lens = [1, 64, 65, 130, 7, 200, 64, 333]
    takes = [1, 64, 65, 100, 7, 129, 10, 333]          # whole lists and lists cut mid-way (the plan's last segment)
    cw, codes, Q = synth(D, M, Ks, sum(lens), 1, seed=3 * M + nwarps)
    T = O.dtable(Q[0], cw, 16).reshape(M, Ks)
    seg_codes, o = [], 0
    for n in lens:
        seg_codes.append(codes[o:o + n])
        o += n
    seg_windows = [skew64_build(c) for c in seg_codes]
    got = walk_segments(seg_windows, takes, T, M, nwarps)
    assert len(got) == sum(takes)
    for s, (c, take) in enumerate(zip(seg_codes, takes)):
        exp = O.adist_all(T, c[:take])
        g = np.array([got[(s, r)] for r in range(take)], np.float32)
        assert np.array_equal(bits(g), bits(exp)), "segment %d" % s


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: the same walk as a list of RUNS (ST_ISSUE of scan_stream.cuh / PS_ISSUE of scan_persist.cuh): inside a run the
# issue logic is a pointer increment and a counter (run_left), the segment tables are read by next_run() only, and the
# row-valid bits are 3 except for the last group of a run that ends with its segment (last_bits from the take count).
def issue_blocks_runs(segs_groups, takes, f0, f_end, H):
    """Block sequence of one warp from the run-based logic: list of (segment, block in segment, descriptor or None) with
    descriptor = (flattened group, segment, group in segment, valid-bits function of the lane)."""
    gcum = np.cumsum(segs_groups)
    out = []
    if f_end <= f0:
        return out
    seg = int(np.searchsorted(gcum, f0, side="right"))
    seg_g0, seg_gend = (int(gcum[seg - 1]) if seg else 0), int(gcum[seg])
    cur_f, hb, run_left, bp, tail = f0, 0, 0, None, 64
    sb = int(np.searchsorted(gcum, f_end - 1, side="right"))
    nblk = ((f_end - f0) + (sb - seg + 1)) * H
    for _ in range(nblk):
        if run_left == 0:  # next_run()
            if cur_f == seg_gend:
                seg += 1
                seg_g0, seg_gend = seg_gend, int(gcum[seg])
            end = min(seg_gend, f_end)
            run_left = (end - cur_f + 1) * H
            bp = (cur_f - seg_g0) * H
            tail = takes[seg] - (seg_gend - seg_g0 - 1) * 64 if end == seg_gend else 64
        if run_left <= H:
            d = None
        elif H == 2 and hb:
            d, hb = None, 0
            cur_f += 1
        else:
            g = cur_f - seg_g0
            last = run_left <= 2 * H
            d = (cur_f, seg, g, (tail if last else 64))
            if H == 1:
                cur_f += 1
            else:
                hb = 1
        out.append((seg, bp, d))
        bp += 1
        run_left -= 1
    return out


@pytest.mark.parametrize("H", [1, 2])
def test_run_based_walk_issues_the_same_blocks_and_valid_rows(H):
    rng = np.random.default_rng(40 + H)
    for trial in range(300):
        nseg = int(rng.integers(1, 9))
        lens = [int(rng.integers(1, 400)) for _ in range(nseg)]
        takes = list(lens)
        if rng.random() < 0.5:
            takes[-1] = int(rng.integers(1, lens[-1] + 1))      # the plan's last segment may be cut mid-way
        groups = [(t + 63) // 64 for t in takes]
        G = sum(groups)
        nw = int(rng.choice([1, 3, 6, 11, 12]))
        per = (G + nw - 1) // nw
        for w in range(nw):
            f0, f_end = min(w * per, G), min((w + 1) * per, G)
            old = issue_blocks(groups, takes, f0, f_end, H) if f_end > f0 else []
            new = issue_blocks_runs(groups, takes, f0, f_end, H)
            assert [(s, b) for s, b, _ in old] == [(s, b) for s, b, _ in new], (trial, w)
            for (s, b, d0), (_, _, d1) in zip(old, new):
                assert (d0 is None) == (d1 is None)
                if d0 is not None:
                    f, s2, g2 = d0
                    assert (f, s2, g2) == d1[:3]
                    for y in range(2):                           # valid bits == "row < take" of the per-block logic
                        for l in (0, 1, 31):
                            r = 64 * g2 + 32 * y + l
                            assert (r < takes[s2]) == (32 * y + l < d1[3]), (trial, w, f, y, l)


def test_top1_selection_rule_equals_the_global_minimum_under_dist_id():
    """The topk = 1 instantiations keep, per warp, the best (distance, position) of the groups they walk: a later candidate
    replaces it only with a strictly smaller distance -- or, on an exact tie in ANOTHER segment, with a smaller id (ids ascend
    with the position inside a segment, not across segments).  The merge takes the minimum (distance, id) over the warps.
    Model of that rule against the brute-force minimum, with heavy ties."""
    rng = np.random.default_rng(7)
    for trial in range(400):
        nseg = int(rng.integers(1, 7))
        lens = [int(rng.integers(1, 150)) for _ in range(nseg)]
        total = sum(lens)
        ids = [np.sort(rng.choice(10 * total, n, replace=False)) for n in lens]       # ascending inside a segment
        all_ids = np.concatenate(ids)
        if len(set(all_ids.tolist())) != total:
            continue
        dist = [rng.integers(0, 4, n).astype(np.float32) for n in lens]               # 4 distinct values: ties everywhere
        groups = [(n + 63) // 64 for n in lens]
        G = sum(groups)
        gcum = np.cumsum(groups)
        nw = int(rng.choice([1, 2, 5, 11]))
        per = (G + nw - 1) // nw
        finals = []
        for w in range(nw):
            f0, f_end = min(w * per, G), min((w + 1) * per, G)
            best = None  # (dist, pos, seg, id)
            for f in range(f0, f_end):
                s = int(np.searchsorted(gcum, f, side="right"))
                g = f - (int(gcum[s - 1]) if s else 0)
                rows = [r for r in range(64 * g, min(64 * g + 64, lens[s]))]
                if best is not None:
                    rows_pass = [r for r in rows if dist[s][r] <= best[0]]
                else:
                    rows_pass = rows
                if not rows_pass:
                    continue
                mn = min(dist[s][r] for r in rows_pass)
                # lowest position at that distance: x rows (lane) before y rows (32 + lane) == ascending row inside the group
                r = min(r for r in rows_pass if dist[s][r] == mn)
                if best is None or mn < best[0]:
                    best = (mn, (f, r), s, int(ids[s][r]))
                elif s != best[2] and int(ids[s][r]) < best[3]:
                    best = (mn, (f, r), s, int(ids[s][r]))
            if best is not None:
                finals.append((best[0], best[3]))
        got = min(finals)
        flat_d = np.concatenate(dist)
        o = np.lexsort((all_ids, flat_d))[0]
        assert got == (flat_d[o], int(all_ids[o])), trial
