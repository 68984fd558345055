"""Deterministic cost model of the two search methods on one B200 -- the replacement of the reference's wall-clock
threshold tuner (rii/rii.py:403-486, `estimate_best_threshold_function`).

The reference times `query_linear` against `query_ivf` on the host, doubles |S| until the inverted index wins, bisects and
fits a line thre_{|S|} = f(L).  On the GPU a single call is dominated by launch latency (a few microseconds), so wall-clock
bisection fits noise.  Here both methods are costed from what they move and launch:

    seconds = launches * T_LAUNCH + T_TABLE + lookups / LOOKUP_RATE + bytes / BUILD_BW

* lookups: one per code byte of every candidate (the scan engines are bound by the shared-memory lookup rate);
  rows are padded to 32 or 64 bytes on the streaming engine.
* linear, no subset : N candidates.                       IVF, no subset: nlist centers + L candidates.
* linear over target_ids: + a compact copy of the S target rows (read S ids + S rows, write S rows).
* IVF over target_ids   : + the sub-index of the members (key generation, a stable radix sort of S (list, id) pairs,
  bounds, the skew64 copy of the S member rows): ~ SUB_BYTES_PER_TARGET bytes and SUB_LAUNCHES launches per target set,
  shared by the `batch` queries of one call.

The constants are measurements on B200 (profiles/r02_*): see DESIGN.md section 4.  Everything is a pure function of
(N, nlist, M, L, |S|, batch): `method='auto'` is deterministic.
"""

T_LAUNCH = 4.0e-6            # s per kernel launch on an idle stream (launch + drain)
T_TABLE = 5.0e-6             # s: distance table build + merge of a query (K1, ~10 K cycles)
LOOKUP_RATE = 4.5e12         # table lookups / s of the streaming engine (4.7 T persistent IVF batches .. 5.1 T long linear scans)
BUILD_BW = 2.5e12            # bytes / s of the gather / sort passes that build sub-indexes
SUB_BYTES_PER_TARGET = 104   # keys 8 + 3 radix passes x 16 + row gather 32 + skew write 32 (M = 32)
SUB_LAUNCHES = 10
LIN_SUB_LAUNCHES = 3


def _row_bytes(M):
    if 12 <= M <= 32:
        return 32
    if 32 < M <= 64:
        return 64
    return M


class CostModel(object):
    def __init__(self, N, nlist, M):
        self.N, self.nlist, self.M, self.rb = int(N), int(nlist), int(M), _row_bytes(int(M))

    def linear(self, S, subset, batch=1):
        """seconds per query of a linear scan over S candidates (S = N without target_ids)."""
        t = T_LAUNCH + T_TABLE + S * self.rb / LOOKUP_RATE  # one launch: table + scan + merge by the last CTA
        if subset:
            t += (LIN_SUB_LAUNCHES * T_LAUNCH + S * (8 + 2 * self.rb) / BUILD_BW) / batch
        return t

    def ivf(self, L, S, subset, batch=1):
        """seconds per query of an inverted-index search that evaluates L candidates."""
        # one fused launch (table, coarse pass, selection, plan, scan, merge) while every coarse distance fits shared memory
        t = (1 if (batch >= 148 or self.nlist <= 1024) else 3) * T_LAUNCH + T_TABLE + (self.nlist + L) * self.rb / LOOKUP_RATE
        if subset:
            t += (SUB_LAUNCHES * T_LAUNCH + S * (SUB_BYTES_PER_TARGET - 64 + 2 * self.rb) / BUILD_BW) / batch
        return t

    def use_linear(self, S, L, subset, batch=1):
        return self.linear(S, subset, batch) <= self.ivf(L, S, subset, batch)

    def threshold(self, L, batch=1):
        """|S| below which the linear scan over target_ids is the cheaper method (the reference's thre_{|S|} = f(L)).
        Both costs are affine in |S|; when the sub-index costs more per target than scanning it, linear never loses and
        the threshold is N."""
        a1 = self.linear(1, True, batch) - self.linear(0, True, batch)
        b1 = self.ivf(L, 1, True, batch) - self.ivf(L, 0, True, batch)
        a0, b0 = self.linear(0, True, batch), self.ivf(L, 0, True, batch)
        if a1 <= b1:
            return float(self.N)
        return max(0.0, min(float(self.N), (b0 - a0) / (a1 - b1)))


class Threshold(object):
    """Callable kept in `Rii.threshold` (rii/rii.py:147-150 stores a numpy poly1d there): thre_{|S|} = f(L)."""

    def __init__(self, model):
        self.model = model

    def __call__(self, L, batch=1):
        return self.model.threshold(L, batch)

    def __repr__(self):
        m = self.model
        return "CostModel(N=%d, nlist=%d, M=%d): thre(L0)=%.0f" % (m.N, m.nlist, m.M, m.threshold(max(1, m.N // max(1, m.nlist))))
