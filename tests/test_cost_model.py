"""The deterministic threshold model that replaces the reference's wall-clock tuner (rii/rii.py:403-486)."""
import numpy as np

from rii_b200.cost_model import CostModel, Threshold


def test_no_subset_prefers_the_inverted_index_when_it_scans_less():
    m = CostModel(N=1000000, nlist=1000, M=32)
    assert not m.use_linear(m.N, 32000, subset=False)          # C2: 33 k lookups rows vs 1 M
    assert m.use_linear(m.N, m.N, subset=False)                # L = N: IVF scans everything plus the centers
    small = CostModel(N=10000, nlist=100, M=32)                # C1-sized: both are launch bound; the scan is tiny
    assert small.linear(small.N, False) < 3e-5 and small.ivf(100, small.N, False) < 3e-5


def test_threshold_is_a_crossover_and_deterministic():
    for batch in (1, 256, 4096):
        m = CostModel(N=100000000, nlist=10000, M=64)
        t = Threshold(m)
        for L in (10000, 100000, 1000000):
            s = t(L, batch)
            assert s == Threshold(CostModel(100000000, 10000, 64))(L, batch)
            assert 0 <= s <= m.N
            if 0 < s < m.N:  # linear below the crossover, ivf above it
                assert m.use_linear(int(s * 0.9), L, True, batch)
                assert not m.use_linear(int(s * 1.1) + 1, L, True, batch)
        ts = [t(L, batch) for L in (10000, 100000, 1000000)]
        assert ts == sorted(ts), "a longer candidate list can only favour the linear scan"


def test_single_query_subset_search_is_linear_on_the_gpu():
    """One query pays the whole sub-index build (a radix sort of the target ids): scanning the targets is cheaper."""
    m = CostModel(N=1000000, nlist=1000, M=32)
    assert Threshold(m)(1000, batch=1) == m.N
    # ... while a batch amortises it: the inverted index wins for large target sets
    assert Threshold(m)(1000, batch=4096) < m.N
    assert "thre" in repr(Threshold(m)) and np.isfinite(Threshold(m)(1000))
