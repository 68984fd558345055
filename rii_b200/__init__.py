"""rii_b200: B200-native (sm_100a) implementation of Rii's ADC hot path behind the `rii.Rii` API."""
from .rii import Rii
from .pq import PQ, OPQ

__all__ = ["Rii", "PQ", "OPQ"]
__version__ = "0.2.12"
