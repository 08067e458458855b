"""openkeonspark_b200 — B200-native (sm_100a) knowledge-graph-embedding hot path behind the
OpenKEonSpark Config / Model API.  See DESIGN.md and include/okb200.h."""
from .Config import Config
from .TransD import TransD
from .TransE import TransE
from .TransH import TransH
from .TransR import TransR
from ._native import OkbError

__all__ = ["Config", "TransE", "TransH", "TransR", "TransD", "OkbError"]
