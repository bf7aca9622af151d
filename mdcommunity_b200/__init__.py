"""B200-native (sm_100a) greedy-dismantling rollout for multiplex networks.

Drop-in for one hot path of KelvinRyman/MDCommunity (MultiDismantler): graph state -> aggregation + fusion ->
Q scoring + arg-max -> LMCC recompute + AUDC.  Hand-written CUDA behind a C ABI (include/mdcommunity_b200.h);
PyTorch supplies device memory and streams only.  No CPU fallback.
"""
from .graph import Graph_test, GSet, GraphBatch  # noqa: F401
from .rollout import BatchRollout, QWeights  # noqa: F401
from .env import MvcEnv  # noqa: F401
from .dismantler import MultiDismantler  # noqa: F401

__version__ = "0.1.0"
