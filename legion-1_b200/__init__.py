"""legion_b200: B200-native (sm_100a) implementation of Legion's GPU-initiated mini-batch
pipeline -- k-hop sampling, presampling hotness, tiered feature extraction -- behind a
C-ABI (include/legion_b200.h).  Package directory: legion-1_b200/ (import as legion_b200).
"""
from . import _lib, synth  # noqa: F401
from ._lib import (DevArray, MappedHostArray, LegionError, RNG_MINSTD, RNG_PHILOX,  # noqa: F401
                   MODE_TRAIN, MODE_VALID, MODE_TEST, build, lib)
from .runner import (Runner, hot_order, place, fill_feature_shard, fill_topo_shard, cost_model,  # noqa: F401
                     coordinate, place_hybrid, fill_feature_shard_hybrid, plan_hybrid, place_compact, fill_feature_shard_compact, Stream, shared_alloc, shared_import, shared_free)
