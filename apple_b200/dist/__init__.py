"""Sharding of one mesh across the GPUs of a node (new functionality: the reference is single-GPU,
SURVEY.md section 2.2 / 8e).

Tets are split into contiguous chunks of the Morton order, one chunk per rank (one process per GPU).
A rank keeps a local copy of every vertex its tets touch.  Element operators produce PARTIAL nodal
sums on shared vertices; ``HaloExchange.sum_`` adds the partials of all sharers (NCCL all-to-all over
NVLink) in a fixed rank order, so every replica of a shared vertex holds bit-identical values and
the element-wise PNCG vector updates can simply be replicated on ghosts.  Scalars (energy, p.Hp,
the PNCG dot products) are reduced over ``counted`` (owned) entries and all-reduced.
"""

from ._halo import HaloExchange, PeerHaloExchange
from ._partition import Shard, partition_mesh, slab_shard, slab_shard_device
from ._pncg import ShardedPNCG
from ._model import ShardedOperators

__all__ = ["HaloExchange", "PeerHaloExchange", "Shard", "ShardedOperators", "ShardedPNCG", "partition_mesh", "slab_shard",
           "slab_shard_device"]
