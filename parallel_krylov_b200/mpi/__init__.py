"""Row-partitioned multi-GPU entry points (≙ /root/reference/v3/gpu/mpi): ``fn(comm, local_A, b, ...)``."""
from .cg import cg
from .mrr import mrr
from .kskipcg import kskipcg
from .kskipmrr import kskipmrr
from .adaptivekskipmrr import adaptivekskipmrr
from .cgcg import cgcg
from ._dist import DistOperator, build_halo_plan, row_offsets_from_local

__all__ = ["cg", "mrr", "kskipcg", "kskipmrr", "adaptivekskipmrr", "cgcg", "DistOperator", "build_halo_plan",
           "row_offsets_from_local"]
