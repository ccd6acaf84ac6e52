"""rs-bann_b200: B200-native HMC/Gibbs hot path of rs-bann behind a C ABI (include/bann.h).

Importable as `rs_bann_b200` through the shim at the repository root.  The CUDA shared library
must be present (`__graft_entry__.build()`); there is no CPU fallback."""
from ._lib import (ACT_NAMES, HMC_ACCEPTED, HMC_REJECTED, HMC_REJECTED_EARLY, LIB_PATH, MODEL_NAMES, PROTOTYPES,
                   STEP_NAMES, BannError, lib)
from .api import Context, Genotypes, HMCStepResult, MCMCCfg, Net, cuda_available, pinned_empty, stats_from_counts


def launch_count(reset: bool = False) -> int:
    return int(lib.bann_launch_count(int(reset)))
from .dist import connect_net, connect_ranks, global_col_stats, row_shard, shard_payload  # noqa: E402
from . import architectures, files  # noqa: E402,F401  (host-side file formats and net construction; cli.py is the rs-bann command surface)
