"""toycluster_b200 -- B200 (sm_100a) replacement for Toycluster's SPH-density + WVT-relaxation
hot path (peano.c / sort.c / tree.c / sph.c / wvt_relax.c), behind a C ABI (include/toygpu.h).

``HotPath`` mirrors the reference's three operators; ``workloads`` builds synthetic
``cluster.par`` inputs.  There is deliberately no CPU implementation in this package.
"""
from .api import (DESNNGB, NGBMAX, NUMITER, WVT_SEQUENTIAL, EXACT_NEIGHBOURS, FAST, EXPORTS, HotPath, ToyGpuError, build,
                  load, LIB_PATH)
from . import workloads

__all__ = ["HotPath", "ToyGpuError", "build", "load", "workloads", "DESNNGB", "NGBMAX",
           "NUMITER", "WVT_SEQUENTIAL", "EXACT_NEIGHBOURS", "FAST", "EXPORTS", "LIB_PATH"]
