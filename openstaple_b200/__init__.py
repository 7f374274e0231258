"""openstaple_b200 -- B200-native (CUDA sm_100a) drop-in for OpenStaPLE's staggered fermion-solver
hot path.  The product is the C-ABI library ``libstaple_b200.so`` (include/staple_b200.h); this
package is the thin host-side mirror of the reference's operator/solver interface used by the
tests and benches.  PyTorch only supplies device memory, streams and torch.distributed."""
from .lib import load_library, library_path  # noqa: F401
from .api import (Lattice, FermParam, RationalApprox, InverterPackage, HostArray, geometry_plan, tamat_fields,  # noqa: F401
                  INVERTER_SUCCESS, INVERTER_FAILURE, CONVERGENCE_CRITICAL, CONVERGENCE_NONCRITICAL)
