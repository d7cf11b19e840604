// Cluster / shared-memory-resident MMTM kernels (single pass over HBM).  Placeholder until the
// streaming path is parity-green on hardware: reports "unsupported" so capi.cu falls back.
#include "common.cuh"
#include "kernels.h"

namespace gml {

bool fused_supported(int, int, int, int, int, int, int) { return false; }
int launch_fused_fwd(const FusedFwdArgs&, cudaStream_t) { return GML_E_UNSUPPORTED; }
int launch_fused_bwd(const FusedBwdArgs&, cudaStream_t) { return GML_E_UNSUPPORTED; }

}  // namespace gml
