// Instantiations of the fused SRA / one-shot kernels for T = __half (one TU per dtype so the
// three compile in parallel).
#include "sra_kernels.cuh"

namespace cgx {
cudaError_t launch_sra_f16(const SraParams& p, cudaStream_t stream) { return launch_sra_t<__half>(p, stream); }
int sra_resident_per_sm_f16() { return sra_resident_per_sm_t<__half>(); }
}  // namespace cgx
