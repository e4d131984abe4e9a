// Instantiations of the fused SRA / one-shot kernels for T = float (one TU per dtype so the
// three compile in parallel).
#include "sra_kernels.cuh"

namespace cgx {
cudaError_t launch_sra_f32(const SraParams& p, cudaStream_t stream) { return launch_sra_t<float>(p, stream); }
int sra_resident_per_sm_f32() { return sra_resident_per_sm_t<float>(); }
}  // namespace cgx
