// Instantiations of the fused SRA / one-shot kernels for T = __nv_bfloat16 (one TU per dtype so the
// three compile in parallel).
#include "sra_kernels.cuh"

namespace cgx {
cudaError_t launch_sra_bf16(const SraParams& p, cudaStream_t stream) { return launch_sra_t<__nv_bfloat16>(p, stream); }
int sra_resident_per_sm_bf16() { return sra_resident_per_sm_t<__nv_bfloat16>(); }
}  // namespace cgx
