// explicit instantiation of a CTA-pair (cta_group::2) vy_gemm kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"
namespace vy {
template int launch_gemm<__nv_bfloat16, 128, true, true, true>(const VyGemm*, const GemmDev&);
}  // namespace vy
