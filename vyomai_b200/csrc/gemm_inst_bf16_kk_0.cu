// explicit instantiations of the vy_gemm kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"
namespace vy {
template int launch_gemm<__nv_bfloat16, 32, false, false, false>(const VyGemm*, const GemmDev&);
template int launch_gemm<__nv_bfloat16, 64, false, false, false>(const VyGemm*, const GemmDev&);
template int launch_gemm<__nv_bfloat16, 128, false, false, false>(const VyGemm*, const GemmDev&);
}  // namespace vy
