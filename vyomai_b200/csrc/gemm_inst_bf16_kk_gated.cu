// explicit instantiations of the vy_gemm kernels whose epilogue is the staged SwiGLU one (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"
namespace vy {
template int launch_gemm<__nv_bfloat16, 128, false, false, false, true>(const VyGemm*, const GemmDev&);
template int launch_gemm<__nv_bfloat16, 256, false, false, false, true>(const VyGemm*, const GemmDev&);
template int launch_gemm<__nv_bfloat16, 128, false, false, true, true>(const VyGemm*, const GemmDev&);
template int launch_gemm<__nv_bfloat16, 256, false, false, true, true>(const VyGemm*, const GemmDev&);
}  // namespace vy
