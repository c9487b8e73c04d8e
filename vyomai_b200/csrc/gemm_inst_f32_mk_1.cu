// explicit instantiations of the vy_gemm kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"
namespace vy {
template int launch_gemm<float, 256, true, false, false>(const VyGemm*, const GemmDev&);
}  // namespace vy
