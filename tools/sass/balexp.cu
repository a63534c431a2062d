// scratch translation unit for SASS accounting: only the config-3 kernels of ntt_bal.cu (compiles in seconds)
#include "../../gpu-homomorphic-encryption_b200/csrc/ntt_bal.cu"
namespace fhe_b200 {
template __global__ void bal_a_kernel<8, 16, true, false>(const BalArgs);
template __global__ void bal_a_kernel<8, 16, true, true>(const BalArgs);
template __global__ void bal_b_kernel<8, 16, true, false>(const BalArgs);
template __global__ void bal_b_kernel<8, 16, true, true>(const BalArgs);
}
