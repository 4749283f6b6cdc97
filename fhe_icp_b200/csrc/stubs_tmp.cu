#include "kernels.h"
namespace fhe {
cudaError_t launch_ksk_gen(const fhe_b200_pbs_params&, const uint8_t*, const uint8_t*, uint64_t, uint64_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_bsk_gen(const fhe_b200_pbs_params&, const uint8_t*, const uint8_t*, uint64_t, uint64_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_keyswitch(const fhe_b200_pbs_params&, const uint64_t*, const uint64_t*, int64_t, uint64_t*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_bsk_to_fourier(const fhe_b200_pbs_params&, const uint64_t*, double*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_pbs(const fhe_b200_pbs_params&, const double*, const uint64_t*, int64_t, const uint64_t*, const int32_t*, uint64_t*, int, cudaStream_t) { return cudaErrorNotSupported; }
bool pbs_params_supported(const fhe_b200_pbs_params&, const char** why) { *why = "stub"; return false; }
}
