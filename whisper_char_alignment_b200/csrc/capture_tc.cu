// tcgen05 / TMEM / TMA cross-attention capture (north-star kernel 1) -- placeholder until
// the kernel lands; the dispatcher in cabi.cu falls through to the CUDA-core kernel.
#include "common.cuh"

namespace wca {

bool capture_tc_supported(int, int, int) { return false; }

int launch_capture_tc(const float *const *, const float *const *, int, int, int64_t, int64_t, const wca_utt_t *, int,
                      int, int, int, float, float *, unsigned, int, cudaStream_t) {
    set_error("capture_tc: not built");
    return WCA_ERR_UNSUPPORTED;
}

}  // namespace wca
