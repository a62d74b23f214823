// tcgen05 / TMA / TMEM path (VBG_PREC_TF32) -- placeholder until the kernel lands.
#include "vbg_common.cuh"
namespace vbg {
bool tc_available() { return false; }
int gemm_tc(const float*, int, const float*, int, int, const float*, int, float*, int, int, int, int,
            const vbg_epilogue_t*, cudaStream_t) { return VBG_EUNSUPPORTED; }
int conv_tc(const float*, int, int, int, int, const float*, int, int, int, int, int, float*,
            const vbg_epilogue_t*, cudaStream_t) { return VBG_EUNSUPPORTED; }
}  // namespace vbg
