// tcgen05 / TMEM / TMA region-word pair kernels (sm_100a).  Placeholder until the kernels land.
#include "common.cuh"
namespace dvlp {
bool tc_supported(int R, int W, int d) { return false; }
bool tc_bwd_available() { return false; }
int run_pair_tc_fwd(const void*, const void*, int, int, int, int, int, float, int, float*, char*, const Workspace&,
                    cudaStream_t) { return DVLP_ERR_BAD_SHAPE; }
int run_pair_tc_bwd(const void*, const void*, int, int, int, int, int, float, int, const float*, char*,
                    const Workspace&, cudaStream_t) { return DVLP_ERR_BAD_SHAPE; }
}  // namespace dvlp
