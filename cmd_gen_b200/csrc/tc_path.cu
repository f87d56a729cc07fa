// tcgen05 tensor-core path (DP_TF32 / DP_BF16 / DP_F16) — placeholder until the kernels land.
#include "common.cuh"

struct TcWeights { int unused; };

int tc_init() { return DP_OK; }
int tc_prepare_weights(dp_handle*, const float*) { return DP_OK; }
void tc_free_weights(dp_handle*) {}
int launch_linear_tc(dp_handle*, const LinearArgs&, int, cudaStream_t)
{
    dp_set_error("tensor-core precision modes are not built yet; use DP_FP32");
    return DP_ERR_INVALID;
}
int launch_edge_tc(dp_handle*, const EdgeArgs&, int, cudaStream_t)
{
    dp_set_error("tensor-core precision modes are not built yet; use DP_FP32");
    return DP_ERR_INVALID;
}
