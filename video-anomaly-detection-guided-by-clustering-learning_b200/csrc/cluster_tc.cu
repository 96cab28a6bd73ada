// cluster_tc.cu — tcgen05 / TMEM / TMA fused cluster forward (placeholder until
// the kernel lands: reports UNSUPPORTED so VADC_IMPL_AUTO takes the SIMT path).
#include "common.cuh"
#include "cluster.h"

extern "C" size_t vadc_cluster_tc_extra_workspace_bytes(int64_t, int, int) { return 0; }

int vadc_cluster_fwd_tc(const float*, const float*, const float*, const float*, int64_t, int, int,
                        float, float, float*, float*, float*, float*, int64_t*, float*, float*,
                        float*, void*, size_t, cudaStream_t) {
  return VADC_ERR_UNSUPPORTED;
}
