// blur_fused.cu -- fused X/Y/Z separable Gaussian (fast path).  Placeholder until
// the generic path is parity-green on the GPU.
#include "common.cuh"

bool s3d_blur_fused_eligible(int, int, int, int, const TapSet &, const float[3]) { return false; }

int s3d_blur_fused(s3d_engine *e, const float *, float *, int, int, int, const TapSet &,
                   const float[3])
{
    return s3d_fail(e, "fused blur not built", cudaSuccess, __FILE__, __LINE__);
}
