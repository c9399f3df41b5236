// cl2_cdbscan.cu -- cDBSCAN2 (cLoops2/cDBSCAN2.py) placeholder until the kernels land.
#include "cl2_points.cuh"
extern "C" int cl2_cdbscan2(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                            int32_t *n_clusters) {
    (void)pts; (void)eps; (void)minPts; (void)labels_out; (void)n_clusters;
    return cl2_fail(ctx, CL2_ERR_ARG, "cl2_cdbscan2: not implemented in this build");
}
