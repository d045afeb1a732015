// Calibrated projection of one LiDAR point, shared by lidar2Img, the fused path and the crop kernels.
#pragma once
#include "common.cuh"

namespace mvx {
namespace {

// ---- projection: (R0@Tr) @ [x y z 1], then P2 @ ., then divide (Calib.py:65-70) -----------------------
// Accumulation order = sequential FMA over k (what torch's CPU sgemm does for a 4x4 operand; pinned by
// tests/test_gpu_parity.py on the reference-generated golden projections).
__device__ __forceinline__ void project_point_z(const float *__restrict__ c32, float x, float y, float z, float &u,
                                                float &v, float &cam_z) {
    float cam[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = __fmul_rn(c32[i * 4 + 0], x);
        a = __fmaf_rn(c32[i * 4 + 1], y, a);
        a = __fmaf_rn(c32[i * 4 + 2], z, a);
        a = __fmaf_rn(c32[i * 4 + 3], 1.0f, a);
        cam[i] = a;
    }
    float img[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float a = __fmul_rn(c32[16 + i * 4 + 0], cam[0]);
        a = __fmaf_rn(c32[16 + i * 4 + 1], cam[1], a);
        a = __fmaf_rn(c32[16 + i * 4 + 2], cam[2], a);
        a = __fmaf_rn(c32[16 + i * 4 + 3], cam[3], a);
        img[i] = a;
    }
    u = __fdiv_rn(img[0], img[2]);
    v = __fdiv_rn(img[1], img[2]);
    cam_z = cam[2];
}
__device__ __forceinline__ void project_point(const float *__restrict__ c32, float x, float y, float z, float &u, float &v) {
    float cz;
    project_point_z(c32, x, y, z, u, v, cz);
}


}  // namespace
}  // namespace mvx
