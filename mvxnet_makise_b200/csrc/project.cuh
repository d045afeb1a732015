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

// The same projection in fp64: what the reference's NUMPY branches evaluate when the calibration matrices are the float64
// arrays `readCalib` returns (Load.py:24-41: `np.zeros((4,4))`, `np.concatenate([fp32 block, [[0,0,0,1]]])` -> float64):
// cropToSight at Load.py:73 and lidar2Img of the pasted ground-truth sets at train.py:36-39. fp32 coordinates promoted to
// double, c64 = [R0@Tr (the fp64 4x4 product, formed on the host by numpy like Calib.py:65) | P2]. The dot products are
// sequential FMAs; a BLAS dgemm may round the last bit differently, which survives only where a decision or the final
// fp64 -> fp32 rounding (`torch.Tensor(voxel)`, train.py:125) sits within 1 ulp(fp64) of a tie.
__device__ __forceinline__ void project_point_z_f64(const double *__restrict__ c64, float x, float y, float z, double &u,
                                                    double &v, double &cam_z) {
    const double xd = x, yd = y, zd = z;
    double cam[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = __dmul_rn(c64[i * 4 + 0], xd);
        a = __fma_rn(c64[i * 4 + 1], yd, a);
        a = __fma_rn(c64[i * 4 + 2], zd, a);
        a = __fma_rn(c64[i * 4 + 3], 1.0, a);
        cam[i] = a;
    }
    double img[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double a = __dmul_rn(c64[16 + i * 4 + 0], cam[0]);
        a = __fma_rn(c64[16 + i * 4 + 1], cam[1], a);
        a = __fma_rn(c64[16 + i * 4 + 2], cam[2], a);
        a = __fma_rn(c64[16 + i * 4 + 3], cam[3], a);
        img[i] = a;
    }
    u = __ddiv_rn(img[0], img[2]);
    v = __ddiv_rn(img[1], img[2]);
    cam_z = cam[2];
}

}  // namespace
}  // namespace mvx
