// §8f rank 1 — the step immediately BEFORE the path: `crop` (range filter, modules/data/Preprocessing.py:12-17) and
// `cropToSight` (in front of the camera and inside the image, Preprocessing.py:26-55; callers modules/data/Load.py:59,73 and
// cropdata.py:32-65) on the GPU, with an order-preserving stream compaction, so raw 4-float KITTI sweeps can go
// straight to the device. Decisions use the reference's own arithmetic: the range test compares the fp32 coordinates
// promoted to fp64 against the fp64 bounds (numpy promotion of `low <= roi`), the sight test uses the same
// sequential-FMA projection as lidar2Img (project.cuh) and compares the fp32 pixel coordinates, promoted to fp64,
// against `imsize - 1e-3` (the reference's fudge against numpy / torch disagreement at the image border).
#include "project.cuh"
#include "voxelize.cuh"

namespace mvx {

namespace {

constexpr int kCropBlock = 1024;

struct CropParams {
    FrameOffsets fo;
    int B;
    const float *points;
    int stride;
    int use_range, use_sight;
    double lo[3], hi[3];
    const float *calib32;      // [B][32]
    const double *calib64;     // [B][32] fp64 calibrations (readCalib's float64 matrices): the reference's numpy branch, or NULL
    double lim_w, lim_h;       // imsize - 1e-3 (fp64)
    unsigned char *flag;       // [sum P]
    int *blocksum;             // [B][nblk] kept points per 1024-point block, then (in place) their exclusive scan
    int nblk;
    float *out;                // same offsets as the input
    int *out_counts;           // [B]
};

__device__ __forceinline__ bool keep_point(const CropParams &p, const float *c32, const double *c64, const float *q) {
    const float x = q[0], y = q[1], z = q[2];
    if (p.use_range) {
        const double xd = x, yd = y, zd = z;
        if (!(p.lo[0] <= xd && xd < p.hi[0] && p.lo[1] <= yd && yd < p.hi[1] && p.lo[2] <= zd && zd < p.hi[2])) return false;
    }
    if (p.use_sight && c64) {                               // float64 calibration: the whole test in fp64 (Load.py:73)
        double u, v, cz;
        project_point_z_f64(c64, x, y, z, u, v, cz);
        if (!(cz > 0.0)) return false;
        if (!(u >= 0.0 && v >= 0.0 && u < p.lim_w && v < p.lim_h)) return false;
    } else if (p.use_sight) {
        float u, v, cz;
        project_point_z(c32, x, y, z, u, v, cz);
        if (!(cz > 0.f)) return false;                      // behind the camera (Preprocessing.py:46-48)
        if (!(u >= 0.f && v >= 0.f && (double)u < p.lim_w && (double)v < p.lim_h)) return false;   // :52
    }
    return true;
}

__global__ void __launch_bounds__(kCropBlock) crop_flag_kernel(CropParams p) {
    __shared__ float c32[32];
    __shared__ double c64[32];
    const int f = blockIdx.y;
    if (p.use_sight && threadIdx.x < 32) {
        if (p.calib64) c64[threadIdx.x] = p.calib64[f * 32 + threadIdx.x];
        else c32[threadIdx.x] = p.calib32[f * 32 + threadIdx.x];
    }
    __syncthreads();
    const int n = p.fo.off[f + 1] - p.fo.off[f];
    const int i = blockIdx.x * kCropBlock + threadIdx.x;
    if (blockIdx.x * kCropBlock >= n) return;
    bool k = false;
    if (i < n) {
        k = keep_point(p, c32, p.calib64 ? c64 : nullptr, p.points + (size_t)(p.fo.off[f] + i) * p.stride);
        p.flag[p.fo.off[f] + i] = k ? 1 : 0;
    }
    const int cnt = __syncthreads_count(k ? 1 : 0);
    if (threadIdx.x == 0) p.blocksum[f * p.nblk + blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) crop_scan_kernel(CropParams p) {   // one CTA per frame
    const int f = blockIdx.x;
    const int n = p.fo.off[f + 1] - p.fo.off[f];
    const int nb = (n + kCropBlock - 1) / kCropBlock;
    int *bs = p.blocksum + f * p.nblk;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < nb ? bs[b] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (b < nb) bs[b] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.out_counts[f] = carry;
}

__global__ void __launch_bounds__(kCropBlock) crop_emit_kernel(CropParams p) {
    const int f = blockIdx.y;
    const int n = p.fo.off[f + 1] - p.fo.off[f];
    if (blockIdx.x * kCropBlock >= n) return;
    const int i = blockIdx.x * kCropBlock + threadIdx.x;
    const int k = (i < n && p.flag[p.fo.off[f] + i]) ? 1 : 0;
    int total;
    const int ex = block_exclusive_scan(k, &total);
    if (k) {
        const float *src = p.points + (size_t)(p.fo.off[f] + i) * p.stride;
        float *dst = p.out + (size_t)(p.fo.off[f] + p.blocksum[f * p.nblk + blockIdx.x] + ex) * p.stride;
        for (int c = 0; c < p.stride; ++c) dst[c] = src[c];
    }
}

}  // namespace
}  // namespace mvx

extern "C" int mvx_crop_workspace_bytes(int32_t B, int64_t total_points, int64_t max_points, size_t *bytes) {
    MVX_REQUIRE(bytes && B >= 1 && B <= mvx::kMaxFrames && total_points >= 0 && max_points >= 0, MVX_EINVAL, "bad crop extents");
    const size_t nblk = (size_t)((max_points + mvx::kCropBlock - 1) / mvx::kCropBlock) + 1;
    *bytes = ((size_t)total_points + 255) / 256 * 256 + (size_t)B * nblk * sizeof(int);
    return MVX_OK;
}

static int crop_points_impl(const float *points, int32_t point_stride, int32_t B, const int32_t *pt_off_host, const double *range6,
                            const float *calib32, const double *calib64, double imsize_w, double imsize_h, float *out_points,
                            int32_t *out_counts, void *workspace, size_t workspace_bytes, void *stream) {
    MVX_REQUIRE(pt_off_host && out_counts && B >= 1 && B <= mvx::kMaxFrames && point_stride >= 3, MVX_EINVAL, "bad crop argument");
    MVX_REQUIRE(range6 || calib32 || calib64, MVX_EINVAL, "crop: give a range, a calibration, or both");
    mvx::CropParams p{};
    long long maxp = 0;
    for (int f = 0; f <= B; ++f) p.fo.off[f] = pt_off_host[f];
    for (int f = 0; f < B; ++f) {
        MVX_REQUIRE(pt_off_host[f + 1] >= pt_off_host[f], MVX_EINVAL, "point offsets must be non-decreasing");
        maxp = std::max<long long>(maxp, pt_off_host[f + 1] - pt_off_host[f]);
    }
    const long long total = pt_off_host[B] - pt_off_host[0];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (total == 0) {
        MVX_CUDA_CHECK(cudaMemsetAsync(out_counts, 0, (size_t)B * sizeof(int), st));
        return MVX_OK;
    }
    MVX_REQUIRE(points && out_points && workspace, MVX_EINVAL, "null pointer");
    size_t need = 0;
    mvx_crop_workspace_bytes(B, pt_off_host[B], maxp, &need);
    MVX_REQUIRE(workspace_bytes >= need, MVX_ESPACE, "crop workspace too small");
    p.B = B, p.points = points, p.stride = point_stride, p.out = out_points, p.out_counts = out_counts;
    p.use_range = range6 != nullptr, p.use_sight = calib32 != nullptr || calib64 != nullptr;
    if (range6)
        for (int d = 0; d < 3; ++d) p.lo[d] = range6[d], p.hi[d] = range6[3 + d];
    p.calib32 = calib32, p.calib64 = calib64;
    p.lim_w = imsize_w - 1e-3, p.lim_h = imsize_h - 1e-3;   // Preprocessing.py:37
    p.nblk = (int)((maxp + mvx::kCropBlock - 1) / mvx::kCropBlock) + 1;
    p.flag = static_cast<unsigned char *>(workspace);
    p.blocksum = reinterpret_cast<int *>(static_cast<char *>(workspace) + ((size_t)pt_off_host[B] + 255) / 256 * 256);
    const dim3 grid((unsigned)((maxp + mvx::kCropBlock - 1) / mvx::kCropBlock), B);
    mvx::crop_flag_kernel<<<grid, mvx::kCropBlock, 0, st>>>(p);
    MVX_LAUNCH_CHECK();
    mvx::crop_scan_kernel<<<B, 1024, 0, st>>>(p);
    MVX_LAUNCH_CHECK();
    mvx::crop_emit_kernel<<<grid, mvx::kCropBlock, 0, st>>>(p);
    MVX_LAUNCH_CHECK();
    return MVX_OK;
}

extern "C" int mvx_crop_points(const float *points, int32_t point_stride, int32_t B, const int32_t *pt_off_host, const double *range6,
                               const float *calib32, double imsize_w, double imsize_h, float *out_points, int32_t *out_counts,
                               void *workspace, size_t workspace_bytes, void *stream) {
    return crop_points_impl(points, point_stride, B, pt_off_host, range6, calib32, nullptr, imsize_w, imsize_h, out_points, out_counts,
                            workspace, workspace_bytes, stream);
}

extern "C" int mvx_crop_points_f64(const float *points, int32_t point_stride, int32_t B, const int32_t *pt_off_host,
                                   const double *range6, const double *calib64, double imsize_w, double imsize_h, float *out_points,
                                   int32_t *out_counts, void *workspace, size_t workspace_bytes, void *stream) {
    return crop_points_impl(points, point_stride, B, pt_off_host, range6, nullptr, calib64, imsize_w, imsize_h, out_points, out_counts,
                            workspace, workspace_bytes, stream);
}
