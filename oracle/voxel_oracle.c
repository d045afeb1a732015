/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or called by the product path
 * (mvxnet_makise_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the CPU baseline.
 *
 * Plain-C, single-threaded restatement of the reference voxelizer:
 *   - cell index math   : modules/data/Preprocessing.py:67-69 (group_) and :86-90 (group):
 *                         idx = ((pts - low) / size).astype(int32), computed in fp64 because
 *                         `low`/`size` are float64 arrays (SURVEY.md trap 1), C truncation.
 *   - grouping          : cpp/voxelutil.cpp:325-360 (`group`, exported as `_group`) and the numba
 *                         loop Preprocessing.py:94-104: one pass in input order, a voxel is created
 *                         at the first occurrence of its (ix,iy,iz) key, a point is kept iff its
 *                         voxel currently holds < T points.
 *   - (V,T,9) assembly  : Preprocessing.py:103 (columns x,y,z,_,_,_,r,row,col) and :112-115
 *                         (centroid = sum over the T slots / cnt in fp64; cols 3..5 = xyz - centroid
 *                         for ALL T slots, pads included).
 * Parity pinning: the reference holds no golden vectors; this file is pinned
 * against the reference's own code executed in the build container (tests/golden/make_golden.py
 * -> tests/golden/*.npz, and oracle/_ref/voxelutil*.so built from /root/reference in place).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t x, y, z; int64_t vid; } slot_t;

static uint64_t mix(int32_t x, int32_t y, int32_t z) {
    uint64_t h = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)y + 0x7F4A7C159E3779B9ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)z + 0x165667B19E3779F9ull) * 0x27D4EB2F165667C5ull;
    h ^= h >> 29;
    return h;
}

/* idx[i][d] = (int32) trunc(((double)pcd[i][d] - low[d]) / size[d]);  pcd row stride = `stride` floats */
void oracle_cell_index(const float *pcd, int64_t P, int64_t stride, const double *low, const double *size,
                       int32_t *idx) {
    for (int64_t i = 0; i < P; i++)
        for (int d = 0; d < 3; d++) {
            double q = ((double)pcd[i * stride + d] - low[d]) / size[d];
            idx[i * 3 + d] = (int32_t)q;
        }
}

/*
 * Grouping pass. Outputs (caller-allocated, worst case V = P):
 *   vid_of_point[i]  : voxel id of point i, or -1 if dropped by the T cap
 *   slot_of_point[i] : slot inside the voxel (0..T-1), or -1
 *   coords[v][3]     : (ix,iy,iz) of voxel v in first-occurrence order
 *   cnt[v]           : min(#points, T)
 * Returns V.
 */
int64_t oracle_group_assign(const int32_t *idx, int64_t P, int32_t T, int64_t *vid_of_point,
                            int32_t *slot_of_point, int64_t *coords, int64_t *cnt) {
    uint64_t cap = 16;
    while (cap < (uint64_t)P * 2 + 2) cap <<= 1;
    slot_t *tab = (slot_t *)malloc(cap * sizeof(slot_t));
    for (uint64_t k = 0; k < cap; k++) tab[k].vid = -1;
    int64_t V = 0;
    for (int64_t i = 0; i < P; i++) {
        int32_t x = idx[i * 3], y = idx[i * 3 + 1], z = idx[i * 3 + 2];
        uint64_t h = mix(x, y, z) & (cap - 1);
        while (tab[h].vid >= 0 && !(tab[h].x == x && tab[h].y == y && tab[h].z == z)) h = (h + 1) & (cap - 1);
        if (tab[h].vid < 0) {                  /* first occurrence -> new voxel (voxelutil.cpp:333-338) */
            tab[h].x = x; tab[h].y = y; tab[h].z = z; tab[h].vid = V;
            coords[V * 3] = x; coords[V * 3 + 1] = y; coords[V * 3 + 2] = z;
            cnt[V] = 1;
            vid_of_point[i] = V; slot_of_point[i] = 0;
            V++;
        } else {
            int64_t v = tab[h].vid;
            if (cnt[v] < T) {                  /* voxelutil.cpp:339-341 */
                vid_of_point[i] = v; slot_of_point[i] = (int32_t)cnt[v];
                cnt[v]++;
            } else {
                vid_of_point[i] = -1; slot_of_point[i] = -1;
            }
        }
    }
    free(tab);
    return V;
}

/* `_group` output: (V,T,7) fp32 zero-initialised, cols 0-2 xyz, col 6 = pcd[:,3] (voxelutil.cpp:344-357) */
void oracle_emit_group7(const float *pcd, int64_t P, int64_t stride, int32_t T, const int64_t *vid_of_point,
                        const int32_t *slot_of_point, int64_t V, float *voxel) {
    memset(voxel, 0, (size_t)V * T * 7 * sizeof(float));
    for (int64_t i = 0; i < P; i++) {
        if (vid_of_point[i] < 0) continue;
        float *o = voxel + ((size_t)vid_of_point[i] * T + slot_of_point[i]) * 7;
        o[0] = pcd[i * stride]; o[1] = pcd[i * stride + 1]; o[2] = pcd[i * stride + 2];
        o[6] = pcd[i * stride + 3];
    }
}

/* numba `group` output: (V,T,9) fp64 (Preprocessing.py:94-115); pcd is (P,6) = [x,y,z,r,row,col] */
void oracle_emit_group9(const float *pcd, int64_t P, int64_t stride, int32_t T, const int64_t *vid_of_point,
                        const int32_t *slot_of_point, int64_t V, const int64_t *cnt, double *voxel) {
    memset(voxel, 0, (size_t)V * T * 9 * sizeof(double));
    for (int64_t i = 0; i < P; i++) {
        if (vid_of_point[i] < 0) continue;
        double *o = voxel + ((size_t)vid_of_point[i] * T + slot_of_point[i]) * 9;
        const float *p = pcd + i * stride;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[6] = p[3]; o[7] = p[4]; o[8] = p[5];
    }
    for (int64_t v = 0; v < V; v++) {
        double *vx = voxel + (size_t)v * T * 9;
        double c[3];
        for (int d = 0; d < 3; d++) {
            double s = 0.0;
            for (int32_t j = 0; j < T; j++) s += vx[j * 9 + d];
            c[d] = s / (double)cnt[v];
        }
        for (int32_t j = 0; j < T; j++)
            for (int d = 0; d < 3; d++) vx[j * 9 + 3 + d] = vx[j * 9 + d] - c[d];
    }
}
