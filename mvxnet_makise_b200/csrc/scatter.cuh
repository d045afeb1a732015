// Internal interface of stage 4 (dense grid fill), shared with the fused path.
#pragma once
#include "common.cuh"

namespace mvx {
// out (B, C, G) fp32 = cell2vid[b][g] < 0 ? 0 : feat[b][vid][c]; feat frames at stride vcap*C. Needs G % 4 == 0.
int launch_grid_fill(const int *cell2vid, const float *feat, float *out, int B, long long G, int C, int vcap, cudaStream_t st);
void set_grid_mode(int m);  // 1 = bulk-store (TMA) kernel (default), 0 = per-thread streaming stores
}  // namespace mvx
