// Density-grid maintenance helpers (NeRFRenderer.update_extra_state,
// /root/reference/reconstruction/nerf/renderer.py:448-542) -- sm_100a.
#include "common.cuh"

namespace tnl {

__device__ __forceinline__ uint32_t compact3g(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// renderer.py:474-483 / 507-513:  xyz = (2*coord/(H-1) - 1) * (bound_c - hgs) + (2*noise - 1) * hgs
// with torch's CUDA scalar rules (tensor / python-scalar == tensor * fp32(1/scalar)).
__global__ void k_grid_cell_positions(const int32_t* __restrict__ indices, uint32_t n, float rHm1, float scale, float hgs,
                                      const float* __restrict__ noise, float* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t idx = (uint32_t)indices[i];
    const uint32_t c[3] = {compact3g(idx), compact3g(idx >> 1), compact3g(idx >> 2)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = __fmul_rn(__fmul_rn(2.0f, (float)c[a]), rHm1);
        v = __fsub_rn(v, 1.0f);
        v = __fmul_rn(v, scale);
        const float j = __fmul_rn(__fsub_rn(__fmul_rn(noise[3 * (size_t)i + a], 2.0f), 1.0f), hgs);
        xyz[3 * (size_t)i + a] = __fadd_rn(v, j);
    }
}

// renderer.py:526-527: grid = max(grid*decay, tmp) where grid >= 0 and tmp >= 0
__global__ void k_grid_ema(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n, float decay) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = grid[i], t = tmp[i];
    if (g >= 0.f && t >= 0.f) grid[i] = fmaxf(__fmul_rn(g, decay), t);
}

}  // namespace tnl

using namespace tnl;

extern "C" {

int tnl_grid_cell_positions(const int32_t* indices, uint32_t n, uint32_t H, float bound_c, const float* noise, float* xyz,
                            tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(indices && noise && xyz && H >= 2, "bad argument");
    const float hgs = (float)((double)bound_c / (double)H);
    const float scale = (float)((double)bound_c - (double)bound_c / (double)H);
    const float rHm1 = 1.0f / (float)(H - 1);
    k_grid_cell_positions<<<ceil_div(n, 256u), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(indices, n, rHm1, scale, hgs,
                                                                                                noise, xyz);
    return finish_launch("grid_cell_positions");
}

int tnl_grid_ema_update(float* grid, const float* tmp_grid, uint32_t n, float decay, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(grid && tmp_grid, "null pointer");
    k_grid_ema<<<ceil_div(n, 256u), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grid, tmp_grid, n, decay);
    return finish_launch("grid_ema_update");
}

}  // extern "C"
