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

// renderer.py:526-527: grid = max(grid*decay, tmp) where grid >= 0 and tmp >= 0, plus the
// sum of clamp(grid, 0) over all cells (renderer.py:528: the mean that sets the occupancy threshold), accumulated
// in double: one atomicAdd per block
__global__ void __launch_bounds__(256)
k_grid_ema_sum(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n, float decay, double* __restrict__ sum) {
    __shared__ float part[8];
    float acc = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        const float t = tmp[i];
        if (g >= 0.f && t >= 0.f) { g = fmaxf(__fmul_rn(g, decay), t); grid[i] = g; }
        acc += fmaxf(g, 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)part[w];
        atomicAdd(sum, s);
    }
}

// renderer.py:528-534 without the host round trip: mean = sum / n, thresh = min(mean, cap), bitfield = packbits(grid > thresh);
// one thread packs 4 bytes (32 cells); thread 0 also publishes the mean
__global__ void k_packbits_mean(const float* __restrict__ grid, uint32_t n_cells, const double* __restrict__ sum, float cap,
                                float* __restrict__ mean_out, uint8_t* __restrict__ bitfield) {
    const float mean = (float)(*sum / (double)n_cells);
    const float thresh = fminf(mean, cap);
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w == 0) *mean_out = mean;
    if (w >= n_cells / 32) return;
    const float4* g = reinterpret_cast<const float4*>(grid) + (size_t)w * 8;
    uint32_t bits = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = __ldg(g + q);
        bits |= (v.x > thresh ? 1u : 0u) << (4 * q + 0);
        bits |= (v.y > thresh ? 1u : 0u) << (4 * q + 1);
        bits |= (v.z > thresh ? 1u : 0u) << (4 * q + 2);
        bits |= (v.w > thresh ? 1u : 0u) << (4 * q + 3);
    }
    reinterpret_cast<uint32_t*>(bitfield)[w] = bits;
}

// tmp[base + indices[i]] = sigma[i] * scale  (renderer.py:486-488 / 516-518: `tmp_grid[cas, indices] = sigmas`; duplicate
// indices: one of the writers wins, as in the reference's index_put_)
__global__ void k_grid_scatter(const int32_t* __restrict__ indices, const float* __restrict__ sigma, uint32_t n, float scale,
                               float* __restrict__ tmp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tmp[indices[i]] = __fmul_rn(sigma[i], scale);
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

int tnl_grid_ema_update_sum(float* grid, const float* tmp_grid, uint32_t n, float decay, double* sum, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(grid && tmp_grid && sum, "null pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaMemsetAsync(sum, 0, sizeof(double), s);
    k_grid_ema_sum<<<min(ceil_div(n, 256u), (uint32_t)kNumSM * 8u), 256, 0, s>>>(grid, tmp_grid, n, decay, sum);
    return finish_launch("grid_ema_update_sum");
}

int tnl_packbits_mean(const float* grid, uint32_t n_cells, const double* sum, float thresh_cap, float* mean_out, uint8_t* bitfield,
                      tnl_stream_t stream) {
    if (n_cells == 0) return 0;
    TNL_ARG_CHECK(grid && sum && mean_out && bitfield, "null pointer");
    TNL_ARG_CHECK(n_cells % 32 == 0, "the cell count must be a multiple of 32");
    TNL_ARG_CHECK((reinterpret_cast<uintptr_t>(grid) & 15) == 0 && (reinterpret_cast<uintptr_t>(bitfield) & 3) == 0,
                  "grid must be 16-byte aligned, bitfield 4-byte aligned");
    k_packbits_mean<<<ceil_div(n_cells / 32, 256u), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grid, n_cells, sum, thresh_cap, mean_out,
                                                                                                      bitfield);
    return finish_launch("packbits_mean");
}

int tnl_grid_scatter(const int32_t* indices, const float* sigma, uint32_t n, float scale, float* tmp_grid_cascade, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(indices && sigma && tmp_grid_cascade, "null pointer");
    k_grid_scatter<<<ceil_div(n, 256u), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(indices, sigma, n, scale, tmp_grid_cascade);
    return finish_launch("grid_scatter");
}

}  // extern "C"
