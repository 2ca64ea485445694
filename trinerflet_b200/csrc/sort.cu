// Spatial binning of the step's sample points (counting sort by 3-D cell) -- sm_100a.
//
// Why: the reference evaluates samples in marching order, i.e. ray by ray for randomly drawn pixels, so
// consecutive warps touch unrelated plane texels: almost every 128-byte texel fetched by the gather and every
// read-modify-write of the gradient scatter misses the 126 MB L2 (ncu: k_sample_bwd moves 9 GB through DRAM for
// 0.3 GB of distinct gradient texels).  Visiting the same points in the order of a coarse 3-D grid makes the
// projections on all three planes coherent at once, so texels are fetched from DRAM about once per step.
// Per-point results are unchanged (every point is still evaluated exactly once; outputs are written back to the
// point's original row), only the order of floating-point accumulation into gradients differs.
//
// perm[i] = original row of the i-th point in cell order; cells = G^3 grid over [-bound, bound]^3 in Morton order,
// rows >= *n_valid (padding) are placed last in their original order.
#include "common.cuh"

namespace tnl {

__device__ __forceinline__ uint32_t spread3s(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint32_t cell_key(const float* __restrict__ xyz, uint32_t m, float inv_bound, int G) {
    uint32_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float u = fminf(fmaxf((__ldg(xyz + 3 * (size_t)m + a) * inv_bound + 1.0f) * 0.5f, 0.0f), 1.0f);
        c[a] = min((uint32_t)(u * (float)G), (uint32_t)(G - 1));
    }
    return spread3s(c[0]) | (spread3s(c[1]) << 1) | (spread3s(c[2]) << 2);
}

// pass 1: histogram (bins = G^3 + 1; the last bin collects padding rows)
__global__ void k_cell_hist(const float* __restrict__ xyz, uint32_t M, const int32_t* __restrict__ n_valid, float inv_bound,
                            int G, uint32_t* __restrict__ hist, uint32_t* __restrict__ keys) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const uint32_t nv = n_valid ? (uint32_t)max(*n_valid, 0) : M;
    const uint32_t key = m < nv ? cell_key(xyz, m, inv_bound, G) : (uint32_t)(G * G * G);
    keys[m] = key;
    atomicAdd(hist + key, 1u);
}

// pass 2: exclusive scan of the histogram in place, three small kernels (4096 bins per block):
//   block sums -> scan of the block sums (one block) -> per-block exclusive scan + block offset
constexpr int kBinsPerBlock = 4096;  // 1024 threads x 4 bins

__device__ __forceinline__ uint32_t block_scan4(uint32_t (&v)[4], uint32_t* sm /*[33]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tsum = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    total = sm[32];
    return sm[warp] + incl - tsum;  // exclusive prefix of this thread's first bin within the block
}

__global__ void __launch_bounds__(1024) k_cell_block_sums(const uint32_t* __restrict__ hist, uint32_t nbins, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sm[33];
    const uint32_t i0 = blockIdx.x * kBinsPerBlock + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i0 + j < nbins) ? hist[i0 + j] : 0u;
    uint32_t total;
    block_scan4(v, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_cell_scan_sums(uint32_t* __restrict__ sums, uint32_t nblocks) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += kBinsPerBlock) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < nblocks) ? sums[i0 + j] : 0u;
        uint32_t total;
        uint32_t ex = carry_s + block_scan4(v, sm, total);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < nblocks) sums[i0 + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_cell_scan_apply(uint32_t* __restrict__ hist, uint32_t nbins, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t sm[33];
    const uint32_t i0 = blockIdx.x * kBinsPerBlock + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i0 + j < nbins) ? hist[i0 + j] : 0u;
    uint32_t total;
    uint32_t ex = sums[blockIdx.x] + block_scan4(v, sm, total);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < nbins) hist[i0 + j] = ex;
        ex += v[j];
    }
}

// pass 3: scatter row ids into their bins
__global__ void k_cell_scatter(const uint32_t* __restrict__ keys, uint32_t M, uint32_t* __restrict__ cursor,
                               int32_t* __restrict__ perm) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const uint32_t slot = atomicAdd(cursor + keys[m], 1u);
    perm[slot] = (int32_t)m;
}

}  // namespace tnl

using namespace tnl;

extern "C" {

size_t tnl_cell_sort_workspace(uint32_t M, uint32_t G) {
    const size_t nbins = (size_t)G * G * G + 1;
    return sizeof(uint32_t) * (nbins + M + (nbins + kBinsPerBlock - 1) / kBinsPerBlock + 1);
}

int tnl_cell_sort(const float* xyz, uint32_t M, const int32_t* n_valid, float inv_bound, uint32_t G, int32_t* perm,
                  void* workspace, size_t workspace_bytes, tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(xyz && perm, "null pointer");
    TNL_ARG_CHECK(G >= 2 && G <= 256 && (G & (G - 1)) == 0, "G must be a power of two in [2, 256] (Morton keys index the bins)");
    if (workspace == nullptr || workspace_bytes < tnl_cell_sort_workspace(M, G)) {
        set_error("cell_sort: workspace too small");
        return TNL_ERR_WORKSPACE;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t nbins = G * G * G + 1;
    uint32_t* hist = static_cast<uint32_t*>(workspace);
    uint32_t* keys = hist + nbins;
    uint32_t* sums = keys + M;
    const uint32_t nblocks = ceil_div(nbins, (uint32_t)kBinsPerBlock);
    cudaMemsetAsync(hist, 0, sizeof(uint32_t) * nbins, s);
    k_cell_hist<<<ceil_div(M, 256u), 256, 0, s>>>(xyz, M, n_valid, inv_bound, (int)G, hist, keys);
    k_cell_block_sums<<<nblocks, 1024, 0, s>>>(hist, nbins, sums);
    k_cell_scan_sums<<<1, 1024, 0, s>>>(sums, nblocks);
    k_cell_scan_apply<<<nblocks, 1024, 0, s>>>(hist, nbins, sums);
    k_cell_scatter<<<ceil_div(M, 256u), 256, 0, s>>>(keys, M, hist, perm);
    return finish_launch("cell_sort");
}

}  // extern "C"
