// Dirty-tile bookkeeping for the multi-GPU gradient exchange -- sm_100a.
//
// In ray-sharded data parallelism every rank must sum the gradient of the three feature planes (P bytes, 1.6 GB for
// base-light).  A sample only touches texels under the projection of an occupied density-grid cell, so the plane
// gradient is exactly zero outside those projections (about 20-25 % of the plane area for a centred object).  All ranks
// hold the same density bitfield, so they derive the same "dirty" tile set without communicating; the exchange then
// packs those tiles into a compact buffer, all-reduces it with NCCL and scatters it back.  No reference counterpart
// (the reference is single-GPU).
#include "common.cuh"
#include <cuda_bf16.h>

namespace tnl {

__device__ __forceinline__ uint32_t compact3t(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// one thread per density-grid cell: an occupied cell of cascade `level` covers, per axis,
//   p in [(2n/H - 1) * mb, (2(n+1)/H - 1) * mb],  mb = min(2^level, bound)       (raymarching.cu:370-376 inverted)
// a sample at p reads/writes texels floor(ix), floor(ix)+1 with ix = (p/bound + 1)/2 * (R-1) (+- fp16 rounding of p/bound);
// `margin` texels of slack cover the rounding and the bilinear footprint.
__global__ void k_mark_dirty_tiles(const uint8_t* __restrict__ bitfield, uint32_t cascade, uint32_t H, float bound, int R, int T,
                                   int margin, uint8_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t H3 = H * H * H;
    if (i >= cascade * H3) return;
    if (!((bitfield[i >> 3] >> (i & 7)) & 1)) return;
    const uint32_t level = i / H3, mort = i % H3;
    const float mb = fminf(scalbnf(1.0f, (int)level), bound);
    const uint32_t c[3] = {compact3t(mort), compact3t(mort >> 1), compact3t(mort >> 2)};
    int lo[3], hi[3];
    const int nt = R / T;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p0 = (2.0f * (float)c[a] / (float)H - 1.0f) * mb, p1 = (2.0f * (float)(c[a] + 1) / (float)H - 1.0f) * mb;
        const float x0 = (p0 / bound + 1.0f) * 0.5f * (float)(R - 1), x1 = (p1 / bound + 1.0f) * 0.5f * (float)(R - 1);
        lo[a] = max(0, (int)floorf(x0) - margin) / T;
        hi[a] = min(R - 1, (int)floorf(x1) + 1 + margin) / T;
    }
    // plane 0: (gx, gy) = (x, z); plane 1: (x, y); plane 2: (y, z); gx indexes W (columns), gy indexes H (rows)
    const int ax[3][2] = {{0, 2}, {0, 1}, {1, 2}};
#pragma unroll
    for (int p = 0; p < 3; ++p)
        for (int ty = lo[ax[p][1]]; ty <= hi[ax[p][1]]; ++ty)
            for (int tx = lo[ax[p][0]]; tx <= hi[ax[p][0]]; ++tx) flags[(p * nt + ty) * nt + tx] = 1;
}

// tile id = (p * nt + ty) * nt + tx ; compact layout [n][T][T][C]; one CTA per (tile, group of rows).
// BF16: the compact buffer holds bfloat16 (halves the bytes on NVLink; the plane gradient itself stays fp32).
template <bool PACK, bool BF16>
__global__ void __launch_bounds__(256)
k_tiles_copy(float* __restrict__ planes, void* __restrict__ compact, const int32_t* __restrict__ tile_ids, int R, int C, int T,
             float scale) {
    const int tile = blockIdx.x;
    const int id = tile_ids[tile];
    const int nt = R / T;
    const int p = id / (nt * nt), ty = (id / nt) % nt, tx = id % nt;
    const int n4 = T * C / 4;
    const int rows_per_cta = (T + gridDim.y - 1) / gridDim.y;
    const int row_end = min(T, (int)(blockIdx.y + 1) * rows_per_cta);
    for (int row = blockIdx.y * rows_per_cta; row < row_end; ++row) {
    float4* src = reinterpret_cast<float4*>(planes + (((size_t)p * R + (size_t)ty * T + row) * R + (size_t)tx * T) * C);
    const size_t off4 = (((size_t)tile * T + row) * T) * C / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        if (BF16) {
            uint2* dst = reinterpret_cast<uint2*>(compact) + off4;
            if (PACK) {
                const float4 v = src[i];
                const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
            } else {
                const uint2 u = dst[i];
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                src[i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
            }
        } else {
            float4* dst = reinterpret_cast<float4*>(compact) + off4;
            if (PACK) dst[i] = src[i];
            else {
                float4 v = dst[i];
                v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
                src[i] = v;
            }
        }
    }
    }
}

// zero the listed tiles of planes [3][R][R][C]; the list length is read on the device (grid sized by its capacity)
__global__ void __launch_bounds__(256)
k_tiles_zero(float* __restrict__ planes, const int32_t* __restrict__ tile_ids, const int32_t* __restrict__ count, int R, int C, int T) {
    const int tile = blockIdx.x;
    if (tile >= __ldg(count)) return;
    const int id = __ldg(tile_ids + tile);
    const int nt = R / T;
    const int p = id / (nt * nt), ty = (id / nt) % nt, tx = id % nt;
    const int n4 = T * C / 4;
    const int rows_per_cta = (T + gridDim.y - 1) / gridDim.y;
    const int row_end = min(T, (int)(blockIdx.y + 1) * rows_per_cta);
    for (int row = blockIdx.y * rows_per_cta; row < row_end; ++row) {
        float4* dst = reinterpret_cast<float4*>(planes + (((size_t)p * R + (size_t)ty * T + row) * R + (size_t)tx * T) * C);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// All-reduce of the dirty tiles IN PLACE over NVLink / NVSwitch peer memory (no pack / unpack, exact fp32 sums).
// The plane-gradient buffer of every rank lives in symmetric memory (same size on every rank, peer-mapped; with NVLS also
// mapped through one multicast address).  Tile k of the (replicated) dirty list is reduced by rank k % world:
//   MC   multimem.ld_reduce.add.v4.f32 -- the switch reads the 16 bytes from every rank and returns their sum --, scale by
//        1 / world, multimem.st.v4.f32 -- the switch writes the result into every rank's buffer.  Per rank and direction
//        1 / world of the dirty bytes cross the link;
//   P2P  (no multicast object): ld.global from each peer's mapping, sum in rank order, st.global to each peer's mapping.
// Callers bracket the launch with cross-rank barriers (every rank's scatter done before / every store landed after).
// ------------------------------------------------------------------------------------------------
struct PeerPtrs {
    float* p[8];
};

#ifdef __CUDACC__
__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#else   // host build of the kernels (tests/emu): peers are ordinary buffers; there is no multicast object
inline float4 mc_ld_reduce_add(const float*) { abort(); }
inline void mc_st(float*, float4) { abort(); }
inline float4 peer_ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
inline void peer_st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
#endif

// U independent 16-byte chunks per thread: all loads are issued before the first store, so a thread keeps U x 16 bytes in
// flight over the link (one chunk per thread is latency-bound: 0.93 ms for 360 MB at N = 2, measured)
template <bool MC, int U>
__device__ __forceinline__ void allreduce_batch(const PeerPtrs& peers, float* mc, const size_t (&off)[U], const bool (&ok)[U], int world,
                                                float scale) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        if (MC) {
            v[u] = mc_ld_reduce_add(mc + off[u]);
        } else {
            v[u] = peer_ld(peers.p[0] + off[u]);
        }
    }
    if (!MC) {
        for (int r = 1; r < world; ++r) {
            float4 w[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) w[u] = peer_ld(peers.p[r] + off[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) { v[u].x += w[u].x; v[u].y += w[u].y; v[u].z += w[u].z; v[u].w += w[u].w; }
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
        if (MC) mc_st(mc + off[u], v[u]);
        else
            for (int r = 0; r < world; ++r) peer_st(peers.p[r] + off[u], v[u]);
    }
}

// planes [3][R][R][C] (symmetric).  Work item w = (tile, group of T / 4 rows); a persistent grid strides over the items of the
// tiles this rank owns (the list length is read on the device), so the kernel needs few SM slots and slips in beside the
// scatter of the next plane (its stream has high priority).
template <bool MC>
__global__ void __launch_bounds__(256)
k_tiles_allreduce(PeerPtrs peers, float* mc, const int32_t* __restrict__ tile_ids, const int32_t* __restrict__ count, int R, int C, int T,
                  int rank, int world, float scale) {
    constexpr int U = 8, GROUPS = 4;
    const int n_tiles = __ldg(count);
    const int nt = R / T;
    const int n4 = T * C / 4;                                   // 16-byte chunks per tile row
    const int rows_per_group = (T + GROUPS - 1) / GROUPS;
    const size_t row_stride = (size_t)R * C;
    // owned tiles: rank, rank + world, ...; item index over (owned tile, row group)
    const int n_owned = n_tiles > rank ? (n_tiles - rank + world - 1) / world : 0;
    for (int w = blockIdx.x; w < n_owned * GROUPS; w += gridDim.x) {
        const int tile = rank + (w / GROUPS) * world;
        const int grp = w % GROUPS;
        const int id = __ldg(tile_ids + tile);
        const int p = id / (nt * nt), ty = (id / nt) % nt, tx = id % nt;
        const int row0 = grp * rows_per_group;
        const int nrows = min(T, row0 + rows_per_group) - row0;
        const int total = nrows * n4;
        const size_t base = (((size_t)p * R + (size_t)ty * T + row0) * R + (size_t)tx * T) * C;
        for (int i0 = threadIdx.x; i0 < total; i0 += U * blockDim.x) {
            size_t off[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * blockDim.x;
                ok[u] = i < total;
                const int row = ok[u] ? i / n4 : 0, c4 = ok[u] ? i % n4 : 0;
                off[u] = base + (size_t)row * row_stride + 4 * (size_t)c4;
            }
            allreduce_batch<MC, U>(peers, mc, off, ok, world, scale);
        }
    }
}

// a flat fp32 buffer of n4 16-byte chunks (the MLP weight gradients): rank r reduces the r-th contiguous share
template <bool MC>
__global__ void __launch_bounds__(256)
k_flat_allreduce(PeerPtrs peers, float* mc, uint32_t n4, int rank, int world, float scale) {
    constexpr int U = 4;
    const uint32_t per = (n4 + world - 1) / world;
    const uint32_t lo = rank * per, hi = min(n4, lo + per);
    for (uint32_t i0 = lo + blockIdx.x * blockDim.x * U + threadIdx.x; i0 < hi; i0 += gridDim.x * blockDim.x * U) {
        size_t off[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * blockDim.x;
            ok[u] = i < hi;
            off[u] = 4 * (size_t)(ok[u] ? i : lo);
        }
        allreduce_batch<MC, U>(peers, mc, off, ok, world, scale);
    }
}

}  // namespace tnl

using namespace tnl;

extern "C" {

int tnl_mark_dirty_tiles(const uint8_t* bitfield, uint32_t cascade, uint32_t H, float bound, uint32_t R, uint32_t T,
                         uint32_t margin, uint8_t* flags, tnl_stream_t stream) {
    TNL_ARG_CHECK(bitfield && flags, "null pointer");
    TNL_ARG_CHECK(T >= 4 && R % T == 0, "tile size must divide R");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t nt = R / T;
    cudaMemsetAsync(flags, 0, 3 * nt * nt, s);
    const uint32_t n = cascade * H * H * H;
    k_mark_dirty_tiles<<<ceil_div(n, 256u), 256, 0, s>>>(bitfield, cascade, H, bound, (int)R, (int)T, (int)margin, flags);
    return finish_launch("mark_dirty_tiles");
}

int tnl_tiles_zero(float* planes, const int32_t* tile_ids, const int32_t* count, uint32_t capacity, uint32_t R, uint32_t C, uint32_t T,
                   tnl_stream_t stream) {
    if (capacity == 0) return 0;
    TNL_ARG_CHECK(planes && tile_ids && count, "null pointer");
    TNL_ARG_CHECK(R % T == 0 && (T * C) % 4 == 0, "bad tile geometry");
    k_tiles_zero<<<dim3(capacity, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(planes, tile_ids, count, (int)R, (int)C, (int)T);
    return finish_launch("tiles_zero");
}

int tnl_tiles_pack(const float* planes, const int32_t* tile_ids, uint32_t n_tiles, uint32_t R, uint32_t C, uint32_t T,
                   void* compact, int bf16, tnl_stream_t stream) {
    if (n_tiles == 0) return 0;
    TNL_ARG_CHECK(planes && tile_ids && compact, "null pointer");
    TNL_ARG_CHECK(R % T == 0 && (T * C) % 4 == 0, "bad tile geometry");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (bf16) k_tiles_copy<true, true><<<dim3(n_tiles, 4), 256, 0, s>>>(const_cast<float*>(planes), compact, tile_ids, (int)R, (int)C, (int)T, 1.0f);
    else k_tiles_copy<true, false><<<dim3(n_tiles, 4), 256, 0, s>>>(const_cast<float*>(planes), compact, tile_ids, (int)R, (int)C, (int)T, 1.0f);
    return finish_launch("tiles_pack");
}

int tnl_tiles_unpack(const void* compact, const int32_t* tile_ids, uint32_t n_tiles, uint32_t R, uint32_t C, uint32_t T,
                     float scale, int bf16, float* planes, tnl_stream_t stream) {
    if (n_tiles == 0) return 0;
    TNL_ARG_CHECK(planes && tile_ids && compact, "null pointer");
    TNL_ARG_CHECK(R % T == 0 && (T * C) % 4 == 0, "bad tile geometry");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (bf16) k_tiles_copy<false, true><<<dim3(n_tiles, 4), 256, 0, s>>>(planes, const_cast<void*>(compact), tile_ids, (int)R, (int)C, (int)T, scale);
    else k_tiles_copy<false, false><<<dim3(n_tiles, 4), 256, 0, s>>>(planes, const_cast<void*>(compact), tile_ids, (int)R, (int)C, (int)T, scale);
    return finish_launch("tiles_unpack");
}

static bool fill_peers(PeerPtrs& pp, const void* const* peers, uint32_t world) {
    if (world < 1 || world > 8) return false;
    for (uint32_t r = 0; r < 8; ++r) pp.p[r] = r < world && peers ? static_cast<float*>(const_cast<void*>(peers[r])) : nullptr;
    return true;
}

int tnl_tiles_allreduce(void* multicast, const void* const* peers, const int32_t* tile_ids, const int32_t* count, uint32_t capacity,
                        uint32_t R, uint32_t C, uint32_t T, uint32_t rank, uint32_t world, float scale, tnl_stream_t stream) {
    if (capacity == 0) return 0;
    TNL_ARG_CHECK(tile_ids && count && (multicast || peers), "null pointer");
    TNL_ARG_CHECK(R % T == 0 && (T * C) % 4 == 0 && rank < world, "bad tile geometry / rank");
    PeerPtrs pp;
    TNL_ARG_CHECK(fill_peers(pp, peers, world), "world size must be 1..8 (one NVSwitch domain)");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t blocks = min(ceil_div(capacity, world) * 4u, (uint32_t)kNumSM * 3u);      // persistent: a few CTAs per SM
    if (multicast)
        k_tiles_allreduce<true><<<blocks, 256, 0, s>>>(pp, static_cast<float*>(multicast), tile_ids, count, (int)R, (int)C, (int)T,
                                                    (int)rank, (int)world, scale);
    else
        k_tiles_allreduce<false><<<blocks, 256, 0, s>>>(pp, nullptr, tile_ids, count, (int)R, (int)C, (int)T, (int)rank, (int)world, scale);
    return finish_launch("tiles_allreduce");
}

int tnl_flat_allreduce(void* multicast, const void* const* peers, uint32_t n_floats, uint32_t rank, uint32_t world, float scale,
                       tnl_stream_t stream) {
    if (n_floats == 0) return 0;
    TNL_ARG_CHECK((multicast || peers) && n_floats % 4 == 0 && rank < world, "bad argument (the buffer length must be a multiple of 4 floats)");
    PeerPtrs pp;
    TNL_ARG_CHECK(fill_peers(pp, peers, world), "world size must be 1..8 (one NVSwitch domain)");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t n4 = n_floats / 4;
    const uint32_t blocks = min(ceil_div(ceil_div(n4, world), 1024u), (uint32_t)kNumSM);
    if (multicast) k_flat_allreduce<true><<<blocks, 256, 0, s>>>(pp, static_cast<float*>(multicast), n4, (int)rank, (int)world, scale);
    else k_flat_allreduce<false><<<blocks, 256, 0, s>>>(pp, nullptr, n4, (int)rank, (int)world, scale);
    return finish_launch("flat_allreduce");
}

}  // extern "C"
