// Tile-binned tri-plane sampling (gather) and its adjoint (scatter) through shared memory -- sm_100a.  OPT-IN variant of
// sample.cu (same behavioural contract: triplane_encoder.py:314-332 + ATen grid_sampler_2d arithmetic), written after the
// round-1 profiles showed both sampling kernels at the L2 roofline (forward 11.4 TB/s of corner reads out of L2, backward
// bound by L2 vector atomics): to move fewer bytes than one corner fetch per (point, plane, corner) the texels have to be
// reused on chip.
//
// Idea.  The step's points are counting-sorted by the TILE OF THEIR NORTH-WEST TAP on every axis (tnl_tap_sort: key =
// (t_z * G + t_y) * G + t_x, t_a = floor(ix_a) / TS with ix_a computed exactly as the sampler computes it, fp16 rounding of the
// projected coordinate included).  A plane tile (p, ty, tx) of TS x TS texels is then touched by the points of the G bins
// that share its two in-plane tile indices, whatever their index along the third axis.  One CTA per plane tile:
//   forward : stage the (TS+1)^2 x C texel tile (tile + the one-texel halo of the south / east taps) in shared memory once,
//             stream the points of its G bins through it; the four corner reads of a point are shared-memory reads.
//   backward: accumulate the corner contributions of the tile's own points in shared memory (shared atomics), write the
//             tile's TS x TS texels with plain stores and park the one-texel south / east halo in a small per-tile strip; a
//             second, tiny pass adds each tile's west / north / north-west neighbour strips to its first column / row.
//             No global atomics and no zero fill of the gradient buffer: every listed tile is written (zeros if nothing
//             landed in it), every texel by exactly one CTA per pass.
// Per (point, plane) the backward reads 12 B of coordinates and 2*C B of feature gradient instead of issuing four C*4-byte
// vector atomics to L2.
//
// Results: forward bit-identical to sample.cu (same taps, same FMA order); backward identical up to the order of the
// float additions (as between any two runs of sample.cu).
#include "common.cuh"
#include "sample_coords.cuh"

namespace tnl {

constexpr int kTS = 32;          // tile size in texels (= the dirty-tile size of tiles.cu / idwt_plan.py)
constexpr int kTW = kTS + 1;     // tile + south / east halo

template <int C4>
struct TCfg {
    static constexpr int C = 4 * C4;
    static constexpr int PITCH = C + 4;                 // floats per texel in shared memory (+4: spreads texels over the banks)
    // points in flight per iteration.  The tile fills most of an SM's shared memory (one CTA per SM), so the CTA itself has
    // to bring the warps that hide the latency of the perm / coordinate / feature-row accesses: 24-32 warps
    static constexpr int SLOTS = C4 <= 4 ? 256 : (C4 <= 8 ? 128 : 64);
    static constexpr int NT = SLOTS * C4;               // thread <-> (point slot, 4-channel group); 1024 / 1024 / 768
    static constexpr size_t TILE_F = sizeof(float) * kTW * kTW * PITCH;   // forward: tile + halo
    static size_t smem_fwd(uint32_t G) { return TILE_F + 2 * sizeof(uint32_t) * G; }        // + bin table (G entries)
    static size_t smem_bwd(uint32_t G) { return TILE_F + 2 * sizeof(uint32_t) * G; }        // backward accumulates tile + halo too
};

struct TileGeom {
    int p, tx, ty;     // plane, tile column (gx axis), tile row (gy axis)
    int a, b, c;       // world axes of gx, gy and of the direction the tile is projected along
};

__device__ __forceinline__ bool tile_geom(const int32_t* __restrict__ tile_ids, const int32_t* __restrict__ n_tiles, int G,
                                          TileGeom& t) {
    int id = (int)blockIdx.x;
    if (tile_ids != nullptr) {
        if (id >= __ldg(n_tiles)) return false;
        id = __ldg(tile_ids + id);
    }
    t.tx = id % G;
    t.ty = (id / G) % G;
    t.p = id / (G * G);
    t.a = (t.p == 2) ? 1 : 0;
    t.b = (t.p == 1) ? 1 : 2;
    t.c = 3 - t.a - t.b;
    return true;
}

// [start, end) of the points whose tap tile is (column ta on axis a, row tb on axis b, k on the third axis)
__device__ __forceinline__ void bin_range(const uint32_t* __restrict__ bin_end, const TileGeom& t, int G, int ta, int tb, int k,
                                          uint32_t& start, uint32_t& end) {
    uint32_t tc[3];
    tc[t.a] = (uint32_t)ta;
    tc[t.b] = (uint32_t)tb;
    tc[t.c] = (uint32_t)k;
    const uint32_t bin = (tc[2] * (uint32_t)G + tc[1]) * (uint32_t)G + tc[0];
    start = bin ? __ldg(bin_end + bin - 1) : 0u;
    end = __ldg(bin_end + bin);
}

// The points of a tile are spread over G (forward) or 4G (backward) bins.  Walking the bins one after the other would
// serialise three dependent global loads (bin range -> perm -> coordinates) per bin; instead the bin table is built once in
// shared memory -- pref[k] = number of points in bins 0..k, gend[k] = end offset of bin k in perm -- and the CTA iterates
// over the flat point index j: bin = first k with pref[k] > j, perm index = gend[k] - (pref[k] - j).
struct BinTable {
    uint32_t* pref;   // [n] inclusive prefix of the bin sizes
    uint32_t* gend;   // [n] end offset of the bin in perm
    int n;
    uint32_t total;
};

// n <= 1024 entries; every thread of the CTA calls it (two barriers)
template <typename F>
__device__ __forceinline__ BinTable build_bin_table(uint32_t* smem_words, int n, F range_of) {
    BinTable bt;
    bt.pref = smem_words;
    bt.gend = smem_words + n;
    bt.n = n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        uint32_t start, end;
        range_of(k, start, end);
        bt.pref[k] = end - start;
        bt.gend[k] = end;
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // inclusive scan of the bin sizes by one warp, 32 entries per step (n is at most 1024)
        const int lane = threadIdx.x;
        uint32_t carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int k = base + lane;
            uint32_t v = k < n ? bt.pref[k] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += up;
            }
            if (k < n) bt.pref[k] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    bt.total = n > 0 ? bt.pref[n - 1] : 0u;
    return bt;
}

__device__ __forceinline__ uint32_t bin_lookup(const BinTable& bt, uint32_t j) {
    int lo = 0, hi = bt.n - 1;          // j < total, so the answer exists
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bt.pref[mid] > j) hi = mid; else lo = mid + 1;
    }
    return bt.gend[lo] - (bt.pref[lo] - j);
}

constexpr int kUnroll = 4;   // independent points per thread in flight (perm -> coordinates -> feature row are dependent loads)

template <int C4, bool HALF>
__global__ void __launch_bounds__(TCfg<C4>::NT, 1)
k_tsample_fwd(const float* __restrict__ planes, const float* __restrict__ xyz, int R, int G, float inv_bound, int fp16_coords,
              const uint32_t* __restrict__ bin_end, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_ids,
              const int32_t* __restrict__ n_tiles, void* __restrict__ feat_) {
    using Cfg = TCfg<C4>;
    constexpr int C = Cfg::C;
    extern __shared__ __align__(16) float tile[];
    TileGeom t;
    if (!tile_geom(tile_ids, n_tiles, G, t)) return;
    const int tid = threadIdx.x, cq = tid % C4, slot = tid / C4;
    const int x_base = t.tx * kTS, y_base = t.ty * kTS;
    uint32_t* words = reinterpret_cast<uint32_t*>(tile + kTW * kTW * Cfg::PITCH);
    const BinTable bt = build_bin_table(words, G, [&](int k, uint32_t& start, uint32_t& end) {
        bin_range(bin_end, t, G, t.tx, t.ty, k, start, end);
    });
    if (bt.total == 0) return;   // a listed tile without points (the halo tiles of a work-list step): nothing to stage
    // stage the tile: rows y_base .. y_base+TS, columns x_base .. x_base+TS, clipped to the plane
    const int rows = min(kTW, R - y_base), cols = min(kTW, R - x_base);
    for (int i = tid; i < rows * cols * C4; i += Cfg::NT) {
        const int q = i % C4, xy = i / C4, lx = xy % cols, ly = xy / cols;
        const float4 v = __ldg(reinterpret_cast<const float4*>(planes + (((size_t)t.p * R + (y_base + ly)) * R + (x_base + lx)) * C) + q);
        *reinterpret_cast<float4*>(tile + (ly * kTW + lx) * Cfg::PITCH + 4 * q) = v;
    }
    __syncthreads();
    for (uint32_t j0 = slot; j0 < bt.total; j0 += kUnroll * Cfg::SLOTS) {
        uint32_t m[kUnroll];
        float gx[kUnroll], gy[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t j = j0 + u * Cfg::SLOTS;
            m[u] = j < bt.total ? (uint32_t)__ldg(perm + bin_lookup(bt, j)) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (m[u] != 0xffffffffu) plane_coords(xyz, m[u], t.p, inv_bound, fp16_coords, gx[u], gy[u]);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (m[u] == 0xffffffffu) continue;
            const Tap tp = make_tap(gx[u], gy[u], R);
            const float* base = tile + ((tp.y0 - y_base) * kTW + (tp.x0 - x_base)) * Cfg::PITCH + 4 * cq;
            float4 acc;
            {
                const float4 v = *reinterpret_cast<const float4*>(base);
                acc.x = v.x * tp.nw; acc.y = v.y * tp.nw; acc.z = v.z * tp.nw; acc.w = v.w * tp.nw;
            }
            if (tp.x1ok) {
                const float4 v = *reinterpret_cast<const float4*>(base + Cfg::PITCH);
                acc.x = fmaf(v.x, tp.ne, acc.x); acc.y = fmaf(v.y, tp.ne, acc.y); acc.z = fmaf(v.z, tp.ne, acc.z); acc.w = fmaf(v.w, tp.ne, acc.w);
            }
            if (tp.y1ok) {
                const float4 v = *reinterpret_cast<const float4*>(base + kTW * Cfg::PITCH);
                acc.x = fmaf(v.x, tp.sw, acc.x); acc.y = fmaf(v.y, tp.sw, acc.y); acc.z = fmaf(v.z, tp.sw, acc.z); acc.w = fmaf(v.w, tp.sw, acc.w);
            }
            if (tp.x1ok && tp.y1ok) {
                const float4 v = *reinterpret_cast<const float4*>(base + (kTW + 1) * Cfg::PITCH);
                acc.x = fmaf(v.x, tp.se, acc.x); acc.y = fmaf(v.y, tp.se, acc.y); acc.z = fmaf(v.z, tp.se, acc.z); acc.w = fmaf(v.w, tp.se, acc.w);
            }
            const size_t q4 = ((size_t)m[u] * 3 + t.p) * C4 + cq;   // this thread's 4-channel group of feat[m][p*C ..]
            if (HALF) reinterpret_cast<uint2*>(feat_)[q4] = pack4h(acc);
            else reinterpret_cast<float4*>(feat_)[q4] = acc;
        }
    }
}

// rows >= *n_valid are visited by no tile (they sit in the padding bin): their feature rows are zeros, as sample.cu writes
// them.  One thread per row; in the steady state only the few padding rows do any work.
template <bool HALF>
__global__ void k_tsample_zero_tail(void* __restrict__ feat_, uint32_t M, uint32_t row_quads, const int32_t* __restrict__ n_valid) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M || (int32_t)m < __ldg(n_valid)) return;
    for (uint32_t q = 0; q < row_quads; ++q) {
        const size_t i = (size_t)m * row_quads + q;
        if (HALF) reinterpret_cast<uint2*>(feat_)[i] = make_uint2(0u, 0u);
        else reinterpret_cast<float4*>(feat_)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Backward, pass 1.  The CTA of tile (p, ty, tx) accumulates ALL four corner contributions of its own points (the G bins of
// its column) in a (TS+1)^2 shared-memory tile, writes its own TS x TS texels to g_planes with plain stores, and parks the
// south / east halo -- contributions that belong to the neighbours' first row / column -- in `halo[tile]` (65 texels:
// row TS with lx = 0..TS, then column TS with ly = 0..TS-1).  Pass 2 (k_tsample_halo_fix) lets every tile add the strips of
// its west / north / north-west neighbours to its own first column / row.  Every texel is written by exactly one CTA in
// each pass: no atomics on global memory, no zero fill, and no CTA scans another tile's points.
constexpr int kHaloTexels = 2 * kTS + 1;

template <int C4, bool HALF>
__global__ void __launch_bounds__(TCfg<C4>::NT, 1)
k_tsample_bwd(const void* __restrict__ g_feat_, const float* __restrict__ xyz, int R, int G, float inv_bound, int fp16_coords,
              const uint32_t* __restrict__ bin_end, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_ids,
              const int32_t* __restrict__ n_tiles, float* __restrict__ g_planes, float* __restrict__ halo) {
    using Cfg = TCfg<C4>;
    constexpr int C = Cfg::C;
    extern __shared__ __align__(16) float tile[];
    TileGeom t;
    if (!tile_geom(tile_ids, n_tiles, G, t)) return;
    const int tid = threadIdx.x, cq = tid % C4, slot = tid / C4;
    const int x_base = t.tx * kTS, y_base = t.ty * kTS;
    uint32_t* words = reinterpret_cast<uint32_t*>(tile + kTW * kTW * Cfg::PITCH);
    const BinTable bt = build_bin_table(words, G, [&](int k, uint32_t& start, uint32_t& end) {
        bin_range(bin_end, t, G, t.tx, t.ty, k, start, end);
    });
    const int id = (t.p * G + t.ty) * G + t.tx;
    float4* strip = reinterpret_cast<float4*>(halo + (size_t)id * kHaloTexels * C);
    const int rows = min(kTS, R - y_base), cols = min(kTS, R - x_base);
    if (bt.total == 0) {   // a listed tile without points (the halo tiles of a work-list step): zeros, straight from registers
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < rows * cols * C4; i += Cfg::NT) {
            const int q = i % C4, xy = i / C4, lx = xy % cols, ly = xy / cols;
            *(reinterpret_cast<float4*>(g_planes + (((size_t)t.p * R + (y_base + ly)) * R + (x_base + lx)) * C) + q) = z;
        }
        for (int i = tid; i < kHaloTexels * C4; i += Cfg::NT) strip[i] = z;
        return;
    }
    for (int i = tid; i < kTW * kTW * Cfg::PITCH / 4; i += Cfg::NT) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (uint32_t j0 = slot; j0 < bt.total; j0 += kUnroll * Cfg::SLOTS) {
        uint32_t m[kUnroll];
        float gx[kUnroll], gy[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t j = j0 + u * Cfg::SLOTS;
            m[u] = j < bt.total ? (uint32_t)__ldg(perm + bin_lookup(bt, j)) : 0xffffffffu;
        }
        float4 g[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (m[u] == 0xffffffffu) continue;
            plane_coords(xyz, m[u], t.p, inv_bound, fp16_coords, gx[u], gy[u]);
            const size_t q4 = ((size_t)m[u] * 3 + t.p) * C4 + cq;
            g[u] = HALF ? unpack4h(__ldg(reinterpret_cast<const uint2*>(g_feat_) + q4))
                        : __ldg(reinterpret_cast<const float4*>(g_feat_) + q4);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (m[u] == 0xffffffffu) continue;
            const Tap tp = make_tap(gx[u], gy[u], R);
            float* base = tile + ((tp.y0 - y_base) * kTW + (tp.x0 - x_base)) * Cfg::PITCH + 4 * cq;
            smem_add4(base, g[u], tp.nw);
            if (tp.x1ok) smem_add4(base + Cfg::PITCH, g[u], tp.ne);
            if (tp.y1ok) smem_add4(base + kTW * Cfg::PITCH, g[u], tp.sw);
            if (tp.x1ok && tp.y1ok) smem_add4(base + (kTW + 1) * Cfg::PITCH, g[u], tp.se);
        }
    }
    __syncthreads();
    for (int i = tid; i < rows * cols * C4; i += Cfg::NT) {
        const int q = i % C4, xy = i / C4, lx = xy % cols, ly = xy / cols;
        const float4 v = *reinterpret_cast<const float4*>(tile + (ly * kTW + lx) * Cfg::PITCH + 4 * q);
        *(reinterpret_cast<float4*>(g_planes + (((size_t)t.p * R + (y_base + ly)) * R + (x_base + lx)) * C) + q) = v;
    }
    for (int i = tid; i < kHaloTexels * C4; i += Cfg::NT) {
        const int q = i % C4, e = i / C4;
        const int lx = e <= kTS ? e : kTS, ly = e <= kTS ? kTS : e - kTW;   // row TS (lx 0..TS), then column TS (ly 0..TS-1)
        strip[e * C4 + q] = *reinterpret_cast<const float4*>(tile + (ly * kTW + lx) * Cfg::PITCH + 4 * q);
    }
}

// Backward, pass 2: first column / first row of every listed tile += the halo strips of its west / north / north-west
// neighbours (those that were processed in pass 1: tile_map[id] != 0, or every tile when tile_map is NULL).
template <int C4>
__global__ void __launch_bounds__(256)
k_tsample_halo_fix(int R, int G, const int32_t* __restrict__ tile_ids, const int32_t* __restrict__ n_tiles,
                   const uint8_t* __restrict__ tile_map, const float* __restrict__ halo, float* __restrict__ g_planes) {
    constexpr int C = 4 * C4;
    TileGeom t;
    if (!tile_geom(tile_ids, n_tiles, G, t)) return;
    const int x_base = t.tx * kTS, y_base = t.ty * kTS;
    auto listed = [&](int tx, int ty) {
        if (tx < 0 || ty < 0) return false;
        const int id = (t.p * G + ty) * G + tx;
        return tile_map == nullptr || tile_map[id] != 0;
    };
    const bool west = listed(t.tx - 1, t.ty), north = listed(t.tx, t.ty - 1), nw = listed(t.tx - 1, t.ty - 1);
    auto strip_of = [&](int tx, int ty) {
        return reinterpret_cast<const float4*>(halo + (size_t)((t.p * G + ty) * G + tx) * kHaloTexels * C);
    };
    // work items: 0 = corner (0,0); 1..TS-1 = first row (lx, 0); TS..2TS-2 = first column (0, ly)
    for (int i = threadIdx.x; i < (2 * kTS - 1) * C4; i += blockDim.x) {
        const int q = i % C4, e = i / C4;
        const int lx = (e > 0 && e < kTS) ? e : 0, ly = e >= kTS ? e - kTS + 1 : 0;
        if (x_base + lx >= R || y_base + ly >= R) continue;
        float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lx == 0 && west) {          // west neighbour's column TS, row ly
            const float4 v = __ldg(strip_of(t.tx - 1, t.ty) + (kTW + ly) * C4 + q);
            add.x += v.x; add.y += v.y; add.z += v.z; add.w += v.w;
        }
        if (ly == 0 && north) {         // north neighbour's row TS, column lx
            const float4 v = __ldg(strip_of(t.tx, t.ty - 1) + lx * C4 + q);
            add.x += v.x; add.y += v.y; add.z += v.z; add.w += v.w;
        }
        if (lx == 0 && ly == 0 && nw) { // north-west neighbour's corner (TS, TS)
            const float4 v = __ldg(strip_of(t.tx - 1, t.ty - 1) + kTS * C4 + q);
            add.x += v.x; add.y += v.y; add.z += v.z; add.w += v.w;
        }
        float4* dst = reinterpret_cast<float4*>(g_planes + (((size_t)t.p * R + (y_base + ly)) * R + (x_base + lx)) * C) + q;
        float4 cur = *dst;
        cur.x += add.x; cur.y += add.y; cur.z += add.z; cur.w += add.w;
        *dst = cur;
    }
}

// ---- tap-tile histogram (pass 1 of the counting sort; the scan and scatter passes are those of sort.cu) ----
__global__ void k_tap_hist(const float* __restrict__ xyz, uint32_t M, const int32_t* __restrict__ n_valid, float inv_bound,
                           int fp16_coords, int R, int G, uint32_t* __restrict__ hist, uint32_t* __restrict__ keys) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const uint32_t nv = n_valid ? (uint32_t)max(*n_valid, 0) : M;
    uint32_t key = (uint32_t)(G * G * G);
    if (m < nv) {
        uint32_t tc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float g = axis_coord(xyz, m, a, inv_bound, fp16_coords);
            tc[a] = (uint32_t)((int)floorf(to_pixel(g, R)) / kTS);
        }
        key = (tc[2] * (uint32_t)G + tc[1]) * (uint32_t)G + tc[0];
    }
    keys[m] = key;
    atomicAdd(hist + key, 1u);
}

template <int C4>
static int launch_tsample(bool fwd, const void* in, void* out, const float* xyz, uint32_t R, uint32_t G, float inv_bound,
                          int fp16_coords, const uint32_t* bin_end, const int32_t* perm, const int32_t* tile_ids,
                          const int32_t* n_tiles, uint32_t grid, int half, cudaStream_t s, const uint8_t* tile_map = nullptr,
                          float* halo = nullptr) {
    using Cfg = TCfg<C4>;
    const size_t smem_f = Cfg::smem_fwd(G), smem_b = Cfg::smem_bwd(G);
    static bool attr = false;
    if (!attr) {   // opt in to the largest dynamic shared-memory size these kernels can ask for (G <= 256)
        cudaFuncSetAttribute(k_tsample_fwd<C4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_fwd(256));
        cudaFuncSetAttribute(k_tsample_fwd<C4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_fwd(256));
        cudaFuncSetAttribute(k_tsample_bwd<C4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bwd(256));
        cudaFuncSetAttribute(k_tsample_bwd<C4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bwd(256));
        attr = true;
    }
    if (fwd) {
        if (half) k_tsample_fwd<C4, true><<<grid, Cfg::NT, smem_f, s>>>(static_cast<const float*>(in), xyz, (int)R, (int)G, inv_bound, fp16_coords, bin_end, perm, tile_ids, n_tiles, out);
        else k_tsample_fwd<C4, false><<<grid, Cfg::NT, smem_f, s>>>(static_cast<const float*>(in), xyz, (int)R, (int)G, inv_bound, fp16_coords, bin_end, perm, tile_ids, n_tiles, out);
    } else {
        if (half) k_tsample_bwd<C4, true><<<grid, Cfg::NT, smem_b, s>>>(in, xyz, (int)R, (int)G, inv_bound, fp16_coords, bin_end, perm, tile_ids, n_tiles, static_cast<float*>(out), halo);
        else k_tsample_bwd<C4, false><<<grid, Cfg::NT, smem_b, s>>>(in, xyz, (int)R, (int)G, inv_bound, fp16_coords, bin_end, perm, tile_ids, n_tiles, static_cast<float*>(out), halo);
        k_tsample_halo_fix<C4><<<grid, 256, 0, s>>>((int)R, (int)G, tile_ids, n_tiles, tile_map, halo, static_cast<float*>(out));
    }
    return 0;
}

// sort.cu
void counting_sort_finish(uint32_t* hist, const uint32_t* keys, uint32_t* sums, uint32_t nbins, uint32_t M, int32_t* perm,
                          cudaStream_t s);
size_t counting_sort_workspace(size_t nbins, uint32_t M);

}  // namespace tnl

using namespace tnl;

static bool tsample_geometry_ok(uint32_t R, uint32_t C) {
    return R >= (uint32_t)kTS && R % kTS == 0 && R / kTS <= 256 && (C == 16 || C == 32 || C == 48);
}

extern "C" {

size_t tnl_tap_sort_workspace(uint32_t M, uint32_t R) {
    if (R < (uint32_t)kTS || R % kTS != 0 || R / kTS > 256) return 0;
    const size_t G = R / kTS;
    return counting_sort_workspace(G * G * G + 1, M);
}

int tnl_tap_sort(const float* xyz, uint32_t M, const int32_t* n_valid, float inv_bound, int fp16_coords, uint32_t R, int32_t* perm,
                 void* workspace, size_t workspace_bytes, tnl_stream_t stream) {
    TNL_ARG_CHECK(R >= (uint32_t)kTS && R % kTS == 0 && R / kTS <= 256, "R must be a multiple of 32, at most 8192");
    TNL_ARG_CHECK(perm || M == 0, "null pointer");
    if (workspace == nullptr || workspace_bytes < tnl_tap_sort_workspace(M, R)) {
        set_error("tap_sort: workspace too small");
        return TNL_ERR_WORKSPACE;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t G = R / kTS, nbins = G * G * G + 1;
    uint32_t* hist = static_cast<uint32_t*>(workspace);
    uint32_t* keys = hist + nbins;
    uint32_t* sums = keys + M;
    cudaMemsetAsync(hist, 0, sizeof(uint32_t) * nbins, s);     // M == 0: every bin is empty
    if (M == 0) return finish_launch("tap_sort");
    TNL_ARG_CHECK(xyz, "null pointer");
    k_tap_hist<<<ceil_div(M, 256u), 256, 0, s>>>(xyz, M, n_valid, inv_bound, fp16_coords, (int)R, (int)G, hist, keys);
    counting_sort_finish(hist, keys, sums, nbins, M, perm, s);
    return finish_launch("tap_sort");
}

int tnl_tsample_forward(const float* planes, const float* xyz, uint32_t M, uint32_t R, uint32_t C, float inv_bound, int fp16_coords,
                        const int32_t* n_valid, const int32_t* perm, const void* bin_end, const int32_t* tile_ids,
                        const int32_t* n_tiles, uint32_t max_tiles, void* feat, int feat_fp16, tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(planes && xyz && perm && bin_end && feat, "null pointer");
    TNL_ARG_CHECK(tsample_geometry_ok(R, C), "tile-binned sampling needs R % 32 == 0, R <= 8192 and C in {16, 32, 48}");
    TNL_ARG_CHECK((tile_ids == nullptr) == (n_tiles == nullptr), "tile_ids and n_tiles must be given together");
    TNL_ARG_CHECK(((uintptr_t)planes & 15) == 0 && ((uintptr_t)feat & 15) == 0, "planes/feat must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t G = R / kTS;
    const uint32_t grid = tile_ids ? max_tiles : 3 * G * G;
    if (n_valid != nullptr) {   // rows past *n_valid sit in the padding bin: zero their feature rows (contract of sample.cu)
        const uint32_t rq = 3 * C / 4;
        if (feat_fp16) k_tsample_zero_tail<true><<<ceil_div(M, 256u), 256, 0, s>>>(feat, M, rq, n_valid);
        else k_tsample_zero_tail<false><<<ceil_div(M, 256u), 256, 0, s>>>(feat, M, rq, n_valid);
    }
    if (grid == 0) return finish_launch("tsample_forward");
    const uint32_t* be = static_cast<const uint32_t*>(bin_end);
    if (C == 16) launch_tsample<4>(true, planes, feat, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s);
    else if (C == 32) launch_tsample<8>(true, planes, feat, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s);
    else launch_tsample<12>(true, planes, feat, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s);
    return finish_launch("tsample_forward");
}

size_t tnl_tsample_backward_workspace(uint32_t R, uint32_t C) {
    if (!tsample_geometry_ok(R, C)) return 0;
    const size_t G = R / kTS;
    return sizeof(float) * 3 * G * G * (size_t)kHaloTexels * C;
}

int tnl_tsample_backward(const void* g_feat, int feat_fp16, const float* xyz, uint32_t M, uint32_t R, uint32_t C, float inv_bound,
                         int fp16_coords, const int32_t* perm, const void* bin_end, const int32_t* tile_ids, const int32_t* n_tiles,
                         uint32_t max_tiles, const uint8_t* tile_map, float* g_planes, void* workspace, size_t workspace_bytes,
                         tnl_stream_t stream) {
    TNL_ARG_CHECK(g_planes && bin_end, "null pointer");
    TNL_ARG_CHECK(M == 0 || (g_feat && xyz && perm), "null pointer");
    TNL_ARG_CHECK(tsample_geometry_ok(R, C), "tile-binned sampling needs R % 32 == 0, R <= 8192 and C in {16, 32, 48}");
    TNL_ARG_CHECK((tile_ids == nullptr) == (n_tiles == nullptr), "tile_ids and n_tiles must be given together");
    TNL_ARG_CHECK(tile_ids != nullptr || tile_map == nullptr, "a tile map only makes sense with a tile list");
    TNL_ARG_CHECK(((uintptr_t)g_planes & 15) == 0 && ((uintptr_t)g_feat & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                  "g_planes / g_feat / workspace must be 16-byte aligned");
    if (workspace == nullptr || workspace_bytes < tnl_tsample_backward_workspace(R, C)) {
        set_error("tsample_backward: workspace too small");
        return TNL_ERR_WORKSPACE;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t G = R / kTS;
    const uint32_t grid = tile_ids ? max_tiles : 3 * G * G;
    if (grid == 0) return 0;
    const uint32_t* be = static_cast<const uint32_t*>(bin_end);
    float* halo = static_cast<float*>(workspace);
    if (C == 16) launch_tsample<4>(false, g_feat, g_planes, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s, tile_map, halo);
    else if (C == 32) launch_tsample<8>(false, g_feat, g_planes, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s, tile_map, halo);
    else launch_tsample<12>(false, g_feat, g_planes, xyz, R, G, inv_bound, fp16_coords, be, perm, tile_ids, n_tiles, grid, feat_fp16, s, tile_map, halo);
    return finish_launch("tsample_backward");
}

}  // extern "C"
