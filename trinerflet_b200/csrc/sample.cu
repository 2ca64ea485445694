// Tri-plane bilinear sampling (gather) and its adjoint (scatter) -- sm_100a, channels-last planes.
//
// Behavioural contract: TriPlaneVolume.forward -> sample_from_planes_aux
// (/root/reference/reconstruction/triplaneencoder/triplane_encoder.py:314-332, 523-530):
//   u = coords / bound ; plane 0 samples (u_x,u_z), plane 1 (u_x,u_y), plane 2 (u_y,u_z) (:250-289);
//   F.grid_sample(bilinear, padding_mode='border', align_corners=True); concat -> feat[m][p*C + c].
// Arithmetic follows ATen's grid_sampler_2d (unnormalise ((g+1)/2)*(R-1), clip to [0,R-1], weights
// nw/ne/sw/se as products of differences, accumulate in that order, out-of-range corners skipped).
//
// Why channels-last: with planes [3][R][R][C] a texel's C features are one contiguous 4*C-byte run, so
// a bilinear tap is 2 x (two adjacent texels) = two contiguous 8*C-byte reads; every 32-byte sector
// fetched is fully used.  The reference's NCHW layout fetches 2 sectors per (plane, channel) for 16
// useful bytes.
//
// Mapping: one thread per (point, plane, 4-channel group): four 128-bit loads (one per corner),
// lerp in registers, one 128-bit store into the contiguous feature row.  Backward: one 128-bit
// load of the feature gradient and four vector atomic adds (red.global.add.v4.f32).
#include "common.cuh"
#include "sample_coords.cuh"

namespace tnl {

// HALF: features are written as fp16 -- the rounding the first nn.Linear applies under autocast anyway (network.py:127),
// moved into the producer so that the feature stream costs half the bytes.
template <bool HALF>
__global__ void __launch_bounds__(256)
k_sample_fwd(const float* __restrict__ planes, const float* __restrict__ xyz, uint32_t M, int R, int C, float inv_bound,
             int fp16_coords, const int32_t* __restrict__ n_valid, const int32_t* __restrict__ perm,
             void* __restrict__ feat_) {
    const int cq_per = C >> 2;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)M * 3 * cq_per;
    if (idx >= total) return;
    const int cq = (int)(idx % cq_per);
    const int p = (int)((idx / cq_per) % 3);
    uint32_t m = (uint32_t)(idx / ((uint64_t)3 * cq_per));
    if (perm) m = (uint32_t)__ldg(perm + m);
    const size_t q4 = ((size_t)m * 3 * C + (size_t)p * C) / 4 + cq;   // index of this thread's 4-channel group
    if (n_valid && (int32_t)m >= *n_valid) {
        if (HALF) reinterpret_cast<uint2*>(feat_)[q4] = make_uint2(0u, 0u);
        else reinterpret_cast<float4*>(feat_)[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    float gx, gy;
    plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
    const Tap t = make_tap(gx, gy, R);
    const float4* base = reinterpret_cast<const float4*>(planes + (((size_t)p * R + t.y0) * R + t.x0) * C) + cq;
    const size_t dx = (size_t)cq_per, dy = (size_t)R * cq_per;
    float4 acc;
    {
        const float4 v = __ldg(base);
        acc.x = v.x * t.nw; acc.y = v.y * t.nw; acc.z = v.z * t.nw; acc.w = v.w * t.nw;
    }
    if (t.x1ok) {
        const float4 v = __ldg(base + dx);
        acc.x = fmaf(v.x, t.ne, acc.x); acc.y = fmaf(v.y, t.ne, acc.y); acc.z = fmaf(v.z, t.ne, acc.z); acc.w = fmaf(v.w, t.ne, acc.w);
    }
    if (t.y1ok) {
        const float4 v = __ldg(base + dy);
        acc.x = fmaf(v.x, t.sw, acc.x); acc.y = fmaf(v.y, t.sw, acc.y); acc.z = fmaf(v.z, t.sw, acc.z); acc.w = fmaf(v.w, t.sw, acc.w);
    }
    if (t.x1ok && t.y1ok) {
        const float4 v = __ldg(base + dy + dx);
        acc.x = fmaf(v.x, t.se, acc.x); acc.y = fmaf(v.y, t.se, acc.y); acc.z = fmaf(v.z, t.se, acc.z); acc.w = fmaf(v.w, t.se, acc.w);
    }
    if (HALF) reinterpret_cast<uint2*>(feat_)[q4] = pack4h(acc);
    else reinterpret_cast<float4*>(feat_)[q4] = acc;
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(v.x * w), "f"(v.y * w), "f"(v.z * w),
                 "f"(v.w * w)
                 : "memory");
}

template <bool HALF>
__global__ void __launch_bounds__(256)
k_sample_bwd(const void* __restrict__ g_feat_, const float* __restrict__ xyz, uint32_t M, int R, int C, float inv_bound,
             int fp16_coords, const int32_t* __restrict__ n_valid, const int32_t* __restrict__ perm,
             float* __restrict__ g_planes) {
    const int cq_per = C >> 2;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)M * 3 * cq_per;
    if (idx >= total) return;
    const int cq = (int)(idx % cq_per);
    const int p = (int)((idx / cq_per) % 3);
    uint32_t m = (uint32_t)(idx / ((uint64_t)3 * cq_per));
    if (perm) m = (uint32_t)__ldg(perm + m);
    if (n_valid && (int32_t)m >= *n_valid) return;
    const size_t q4 = ((size_t)m * 3 * C + (size_t)p * C) / 4 + cq;
    const float4 g = HALF ? unpack4h(__ldg(reinterpret_cast<const uint2*>(g_feat_) + q4))
                          : __ldg(reinterpret_cast<const float4*>(g_feat_) + q4);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;  // adding zeros is a no-op
    float gx, gy;
    plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
    const Tap t = make_tap(gx, gy, R);
    float* base = g_planes + (((size_t)p * R + t.y0) * R + t.x0) * C + 4 * cq;
    const size_t dx = (size_t)C, dy = (size_t)R * C;
    red_add_v4(base, g, t.nw);
    if (t.x1ok) red_add_v4(base + dx, g, t.ne);
    if (t.y1ok) red_add_v4(base + dy, g, t.sw);
    if (t.x1ok && t.y1ok) red_add_v4(base + dy + dx, g, t.se);
}

// ------------------------------------------------------------------------------------------------
// Specialised variants for C in {16, 32, 48}: 8 channels per thread (two 128-bit loads per corner), thread -> (point,
// plane, channel group) decoded with compile-time constants (the generic kernels above spend most of their issue slots
// on 64-bit index arithmetic: ncu showed them issue-bound at 20-25 % of DRAM bandwidth once the visiting order is sorted).
// ------------------------------------------------------------------------------------------------
template <int TPP>  // threads per (point, plane) = C / 8
struct SampleCfg {
    static constexpr int C = 8 * TPP;
    static constexpr int PPB = 16;                     // points per block
    static constexpr int NT = PPB * 3 * TPP;           // threads per block
};

__device__ __forceinline__ void fma4(float4& acc, const float4 v, const float w) {
    acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
}

template <int TPP, bool HALF>
__global__ void __launch_bounds__(SampleCfg<TPP>::NT)
k_sample_fwd8(const float* __restrict__ planes, const float* __restrict__ xyz, uint32_t M, int R, float inv_bound,
              int fp16_coords, const int32_t* __restrict__ n_valid, const int32_t* __restrict__ perm,
              void* __restrict__ feat_) {
    using Cfg = SampleCfg<TPP>;
    constexpr int C = Cfg::C;
    const int tid = threadIdx.x;
    const int cg = tid % TPP, p = (tid / TPP) % 3, lp = tid / (3 * TPP);
    uint32_t m = blockIdx.x * Cfg::PPB + lp;
    if (m >= M) return;
    if (perm) m = (uint32_t)__ldg(perm + m);
    const size_t q8 = ((size_t)m * 3 + p) * TPP + cg;   // index of this thread's 8-channel group
    if (n_valid && (int32_t)m >= *n_valid) {
        if (HALF) reinterpret_cast<uint4*>(feat_)[q8] = make_uint4(0u, 0u, 0u, 0u);
        else { reinterpret_cast<float4*>(feat_)[2 * q8] = make_float4(0.f, 0.f, 0.f, 0.f); reinterpret_cast<float4*>(feat_)[2 * q8 + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        return;
    }
    float gx, gy;
    plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
    const Tap t = make_tap(gx, gy, R);
    const float4* base = reinterpret_cast<const float4*>(planes + (((size_t)p * R + t.y0) * R + t.x0) * C) + 2 * cg;
    const size_t dx = C / 4, dy = (size_t)R * (C / 4);
    float4 a0, a1;
    {
        const float4 v0 = __ldg(base), v1 = __ldg(base + 1);
        a0 = make_float4(v0.x * t.nw, v0.y * t.nw, v0.z * t.nw, v0.w * t.nw);
        a1 = make_float4(v1.x * t.nw, v1.y * t.nw, v1.z * t.nw, v1.w * t.nw);
    }
    if (t.x1ok) { fma4(a0, __ldg(base + dx), t.ne); fma4(a1, __ldg(base + dx + 1), t.ne); }
    if (t.y1ok) { fma4(a0, __ldg(base + dy), t.sw); fma4(a1, __ldg(base + dy + 1), t.sw); }
    if (t.x1ok && t.y1ok) { fma4(a0, __ldg(base + dy + dx), t.se); fma4(a1, __ldg(base + dy + dx + 1), t.se); }
    if (HALF) {
        const uint2 lo = pack4h(a0), hi = pack4h(a1);
        reinterpret_cast<uint4*>(feat_)[q8] = make_uint4(lo.x, lo.y, hi.x, hi.y);
    } else {
        reinterpret_cast<float4*>(feat_)[2 * q8] = a0;
        reinterpret_cast<float4*>(feat_)[2 * q8 + 1] = a1;
    }
}

// backward: 4 channels per thread (one vector atomic per corner) -- measured faster than 8 (more atomics in flight)
template <int TPP4>  // threads per (point, plane) = C / 4
struct ScatterCfg {
    static constexpr int C = 4 * TPP4;
    static constexpr int PPB = 8;
    static constexpr int NT = PPB * 3 * TPP4;
};

template <int TPP4, bool HALF>
__global__ void __launch_bounds__(ScatterCfg<TPP4>::NT)
k_sample_bwd4(const void* __restrict__ g_feat_, const float* __restrict__ xyz, uint32_t M, int R, float inv_bound,
              int fp16_coords, const int32_t* __restrict__ n_valid, const int32_t* __restrict__ perm,
              float* __restrict__ g_planes) {
    using Cfg = ScatterCfg<TPP4>;
    constexpr int C = Cfg::C;
    const int tid = threadIdx.x;
    const int cq = tid % TPP4, p = (tid / TPP4) % 3, lp = tid / (3 * TPP4);
    uint32_t m = blockIdx.x * Cfg::PPB + lp;
    if (m >= M) return;
    if (perm) m = (uint32_t)__ldg(perm + m);
    if (n_valid && (int32_t)m >= *n_valid) return;
    const size_t q4 = ((size_t)m * 3 + p) * TPP4 + cq;
    const float4 g = HALF ? unpack4h(__ldg(reinterpret_cast<const uint2*>(g_feat_) + q4))
                          : __ldg(reinterpret_cast<const float4*>(g_feat_) + q4);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;  // adding zeros is a no-op
    float gx, gy;
    plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
    const Tap t = make_tap(gx, gy, R);
    float* base = g_planes + (((size_t)p * R + t.y0) * R + t.x0) * C + 4 * cq;
    const size_t dx = (size_t)C, dy = (size_t)R * C;
    red_add_v4(base, g, t.nw);
    if (t.x1ok) red_add_v4(base + dx, g, t.ne);
    if (t.y1ok) red_add_v4(base + dy, g, t.sw);
    if (t.x1ok && t.y1ok) red_add_v4(base + dy + dx, g, t.se);
}

// the same scatter restricted to ONE plane (thread <-> (point, 4 channels); 24 points per block): the multi-GPU step scatters plane
// by plane so that the exchange of plane p overlaps the scatter of plane p + 1 (parallel.PeerGradExchange)
template <int TPP4, bool HALF>
__global__ void __launch_bounds__(ScatterCfg<TPP4>::NT)
k_sample_bwd4_plane(const void* __restrict__ g_feat_, const float* __restrict__ xyz, uint32_t M, int R, float inv_bound,
                    int fp16_coords, const int32_t* __restrict__ n_valid, const int32_t* __restrict__ perm,
                    float* __restrict__ g_planes, int p) {
    using Cfg = ScatterCfg<TPP4>;
    constexpr int C = Cfg::C;
    const int tid = threadIdx.x;
    const int cq = tid % TPP4, lp = tid / TPP4;
    uint32_t m = blockIdx.x * (3 * Cfg::PPB) + lp;
    if (m >= M) return;
    if (perm) m = (uint32_t)__ldg(perm + m);
    if (n_valid && (int32_t)m >= *n_valid) return;
    const size_t q4 = ((size_t)m * 3 + p) * TPP4 + cq;
    const float4 g = HALF ? unpack4h(__ldg(reinterpret_cast<const uint2*>(g_feat_) + q4))
                          : __ldg(reinterpret_cast<const float4*>(g_feat_) + q4);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
    float gx, gy;
    plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
    const Tap t = make_tap(gx, gy, R);
    float* base = g_planes + (((size_t)p * R + t.y0) * R + t.x0) * C + 4 * cq;
    const size_t dx = (size_t)C, dy = (size_t)R * C;
    red_add_v4(base, g, t.nw);
    if (t.x1ok) red_add_v4(base + dx, g, t.ne);
    if (t.y1ok) red_add_v4(base + dy, g, t.sw);
    if (t.x1ok && t.y1ok) red_add_v4(base + dy + dx, g, t.se);
}


// ------------------------------------------------------------------------------------------------
// Gradient with respect to the sample POSITIONS: grid_sampler_2d_backward w.r.t. the grid, chained through the projection
// u = xyz * inv_bound.  Callers that differentiate the field in space need it (the super_resolution application's analytic
// normals, super_resolution/threestudio/models/geometry/implicit_volume.py:218-226, through the F.grid_sample call of
// super_resolution/threestudio/models/triplaneencoder/triplane_encoder.py:262); the NeRF training step never does.
// Block = 32 points x 3 planes.  Thread (point, plane) walks the C channels of its four corner texels (contiguous 4*C-byte
// runs, every sector fully used) with the term order of ATen's kernel; the three planes of a point meet in shared memory
// and thread (point, axis) writes one component, so the result is deterministic and needs no atomics.
// ------------------------------------------------------------------------------------------------
constexpr int XG_PPB = 32;

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

__global__ void __launch_bounds__(3 * XG_PPB)
k_sample_xyz_grad(const float* __restrict__ g_feat, const float* __restrict__ planes, const float* __restrict__ xyz,
                  uint32_t M, int R, int C, float inv_bound, int fp16_coords, float* __restrict__ g_xyz) {
    __shared__ float s_uv[XG_PPB][3][2];
    const int p = threadIdx.x % 3, lp = threadIdx.x / 3;
    const uint32_t m = blockIdx.x * XG_PPB + lp;
    float gu = 0.f, gv = 0.f;
    if (m < M) {
        float gx, gy, mx, my;
        plane_coords(xyz, m, p, inv_bound, fp16_coords, gx, gy);
        const float ix = to_pixel_grad(gx, R, mx), iy = to_pixel_grad(gy, R, my);
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const bool x1ok = x0 + 1 <= R - 1, y1ok = y0 + 1 <= R - 1;
        const float wx0 = (fx + 1.0f) - ix, wx1 = ix - fx, wy0 = (fy + 1.0f) - iy, wy1 = iy - fy;
        const int cq_per = C >> 2;
        const float4* base = reinterpret_cast<const float4*>(planes + (((size_t)p * R + y0) * R + x0) * C);
        const float4* go = reinterpret_cast<const float4*>(g_feat + ((size_t)m * 3 + p) * C);
        const size_t dx = (size_t)cq_per, dy = (size_t)R * cq_per;
        for (int q = 0; q < cq_per; ++q) {
            const float4 g = __ldg(go + q);
            const float nw = dot4(__ldg(base + q), g);
            const float ne = x1ok ? dot4(__ldg(base + dx + q), g) : 0.f;
            const float sw = y1ok ? dot4(__ldg(base + dy + q), g) : 0.f;
            const float se = (x1ok && y1ok) ? dot4(__ldg(base + dy + dx + q), g) : 0.f;
            gu -= nw * wy0; gv -= nw * wx0;
            gu += ne * wy0; gv -= ne * wx1;
            gu -= sw * wy1; gv += sw * wx0;
            gu += se * wy1; gv += se * wx1;
        }
        gu *= mx;
        gv *= my;
    }
    s_uv[lp][p][0] = gu;
    s_uv[lp][p][1] = gv;
    __syncthreads();
    if (m < M) {
        // axis a = p of this thread: x <- u of planes 0, 1;  y <- v of plane 1, u of plane 2;  z <- v of planes 0, 2
        const float s = (p == 0) ? s_uv[lp][0][0] + s_uv[lp][1][0]
                      : (p == 1) ? s_uv[lp][1][1] + s_uv[lp][2][0]
                                 : s_uv[lp][0][1] + s_uv[lp][2][1];
        g_xyz[3 * (size_t)m + p] = s * inv_bound;
    }
}

template <int TPP>
static void launch_fwd8(const float* planes, const float* xyz, uint32_t M, uint32_t R, float inv_bound, int fp16_coords,
                        const int32_t* n_valid, const int32_t* perm, void* feat, int half, cudaStream_t s) {
    using Cfg = SampleCfg<TPP>;
    const unsigned blocks = ceil_div(M, (uint32_t)Cfg::PPB);
    if (half) k_sample_fwd8<TPP, true><<<blocks, Cfg::NT, 0, s>>>(planes, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, feat);
    else k_sample_fwd8<TPP, false><<<blocks, Cfg::NT, 0, s>>>(planes, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, feat);
}
template <int TPP4>
static void launch_bwd4(const void* g_feat, int half, const float* xyz, uint32_t M, uint32_t R, float inv_bound, int fp16_coords,
                        const int32_t* n_valid, const int32_t* perm, float* g_planes, cudaStream_t s) {
    using Cfg = ScatterCfg<TPP4>;
    const unsigned blocks = ceil_div(M, (uint32_t)Cfg::PPB);
    if (half) k_sample_bwd4<TPP4, true><<<blocks, Cfg::NT, 0, s>>>(g_feat, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, g_planes);
    else k_sample_bwd4<TPP4, false><<<blocks, Cfg::NT, 0, s>>>(g_feat, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, g_planes);
}

template <int TPP4>
static void launch_bwd4_plane(const void* g_feat, int half, const float* xyz, uint32_t M, uint32_t R, float inv_bound, int fp16_coords,
                              const int32_t* n_valid, const int32_t* perm, float* g_planes, int plane, cudaStream_t s) {
    using Cfg = ScatterCfg<TPP4>;
    const unsigned blocks = ceil_div(M, (uint32_t)(3 * Cfg::PPB));
    if (half) k_sample_bwd4_plane<TPP4, true><<<blocks, Cfg::NT, 0, s>>>(g_feat, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, g_planes, plane);
    else k_sample_bwd4_plane<TPP4, false><<<blocks, Cfg::NT, 0, s>>>(g_feat, xyz, M, (int)R, inv_bound, fp16_coords, n_valid, perm, g_planes, plane);
}

}  // namespace tnl

using namespace tnl;

extern "C" {

int tnl_sample_planes_forward(const float* planes, const float* xyz, uint32_t M, uint32_t R, uint32_t C, float inv_bound,
                              int fp16_coords, const int32_t* n_valid, const int32_t* perm, void* feat, int feat_fp16,
                              tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(planes && xyz && feat, "null pointer");
    TNL_ARG_CHECK(C >= 4 && C % 4 == 0 && R >= 2, "C must be a multiple of 4, R >= 2");
    TNL_ARG_CHECK(((uintptr_t)planes & 15) == 0 && ((uintptr_t)feat & 15) == 0, "planes/feat must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C == 16 || C == 32 || C == 48) {
        if (C == 16) launch_fwd8<2>(planes, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, feat, feat_fp16, st);
        else if (C == 32) launch_fwd8<4>(planes, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, feat, feat_fp16, st);
        else launch_fwd8<6>(planes, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, feat, feat_fp16, st);
        return finish_launch("sample_planes_forward");
    }
    const uint64_t total = (uint64_t)M * 3 * (C / 4);
    if (feat_fp16)
        k_sample_fwd<true><<<(unsigned)ceil_div(total, (uint64_t)256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            planes, xyz, M, (int)R, (int)C, inv_bound, fp16_coords, n_valid, perm, feat);
    else
        k_sample_fwd<false><<<(unsigned)ceil_div(total, (uint64_t)256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            planes, xyz, M, (int)R, (int)C, inv_bound, fp16_coords, n_valid, perm, feat);
    return finish_launch("sample_planes_forward");
}

int tnl_sample_planes_backward(const void* g_feat, int feat_fp16, const float* xyz, uint32_t M, uint32_t R, uint32_t C,
                               float inv_bound, int fp16_coords, const int32_t* n_valid, const int32_t* perm,
                               float* g_planes, tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(g_feat && xyz && g_planes, "null pointer");
    TNL_ARG_CHECK(C >= 4 && C % 4 == 0 && R >= 2, "C must be a multiple of 4, R >= 2");
    TNL_ARG_CHECK(((uintptr_t)g_planes & 15) == 0 && ((uintptr_t)g_feat & 15) == 0, "g_planes/g_feat must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C == 16 || C == 32 || C == 48) {
        if (C == 16) launch_bwd4<4>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, st);
        else if (C == 32) launch_bwd4<8>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, st);
        else launch_bwd4<12>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, st);
        return finish_launch("sample_planes_backward");
    }
    const uint64_t total = (uint64_t)M * 3 * (C / 4);
    if (feat_fp16)
        k_sample_bwd<true><<<(unsigned)ceil_div(total, (uint64_t)256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            g_feat, xyz, M, (int)R, (int)C, inv_bound, fp16_coords, n_valid, perm, g_planes);
    else
        k_sample_bwd<false><<<(unsigned)ceil_div(total, (uint64_t)256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            g_feat, xyz, M, (int)R, (int)C, inv_bound, fp16_coords, n_valid, perm, g_planes);
    return finish_launch("sample_planes_backward");
}

int tnl_sample_planes_backward_plane(const void* g_feat, int feat_fp16, const float* xyz, uint32_t M, uint32_t R, uint32_t C,
                                     float inv_bound, int fp16_coords, const int32_t* n_valid, const int32_t* perm,
                                     float* g_planes, uint32_t plane, tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(g_feat && xyz && g_planes && plane < 3, "null pointer / plane index");
    TNL_ARG_CHECK(C == 16 || C == 32 || C == 48, "per-plane scatter: C in {16, 32, 48}");
    TNL_ARG_CHECK(((uintptr_t)g_planes & 15) == 0 && ((uintptr_t)g_feat & 15) == 0, "g_planes/g_feat must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (C == 16) launch_bwd4_plane<4>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, (int)plane, st);
    else if (C == 32) launch_bwd4_plane<8>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, (int)plane, st);
    else launch_bwd4_plane<12>(g_feat, feat_fp16, xyz, M, R, inv_bound, fp16_coords, n_valid, perm, g_planes, (int)plane, st);
    return finish_launch("sample_planes_backward_plane");
}

int tnl_sample_planes_backward_coords(const float* g_feat, const float* planes, const float* xyz, uint32_t M, uint32_t R,
                                      uint32_t C, float inv_bound, int fp16_coords, float* g_xyz, tnl_stream_t stream) {
    if (M == 0) return 0;
    TNL_ARG_CHECK(g_feat && planes && xyz && g_xyz, "null pointer");
    TNL_ARG_CHECK(C >= 4 && C % 4 == 0 && R >= 2, "C must be a multiple of 4, R >= 2");
    TNL_ARG_CHECK(((uintptr_t)planes & 15) == 0 && ((uintptr_t)g_feat & 15) == 0, "planes/g_feat must be 16-byte aligned");
    k_sample_xyz_grad<<<ceil_div(M, (uint32_t)XG_PPB), 3 * XG_PPB, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        g_feat, planes, xyz, M, (int)R, (int)C, inv_bound, fp16_coords, g_xyz);
    return finish_launch("sample_planes_backward_coords");
}

}  // extern "C"
