// On-device step feeder -- sm_100a.  Replaces, for the per-step batch, the reference's
//   get_rays                       /root/reference/reconstruction/nerf/utils.py:64-149
//   shuffle_data / select_batch    nerf/utils.py:228-243   (index into the flattened [B*H*W] ray table)
//   the target gather of collate   nerf/provider.py:708-711
// The reference materialises rays_o / rays_d / images for ALL B*H*W rays (64 M rays = 1.5 GB + 0.77 GB for the Blender
// set), permutes the three tables on the host every epoch and copies a pageable slice to the device every step.  Here
// the poses (6.4 KB) and the images stay resident in HBM and a step's rays are generated from their flat ids.
#include "common.cuh"
#include "rays_core.cuh"

namespace tnl {

// thread <-> ray.  ids == nullptr: id = first_id + thread index (full frames, get_rays with N = -1).
template <int CI>
__global__ void k_rays_from_ids(const float* __restrict__ poses, RayIntrinsics k, uint32_t HW, uint32_t W,
                                const int64_t* __restrict__ ids, int64_t first_id, int64_t n_total, uint32_t n,
                                const float* __restrict__ images, float* __restrict__ rays_o, float* __restrict__ rays_d,
                                float* __restrict__ gt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t id = ids ? ids[t] : first_id + (int64_t)t;
    // out-of-range ids are clamped (the reference would raise an index error on the host; a kernel cannot)
    id = id < 0 ? 0 : (id >= n_total ? n_total - 1 : id);
    uint32_t img, pix;
    split_ray_id(id, HW, img, pix);
    float dir[3], o[3], d[3];
    pixel_direction(k, W, pix, dir);
    const float4* P4 = reinterpret_cast<const float4*>(poses + 16 * (size_t)img);
    const float4 r0 = __ldg(P4), r1 = __ldg(P4 + 1), r2 = __ldg(P4 + 2);
    const float pose[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
    ray_from_pose(pose, dir, o, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_o[3 * (size_t)t + a] = o[a];
        rays_d[3 * (size_t)t + a] = d[a];
    }
    if (CI > 0) {
        const float* src = images + (size_t)id * CI;
#pragma unroll
        for (int c = 0; c < CI; ++c) gt[(size_t)t * CI + c] = __ldg(src + c);
    }
}

}  // namespace tnl

using namespace tnl;

extern "C" {

int tnl_rays_from_ids(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                      const int64_t* ray_ids, int64_t first_id, uint32_t n, const float* images, uint32_t image_channels,
                      float* rays_o, float* rays_d, float* targets, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(poses && rays_o && rays_d, "null pointer");
    TNL_ARG_CHECK(B >= 1 && H >= 1 && W >= 1 && (uint64_t)H * W <= 0x7fffffffull, "bad image geometry");
    TNL_ARG_CHECK(fx != 0.f && fy != 0.f, "zero focal length");
    TNL_ARG_CHECK((images == nullptr) == (targets == nullptr), "images and targets must be given together");
    TNL_ARG_CHECK(images == nullptr || image_channels == 3 || image_channels == 4, "image_channels must be 3 or 4");
    TNL_ARG_CHECK((reinterpret_cast<uintptr_t>(poses) & 15) == 0, "poses must be 16-byte aligned");
    const uint32_t HW = H * W;
    const int64_t n_total = (int64_t)B * HW;
    TNL_ARG_CHECK(ray_ids || (first_id >= 0 && first_id + (int64_t)n <= n_total), "implicit id range outside the ray table");
    const RayIntrinsics k{fx, fy, cx, cy};
    const dim3 grid(ceil_div(n, 256u)), block(256);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t ci = images ? image_channels : 0;
    if (ci == 0)
        k_rays_from_ids<0><<<grid, block, 0, s>>>(poses, k, HW, W, ray_ids, first_id, n_total, n, images, rays_o, rays_d, targets);
    else if (ci == 3)
        k_rays_from_ids<3><<<grid, block, 0, s>>>(poses, k, HW, W, ray_ids, first_id, n_total, n, images, rays_o, rays_d, targets);
    else
        k_rays_from_ids<4><<<grid, block, 0, s>>>(poses, k, HW, W, ray_ids, first_id, n_total, n, images, rays_o, rays_d, targets);
    return finish_launch("rays_from_ids");
}

}  // extern "C"
