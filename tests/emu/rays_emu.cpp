// TEST-ONLY host emulator of the step-feeder kernel: runs the per-ray functions of
// trinerflet_b200/csrc/rays_core.cuh for every "thread", so the CPU test-suite can check the kernel's arithmetic against
// the oracle and the reference-generated golden vectors without a GPU.  Never linked into the product library.
// Build: g++ -O2 -ffp-contract=off (one rounding per operation, as the __f*_rn intrinsics of the device build).
#include <cstdint>
#include "../../trinerflet_b200/csrc/rays_core.cuh"

using namespace tnl;

extern "C" void emu_rays_from_ids(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                                  const int64_t* ids, int64_t first_id, uint32_t n, const float* images, uint32_t ci,
                                  float* rays_o, float* rays_d, float* gt) {
    const uint32_t HW = H * W;
    const int64_t n_total = (int64_t)B * HW;
    const RayIntrinsics k{fx, fy, cx, cy};
    for (uint32_t t = 0; t < n; ++t) {
        int64_t id = ids ? ids[t] : first_id + (int64_t)t;
        id = id < 0 ? 0 : (id >= n_total ? n_total - 1 : id);
        uint32_t img, pix;
        split_ray_id(id, HW, img, pix);
        float dir[3];
        pixel_direction(k, W, pix, dir);
        ray_from_pose(poses + 16 * (size_t)img, dir, rays_o + 3 * (size_t)t, rays_d + 3 * (size_t)t);
        for (uint32_t c = 0; c < ci; ++c) gt[(size_t)t * ci + c] = images[(size_t)id * ci + c];
    }
}
