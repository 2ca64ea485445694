// Core of the on-device step feeder: per-ray arithmetic of the reference's get_rays
// (/root/reference/reconstruction/nerf/utils.py:64-149) for an explicit (image, pixel) pair, written so that the SAME
// code compiles for the device (rays.cu) and for the host-side emulator of the CPU test-suite
// (tests/emu/rays_emu.cpp).  The product path is the CUDA kernel; the emulator is never linked into the library.
//
// Reference op sequence (utils.py:80-82, 136-145), all fp32, one rounding per op:
//     i = (pix % W) + 0.5 ; j = (pix / W) + 0.5                 (meshgrid 'ij' of linspace, transposed, flattened)
//     xs = (i - cx) / fx * 1 ; ys = (j - cy) / fy * 1 ; zs = 1
//     dir = (xs, ys, zs) / sqrt(xs^2 + ys^2 + zs^2)
//     rays_d = dir @ pose[:3,:3]^T ; rays_o = pose[:3,3]
// The two reductions follow the accumulation ATen's CPU kernels perform (fused multiply-adds, operands in index order):
//     |.|^2 = fma(zs, zs, fma(ys, ys, xs*xs)) ;  rays_d[r] = fma(dir2, R[r][2], fma(dir1, R[r][1], dir0*R[r][0]))
// which makes the result bit-identical to the reference-generated golden vectors (tests/golden/rays_ref.npz).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TNL_RHD __host__ __device__ __forceinline__
#else
#define TNL_RHD inline
#endif

namespace tnl {

// single-rounding fp32 operations (no FMA contraction on either side; the host build uses -ffp-contract=off)
#ifdef __CUDA_ARCH__
TNL_RHD float r_add(float a, float b) { return __fadd_rn(a, b); }
TNL_RHD float r_sub(float a, float b) { return __fsub_rn(a, b); }
TNL_RHD float r_mul(float a, float b) { return __fmul_rn(a, b); }
TNL_RHD float r_div(float a, float b) { return __fdiv_rn(a, b); }
TNL_RHD float r_sqrt(float a) { return __fsqrt_rn(a); }
TNL_RHD float r_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
TNL_RHD float r_add(float a, float b) { return a + b; }
TNL_RHD float r_sub(float a, float b) { return a - b; }
TNL_RHD float r_mul(float a, float b) { return a * b; }
TNL_RHD float r_div(float a, float b) { return a / b; }
TNL_RHD float r_sqrt(float a) { return sqrtf(a); }
TNL_RHD float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

struct RayIntrinsics {
    float fx, fy, cx, cy;
};

// camera-space unit direction of the centre of pixel `pix` (row-major, pix = row * W + col)
TNL_RHD void pixel_direction(const RayIntrinsics& k, uint32_t W, uint32_t pix, float dir[3]) {
    const float i = r_add((float)(pix % W), 0.5f);
    const float j = r_add((float)(pix / W), 0.5f);
    const float xs = r_div(r_sub(i, k.cx), k.fx);
    const float ys = r_div(r_sub(j, k.cy), k.fy);
    const float nrm = r_sqrt(r_fma(1.0f, 1.0f, r_fma(ys, ys, r_mul(xs, xs))));
    dir[0] = r_div(xs, nrm);
    dir[1] = r_div(ys, nrm);
    dir[2] = r_div(1.0f, nrm);
}

// pose: 16 floats, row-major cam2world [4][4] -> world-space origin and direction
TNL_RHD void ray_from_pose(const float* pose, const float dir[3], float o[3], float d[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = r_fma(dir[2], pose[4 * r + 2], r_fma(dir[1], pose[4 * r + 1], r_mul(dir[0], pose[4 * r + 0])));
        o[r] = pose[4 * r + 3];
    }
}

// flat ray id (shuffle_data's index into the [B*H*W] view, utils.py:228-236) -> (image, pixel)
TNL_RHD void split_ray_id(int64_t id, uint32_t HW, uint32_t& img, uint32_t& pix) {
    img = (uint32_t)(id / (int64_t)HW);
    pix = (uint32_t)(id - (int64_t)img * (int64_t)HW);
}

}  // namespace tnl
