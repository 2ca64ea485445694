// Epilogue helpers shared by the tcgen05 MLP kernels (mlp_tc.cu: 64-wide heads, mlp_tc128.cu: 128-wide heads).
#pragma once
#include "mlp_math.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

namespace tnl {

// 8 fp32 -> 8 fp16 (one 16-byte tile row chunk), optional ReLU on the rounded values
template <bool RELU>
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 o;
    o.x = pack_h2(v[0], v[1]);
    o.y = pack_h2(v[2], v[3]);
    o.z = pack_h2(v[4], v[5]);
    o.w = pack_h2(v[6], v[7]);
    if (RELU) { o.x = relu_h2(o.x); o.y = relu_h2(o.y); o.z = relu_h2(o.z); o.w = relu_h2(o.w); }
    return o;
}
// zero the fp16 lanes of `d` whose counterpart in `h` (a ReLU output, never negative) is zero
__device__ __forceinline__ uint32_t mask_h2(uint32_t d, uint32_t h) {
    const uint32_t m = (((h & 0x7fffu) != 0u) ? 0x0000ffffu : 0u) | (((h & 0x7fff0000u) != 0u) ? 0xffff0000u : 0u);
    return d & m;
}
__device__ __forceinline__ uint4 mask8(uint4 d, uint4 h) {
    return make_uint4(mask_h2(d.x, h.x), mask_h2(d.y, h.y), mask_h2(d.z, h.z), mask_h2(d.w, h.w));
}
__device__ __forceinline__ uint32_t clamp_valid(const int32_t* n_valid_ptr, uint32_t M) {
    if (!n_valid_ptr) return M;
    const int32_t nv = *n_valid_ptr;
    return nv < 0 ? 0u : ((uint32_t)nv < M ? (uint32_t)nv : M);
}

// 16-byte asynchronous global -> shared copy (LDGSTS) and its completion wait; generic-proxy writes: a
// fence.proxy.async must follow before a tensor-core operand read
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace tnl
