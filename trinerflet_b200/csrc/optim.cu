// Optimizer epilogue of a training step -- sm_100a.
//
// Behavioural contract: `scaler.step(optimizer); scaler.update()` of the reference loop
// (/root/reference/reconstruction/nerf/utils.py:1170-1173) with `torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15)`
// (/root/reference/reconstruction/main_nerf.py:119): unscale the gradients by 1/scale, skip the whole step if any
// gradient is non-finite, otherwise the Adam update of torch's `_single_tensor_adam` (lerp form of the first moment,
// denom = sqrt(v)/sqrt(bias_correction2) + eps, step_size = lr / bias_correction1).  The library path makes ~9 passes
// over the 1.6 GB of coefficients (check + unscale r/w, then the foreach Adam chain); here it is one read pass for the
// check and one pass that reads p, g, m, v and writes p, m, v.  HBM-bound streaming kernels, float4 per thread.
#include "common.cuh"

namespace tnl {

__global__ void __launch_bounds__(256)
k_grad_nonfinite(const float* __restrict__ g, size_t n, float* __restrict__ found_inf) {
    const size_t n4 = n / 4;
    bool bad = false;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !isfinite(g[i]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;   // same value from every writer
}

// state = {step, bias_correction1, sqrt(bias_correction2)}; the step only advances when the update is applied
__global__ void k_adam_prepare(float* __restrict__ state, const float* __restrict__ found_inf, float beta1, float beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (found_inf != nullptr && *found_inf != 0.0f) return;
    const float step = state[0] + 1.0f;
    state[0] = step;
    state[1] = (float)(1.0 - pow((double)beta1, (double)step));
    state[2] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float inv_scale, float lr_over_bc1, float beta1, float beta2,
                                      float eps, float wd, float bc2s) {
    g *= inv_scale;
    if (wd != 0.0f) g = fmaf(wd, p, g);
    m = fmaf(g - m, 1.0f - beta1, m);
    v = fmaf(v, beta2, (1.0f - beta2) * g * g);
    const float denom = sqrtf(v) / bc2s + eps;
    p = p - lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
       const float* __restrict__ inv_scale_ptr, const float* __restrict__ found_inf, const float* __restrict__ state, float lr,
       float beta1, float beta2, float eps, float wd) {
    if (found_inf != nullptr && *found_inf != 0.0f) return;   // GradScaler semantics: a non-finite gradient skips the step
    const float inv_scale = inv_scale_ptr ? __ldg(inv_scale_ptr) : 1.0f;
    const float lr_over_bc1 = lr / __ldg(state + 1), bc2s = __ldg(state + 2);
    const size_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        adam1(pp.x, gg.x, mm.x, vv.x, inv_scale, lr_over_bc1, beta1, beta2, eps, wd, bc2s);
        adam1(pp.y, gg.y, mm.y, vv.y, inv_scale, lr_over_bc1, beta1, beta2, eps, wd, bc2s);
        adam1(pp.z, gg.z, mm.z, vv.z, inv_scale, lr_over_bc1, beta1, beta2, eps, wd, bc2s);
        adam1(pp.w, gg.w, mm.w, vv.w, inv_scale, lr_over_bc1, beta1, beta2, eps, wd, bc2s);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam1(pp, g[i], mm, vv, inv_scale, lr_over_bc1, beta1, beta2, eps, wd, bc2s);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

static inline uint32_t stream_grid(size_t n) {
    const size_t want = (n / 4 + 255) / 256;
    const size_t cap = (size_t)kNumSM * 16;
    return (uint32_t)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace tnl

using namespace tnl;

extern "C" {

int tnl_grad_nonfinite(const float* g, uint64_t n, float* found_inf, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(g && found_inf, "null pointer");
    TNL_ARG_CHECK(((uintptr_t)g & 15) == 0, "gradient must be 16-byte aligned");
    k_grad_nonfinite<<<stream_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, (size_t)n, found_inf);
    return finish_launch("grad_nonfinite");
}

int tnl_adam_prepare(float* state3, const float* found_inf, float beta1, float beta2, tnl_stream_t stream) {
    TNL_ARG_CHECK(state3, "null pointer");
    k_adam_prepare<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state3, found_inf, beta1, beta2);
    return finish_launch("adam_prepare");
}

int tnl_adam_step(float* p, const float* g, float* m, float* v, uint64_t n, const float* inv_scale, const float* found_inf,
                  const float* state3, float lr, float beta1, float beta2, float eps, float weight_decay, tnl_stream_t stream) {
    if (n == 0) return 0;
    TNL_ARG_CHECK(p && g && m && v && state3, "null pointer");
    TNL_ARG_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "tensors must be 16-byte aligned");
    k_adam<<<stream_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, (size_t)n, inv_scale, found_inf, state3, lr, beta1,
                                                                              beta2, eps, weight_decay);
    return finish_launch("adam_step");
}

}  // extern "C"
