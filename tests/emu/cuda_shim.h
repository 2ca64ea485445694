// TEST-ONLY: a small CUDA-on-host execution shim so that the CPU test-suite can run the PRODUCT kernels' source
// (trinerflet_b200/csrc/*.cu, rewritten only at their `<<<...>>>` launch sites by tests/emu/gen_kemu.py) without a GPU.
// Never part of libtrinerflet_b200.so; nothing in the product imports or links it.
//
// Execution model: the blocks of a launch run one after the other; the threads of a block are cooperative fibers
// (ucontext) on one OS thread, scheduled round-robin.  __syncthreads() and the warp collectives (__shfl_*_sync,
// __ballot_sync, __any_sync, __all_sync, __syncwarp) are rendezvous points between fibers, so kernels with block scans
// and warp scans execute with their real data flow.  Atomics are plain read-modify-writes (one OS thread).  fp32
// intrinsics map to the correctly rounded host operations (build with -ffp-contract=off); the approximate ones
// (__expf, ex2.approx) map to libm, which is the only intended numerical difference from the device.
// Not modelled: inline PTX (the generator substitutes the one `red` instruction it knows), tensor-core / TMA / cp.async
// instructions, memory-ordering hazards between unsynchronised threads.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <utility>
#include <vector>

#define TNL_KERNEL_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static const
#define __align__(n) alignas(n)

// ---------------------------------------------------------------------------------------------- vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---------------------------------------------------------------------------------------------- runtime API subset
typedef struct CUstream_st* cudaStream_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorUnknown = 999 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---------------------------------------------------------------------------------------------- fibers
namespace tnl_emu {

struct Warp {
    unsigned alive = 0;       // lanes that exist and have not returned
    unsigned arrived = 0;     // lanes waiting at the current warp rendezvous
    unsigned mask = 0;        // member mask of the current rendezvous (taken from its first arriver)
    unsigned gen = 0;
    uint64_t slot[32];
    uint32_t wide[32][8];     // per-lane operand registers of the emulated warp-wide PTX instructions (mma / ldmatrix)
};

struct Block {
    unsigned nthreads = 0, alive = 0, arrived = 0, gen = 0;
    std::vector<Warp> warps;
};

#if defined(__x86_64__)
#define TNL_EMU_ASM_SWITCH 1
// minimal cooperative context switch (callee-saved registers + stack pointer); swapcontext() costs a sigprocmask system
// call per switch, which dominates kernels that rendezvous often (block scans with 1024 threads)
extern "C" void tnl_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak tnl_emu_switch
.type tnl_emu_switch,@function
tnl_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size tnl_emu_switch,.-tnl_emu_switch
)");
#endif

struct Fiber {
#ifdef TNL_EMU_ASM_SWITCH
    void* sp = nullptr;
#else
    ucontext_t ctx;
#endif
    bool done = true;
    unsigned tid = 0;
};

struct Runtime {
#ifdef TNL_EMU_ASM_SWITCH
    void* main_sp = nullptr;
#else
    ucontext_t main_ctx;
#endif
    std::vector<Fiber> fibers;
    char* stacks = nullptr;
    size_t stacks_for = 0;
    Block block;
    const std::function<void()>* body = nullptr;
    Fiber* cur = nullptr;
    size_t stack_bytes = 64 * 1024;
    unsigned long long progress = 0;   // bumped by every arrival at a rendezvous and every thread exit
    std::vector<uint64_t> dyn_smem;    // dynamic shared memory of the running launch (16-byte aligned, see dyn_smem())
    std::vector<unsigned> order;       // resume order of the fibers of a block (TNL_EMU_SCHED)
};

inline Runtime& rt() {
    static Runtime r;
    return r;
}

}  // namespace tnl_emu

// the built-in index variables: plain globals, rewritten by the scheduler before a fiber resumes
inline uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);

namespace tnl_emu {

inline void set_thread(unsigned tid) {
    threadIdx.x = tid % blockDim.x;
    threadIdx.y = (tid / blockDim.x) % blockDim.y;
    threadIdx.z = tid / (blockDim.x * blockDim.y);
}

inline void to_main(Fiber* f) {
    Runtime& r = rt();
#ifdef TNL_EMU_ASM_SWITCH
    tnl_emu_switch(&f->sp, r.main_sp);
#else
    swapcontext(&f->ctx, &r.main_ctx);
#endif
}
inline void to_fiber(Fiber* f) {
    Runtime& r = rt();
#ifdef TNL_EMU_ASM_SWITCH
    tnl_emu_switch(&r.main_sp, f->sp);
#else
    swapcontext(&r.main_ctx, &f->ctx);
#endif
}

inline void yield() { to_main(rt().cur); }

inline void release_block_if_complete(Block& b) {
    if (b.alive > 0 && b.arrived == b.alive) {
        b.arrived = 0;
        ++b.gen;
    }
}
inline void release_warp_if_complete(Warp& w) {
    const unsigned expected = w.alive & w.mask;   // exited lanes count as arrived
    if (w.arrived != 0 && (w.arrived & expected) == expected) {
        w.arrived = 0;
        w.mask = 0;
        ++w.gen;
    }
}

inline void trampoline() {
    Runtime& r = rt();
    Fiber* f = r.cur;
    (*r.body)();
    // thread exit: it no longer takes part in any rendezvous (hardware counts exited threads as arrived)
    f->done = true;
    ++r.progress;
    Block& b = r.block;
    --b.alive;
    Warp& w = b.warps[f->tid / 32];
    w.alive &= ~(1u << (f->tid % 32));
    release_block_if_complete(b);
    release_warp_if_complete(w);
    to_main(f);
    abort();   // a finished fiber is never resumed
}

inline void run_block(unsigned nthreads, const std::function<void()>& body) {
    Runtime& r = rt();
    if (r.fibers.size() < nthreads) r.fibers.resize(nthreads);
    if (r.stacks_for < nthreads) {
        free(r.stacks);
        r.stacks = static_cast<char*>(malloc((size_t)nthreads * r.stack_bytes));
        r.stacks_for = nthreads;
    }
    r.body = &body;
    Block& b = r.block;
    b.nthreads = b.alive = nthreads;
    b.arrived = 0;
    b.gen = 0;
    b.warps.assign((nthreads + 31) / 32, Warp());
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = r.fibers[t];
        f.done = false;
        f.tid = t;
#ifdef TNL_EMU_ASM_SWITCH
        // initial frame: six zeroed callee-saved registers, then the entry point as the return address of the first switch;
        // after that `ret` the stack pointer is 8 (mod 16), as after a call
        uintptr_t top = reinterpret_cast<uintptr_t>(r.stacks + (size_t)(t + 1) * r.stack_bytes) & ~uintptr_t(15);
        void** sp = reinterpret_cast<void**>(top - 64);
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        sp[6] = reinterpret_cast<void*>(&trampoline);
        sp[7] = nullptr;
        f.sp = sp;
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = r.stacks + (size_t)t * r.stack_bytes;
        f.ctx.uc_stack.ss_size = r.stack_bytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
        b.warps[t / 32].alive |= 1u << (t % 32);
    }
    unsigned remaining = nthreads;
    unsigned idle_rounds = 0;
    // race check: TNL_EMU_SCHED=reverse | random[:seed] changes the order in which runnable threads are resumed.  A kernel
    // whose result depends on it (beyond the order of float atomics) is missing a barrier.
    static const int sched = [] {
        const char* e = getenv("TNL_EMU_SCHED");
        if (e == nullptr) return 0;
        if (strncmp(e, "reverse", 7) == 0) return 1;
        if (strncmp(e, "random", 6) == 0) { srand(e[6] == ':' ? (unsigned)atoi(e + 7) : 1u); return 2; }
        return 0;
    }();
    std::vector<unsigned>& order = r.order;
    order.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) order[t] = sched == 1 ? nthreads - 1 - t : t;
    while (remaining > 0) {
        const unsigned long long before = r.progress;
        if (sched == 2)
            for (unsigned t = nthreads; t > 1; --t) std::swap(order[t - 1], order[(unsigned)rand() % t]);
        for (unsigned k = 0; k < nthreads; ++k) {
            const unsigned t = order[k];
            Fiber& f = r.fibers[t];
            if (f.done) continue;
            r.cur = &f;
            set_thread(t);
            to_fiber(&f);
            if (f.done) --remaining;
        }
        // every fiber was resumed once; if none of them reached a new rendezvous or returned, the state cannot change
        // any more: the kernel has a divergent barrier / collective
        idle_rounds = (r.progress != before) ? 0 : idle_rounds + 1;
        if (idle_rounds > 1) {
            fprintf(stderr, "tnl_emu: deadlock (divergent __syncthreads / warp collective) in block (%u,%u,%u)\n", blockIdx.x,
                    blockIdx.y, blockIdx.z);
            abort();
        }
    }
    r.cur = nullptr;
}

struct LaunchCfg {
    dim3 grid, block;
    size_t smem;
};
inline LaunchCfg cfg(dim3 grid, dim3 block, size_t smem = 0, const void* stream = nullptr) {
    (void)stream;
    return LaunchCfg{grid, block, smem};
}
// `extern __shared__ T name[];` is rewritten by the generator into `T* name = (T*)tnl_emu::dyn_smem();`
inline void* dyn_smem() {
    Runtime& r = rt();
    uintptr_t a = reinterpret_cast<uintptr_t>(r.dyn_smem.data());
    return reinterpret_cast<void*>((a + 15) & ~uintptr_t(15));
}

inline unsigned long long g_launches = 0;

inline void launch(const LaunchCfg& c, const std::function<void()>& body) {
    ++g_launches;
    rt().dyn_smem.assign(c.smem / 8 + 4, 0xfff8dead0000beefull);   // NaN pattern: reads of unwritten shared memory show up
    gridDim = c.grid;
    blockDim = c.block;
    const unsigned nthreads = c.block.x * c.block.y * c.block.z;
    for (unsigned bz = 0; bz < c.grid.z; ++bz)
        for (unsigned by = 0; by < c.grid.y; ++by)
            for (unsigned bx = 0; bx < c.grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                run_block(nthreads, body);
            }
}

// ---- rendezvous primitives (called from inside fibers) ----
inline void block_barrier() {
    Runtime& r = rt();
    Block& b = r.block;
    const unsigned gen = b.gen;
    ++b.arrived;
    ++r.progress;
    release_block_if_complete(b);
    while (b.gen == gen) yield();
}

inline Warp& my_warp(unsigned& lane) {
    Runtime& r = rt();
    lane = r.cur->tid % 32;
    return r.block.warps[r.cur->tid / 32];
}

inline void warp_barrier(Warp& w, unsigned lane, unsigned mask) {
    const unsigned gen = w.gen;
    if (w.arrived == 0) w.mask = mask;
    w.arrived |= 1u << lane;
    ++rt().progress;
    release_warp_if_complete(w);
    while (w.gen == gen) yield();
}

template <typename T>
inline T warp_exchange(unsigned mask, T v, int src_lane_signed, bool src_valid) {
    static_assert(sizeof(T) <= 8, "warp shuffles move at most 64 bits");
    unsigned lane;
    Warp& w = my_warp(lane);
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[lane] = raw;
    warp_barrier(w, lane, mask);
    T out = v;
    if (src_valid && src_lane_signed >= 0 && src_lane_signed < 32) {
        const uint64_t got = w.slot[src_lane_signed];
        memcpy(&out, &got, sizeof(T));
    }
    warp_barrier(w, lane, mask);  // nobody overwrites a slot before everybody has read
    return out;
}

inline unsigned warp_vote(unsigned mask, bool pred) {
    unsigned lane;
    Warp& w = my_warp(lane);
    w.slot[lane] = pred ? 1 : 0;
    const unsigned participants = w.alive & mask;  // sampled before the rendezvous: lanes alive at the vote
    warp_barrier(w, lane, mask);
    unsigned bits = 0;
    for (unsigned l = 0; l < 32; ++l)
        if (((participants >> l) & 1u) && w.slot[l]) bits |= 1u << l;
    warp_barrier(w, lane, mask);
    return bits;
}

// ---- warp-wide PTX instructions of the mma.sync kernels (csrc/mlp.cu); substituted for the asm by gen_kemu.py ----
inline float half_bits_to_float(uint32_t v, int hi) {
    const uint16_t b = hi ? (uint16_t)(v >> 16) : (uint16_t)(v & 0xffffu);
    _Float16 h;
    memcpy(&h, &b, 2);
    return (float)h;
}

// mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32: D = A(16x16, row) * B(16x8, col) + D.  Fragment layout (g = lane / 4,
// t = lane % 4): a0 (g, 2t..2t+1)  a1 (g+8, 2t..)  a2 (g, 2t+8..)  a3 (g+8, 2t+8..);  b0 (k 2t..2t+1, n g)  b1 (k 2t+8.., n g);
// d0 (g, 2t)  d1 (g, 2t+1)  d2 (g+8, 2t)  d3 (g+8, 2t+1).  Products of fp16 values are exact in fp32; the sum is formed in
// fp64 and rounded once (the tensor core's internal accumulation order is not specified; tests carry fp16 tolerances).
inline void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    unsigned lane;
    Warp& w = my_warp(lane);
    for (int i = 0; i < 4; ++i) w.wide[lane][i] = a[i];
    w.wide[lane][4] = b0;
    w.wide[lane][5] = b1;
    warp_barrier(w, lane, 0xffffffffu);
    const int g = (int)lane >> 2, t = (int)lane & 3;
    float out[4];
    for (int i = 0; i < 4; ++i) {
        const int m = g + 8 * (i >> 1), n = 2 * t + (i & 1);
        double acc = (double)d[i];
        for (int k = 0; k < 16; ++k) {
            const float av = half_bits_to_float(w.wide[(m & 7) * 4 + ((k & 7) >> 1)][(m >> 3) + 2 * (k >> 3)], k & 1);
            const float bv = half_bits_to_float(w.wide[n * 4 + ((k & 7) >> 1)][4 + (k >> 3)], k & 1);
            acc += (double)av * (double)bv;
        }
        out[i] = (float)acc;
    }
    warp_barrier(w, lane, 0xffffffffu);
    for (int i = 0; i < 4; ++i) d[i] = out[i];
}

// ldmatrix.sync.aligned.m8n8.xN.trans.shared.b16: lane 8j+i supplies the address of row i of matrix j (8 halves);
// with .trans thread (g, t) receives r[j] = { M_j[2t][g], M_j[2t+1][g] }
template <int NMAT>
inline void ldmatrix_trans(uint32_t (&r)[NMAT], const void* row_ptr) {
    unsigned lane;
    Warp& w = my_warp(lane);
    w.slot[lane] = (uint64_t)reinterpret_cast<uintptr_t>(row_ptr);
    warp_barrier(w, lane, 0xffffffffu);
    const int g = (int)lane >> 2, t = (int)lane & 3;
    for (int j = 0; j < NMAT; ++j) {
        const uint16_t* r0 = reinterpret_cast<const uint16_t*>((uintptr_t)w.slot[8 * j + 2 * t]);
        const uint16_t* r1 = reinterpret_cast<const uint16_t*>((uintptr_t)w.slot[8 * j + 2 * t + 1]);
        r[j] = (uint32_t)r0[g] | ((uint32_t)r1[g] << 16);
    }
    warp_barrier(w, lane, 0xffffffffu);
}

inline unsigned lane_id() {
    unsigned lane;
    my_warp(lane);
    return lane;
}

}  // namespace tnl_emu

// ---------------------------------------------------------------------------------------------- device intrinsics
static inline void __syncthreads() { tnl_emu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    unsigned lane;
    tnl_emu::Warp& w = tnl_emu::my_warp(lane);
    tnl_emu::warp_barrier(w, lane, mask);
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = (int)tnl_emu::lane_id();
    const int s = (lane / width) * width + (src % width);
    return tnl_emu::warp_exchange(mask, v, s, true);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)tnl_emu::lane_id();
    const int s = lane - (int)delta;
    return tnl_emu::warp_exchange(mask, v, s, s >= (lane / width) * width);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)tnl_emu::lane_id();
    const int s = lane + (int)delta;
    return tnl_emu::warp_exchange(mask, v, s, s < (lane / width + 1) * width);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    const int lane = (int)tnl_emu::lane_id();
    const int s = lane ^ lane_mask;
    return tnl_emu::warp_exchange(mask, v, s, s / width == lane / width);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return tnl_emu::warp_vote(mask, pred != 0); }
static inline int __any_sync(unsigned mask, int pred) { return tnl_emu::warp_vote(mask, pred != 0) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return tnl_emu::warp_vote(mask, pred == 0) == 0; }

template <typename T>
static inline T __ldg(const T* p) { return *p; }

template <typename T>
static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
static inline unsigned atomicAdd(unsigned* p, int v) { unsigned old = *p; *p = old + (unsigned)v; return old; }
template <typename T>
static inline T atomicMax(T* p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <typename T>
static inline T atomicMin(T* p, T v) { T old = *p; if (v < old) *p = v; return old; }
template <typename T>
static inline T atomicOr(T* p, T v) { T old = *p; *p = old | v; return old; }
template <typename T>
static inline T atomicExch(T* p, T v) { T old = *p; *p = v; return old; }

// correctly rounded single operations (the translation unit is compiled with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// glibc's <math.h> declares __expf / __logf itself, so these two are macros.  Device: ex2.approx / lg2.approx based (2 ulp).
static inline float tnl_emu_expf(float a) { return expf(a); }
static inline float tnl_emu_logf(float a) { return logf(a); }
#define __expf(x) tnl_emu_expf(x)
#define __logf(x) tnl_emu_logf(x)
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline int __float2int_rz(float a) { return (int)a; }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
static inline int __float2int_rn(float a) { return (int)nearbyintf(a); }
static inline unsigned __float2uint_rz(float a) { return (unsigned)a; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------- fp16 / bf16
struct __half {
    _Float16 v;
};
typedef __half half;
struct alignas(4) __half2 {
    __half x, y;
};
typedef __half2 half2;
static inline __half __float2half_rn(float f) { __half h; h.v = (_Float16)f; return h; }
static inline __half __float2half(float f) { return __float2half_rn(f); }
static inline float __half2float(__half h) { return (float)h.v; }
static inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __float2half_rn(a); r.y = __float2half_rn(b); return r; }
static inline float2 __half22float2(__half2 h) { return float2{__half2float(h.x), __half2float(h.y)}; }

static inline __half2 __float2half2_rn(float a) { return __floats2half2_rn(a, a); }
static inline __half2 __hmax2(__half2 a, __half2 b) {
    __half2 r;
    r.x.v = a.x.v > b.x.v ? a.x.v : b.x.v;
    r.y.v = a.y.v > b.y.v ? a.y.v : b.y.v;
    return r;
}
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

struct __nv_bfloat16 {
    uint16_t bits;
};
struct alignas(4) __nv_bfloat162 {
    __nv_bfloat16 x, y;
};
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = 0x7fff; return r; }            // NaN
    u += 0x7fffu + ((u >> 16) & 1u);                                                // round to nearest even
    r.bits = (uint16_t)(u >> 16);
    return r;
}
static inline float __bfloat162float(__nv_bfloat16 b) {
    const uint32_t u = (uint32_t)b.bits << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
    __nv_bfloat162 r;
    r.x = __float2bfloat16_rn(a);
    r.y = __float2bfloat16_rn(b);
    return r;
}
static inline float2 __bfloat1622float2(__nv_bfloat162 v) { return float2{__bfloat162float(v.x), __bfloat162float(v.y)}; }
