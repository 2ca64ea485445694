// Core of the IDWT level kernels: per-thread phase functions, written so that the SAME code compiles
// for the device (idwt.cu) and for a host-side lock-step emulator used only by the CPU test-suite
// (tests/emu/idwt_emu.cpp) to check the index arithmetic without a GPU.  The product path is the
// CUDA kernel; the emulator is never linked into libtrinerflet_b200.so.
//
// Behavioural contract: one iteration of the level loop of TriPlaneVolume.build_planes
// (/root/reference/reconstruction/triplaneencoder/triplane_encoder.py:379-394):
//     yl = 2*x ; pad yl, yh by 4 ; x' = DWTInverse('bior6.8', mode='zero')((yl, [yh]))
// per axis   y[i] = sum_m lo[m]*g0[i+8-2m] + hi[m]*g1[i+8-2m]   (0 <= i+8-2m < 18, zero outside),
// H axis first on (2*LL, yh0) and (yh1, yh2), then W axis (pytorch_wavelets SFB2D order).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define TNL_HD __host__ __device__ __forceinline__
#else
#define TNL_HD inline
#endif

namespace tnl {

// PyWavelets bior6.8 synthesis taps (rec_lo = g0 has 11 non-zero taps, rec_hi = g1 has 17).
TNL_HD constexpr float rec_lo(int k) {
    switch (k) {
        case 3: return 0.014426282505624435f;
        case 4: return 0.014467504896790148f;
        case 5: return -0.07872200106262882f;
        case 6: return -0.04036797903033992f;
        case 7: return 0.41784910915027457f;
        case 8: return 0.7589077294536541f;
        case 9: return 0.41784910915027457f;
        case 10: return -0.04036797903033992f;
        case 11: return -0.07872200106262882f;
        case 12: return 0.014467504896790148f;
        case 13: return 0.014426282505624435f;
        default: return 0.0f;
    }
}
TNL_HD constexpr float rec_hi(int k) {
    switch (k) {
        case 1: return -0.0019088317364812906f;
        case 2: return -0.0019142861290887667f;
        case 3: return 0.016990639867602342f;
        case 4: return 0.01193456527972926f;
        case 5: return -0.04973290349094079f;
        case 6: return -0.07726317316720414f;
        case 7: return 0.09405920349573646f;
        case 8: return 0.4207962846098268f;
        case 9: return -0.8259229974584023f;
        case 10: return 0.4207962846098268f;
        case 11: return 0.09405920349573646f;
        case 12: return -0.07726317316720414f;
        case 13: return -0.04973290349094079f;
        case 14: return 0.01193456527972926f;
        case 15: return 0.016990639867602342f;
        case 16: return -0.0019142861290887667f;
        case 17: return -0.0019088317364812906f;
        default: return 0.0f;
    }
}

// async 4-byte global->shared copy with zero fill; the host emulator provides a synchronous version
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;  // src-size 0 => destination zero-filled (this IS the zero padding)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#else
inline void cp_async4(float* smem_dst, const float* gsrc, bool valid) { *smem_dst = valid ? *gsrc : 0.0f; }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#endif

template <int CG_, int TW_>
struct IdwtCfg {
    static constexpr int CG = CG_;      // channels per CTA
    static constexpr int TW = TW_;      // fine-resolution columns per strip (32)
    static constexpr int TM = TW / 2;   // coarse-resolution columns per strip (16)
    static constexpr int RA = 3;        // coarse rows per step
    static constexpr int RB = 2 * RA;   // fine rows per step
    static constexpr int WIN = TM + 8;  // coarse columns incl. halo (24) == fine columns incl. halo / 2
    static constexpr int NT = WIN * CG; // threads per CTA
    static constexpr int pad_to(int len) { return len + (((CG - len) % 32) + 32) % 32; }  // stride == CG (mod 32)
    static constexpr int RS_F = pad_to(WIN * CG);      // forward mid row stride (floats)
    static constexpr int RS_B = pad_to(2 * WIN * CG);  // backward mid row stride (48 fine columns)
    static constexpr int MID_F = 2 * RB * RS_F;        // floats per forward mid buffer  (2 bands x 6 rows)
    static constexpr int MID_B = 2 * RA * RS_B;        // floats per backward mid buffer (2 bands x 3 rows)
    static constexpr int STAGE = 4 * RA * NT;          // floats per input stage (12 per thread)
    static constexpr size_t SMEM_F = sizeof(float) * (2 * MID_F + 2 * STAGE);
    static constexpr size_t SMEM_B = sizeof(float) * (2 * MID_B + 2 * STAGE);
    static_assert(RB * (TW / 8) * CG == NT, "forward phase-B item count must equal the thread count");
    static_assert(RA * (TM / 2) * CG == NT, "backward phase-B item count must equal the thread count");
};

struct IdwtBlock {
    int bx, by, bz;
};

// per-thread geometry shared by both directions
struct IdwtGeom {
    int n, C, plane, c0, m0, row_lo, row_hi, col, chan;
    int begin, end, nsteps;  // streamed coarse rows (fwd) / fine row pairs (bwd): [begin, end)
};

template <typename Cfg>
TNL_HD IdwtGeom idwt_geom(int tid, IdwtBlock b, int n, int C, int rows_per_cta) {
    IdwtGeom g;
    g.n = n;
    g.C = C;
    const int chunks = C / Cfg::CG;
    g.plane = b.bz / chunks;
    g.c0 = (b.bz % chunks) * Cfg::CG;
    g.m0 = b.bx * Cfg::TM;
    g.row_lo = b.by * rows_per_cta;
    g.row_hi = g.row_lo + rows_per_cta < n ? g.row_lo + rows_per_cta : n;
    g.col = tid / Cfg::CG;
    g.chan = g.c0 + tid % Cfg::CG;
    // centre row m needs streamed rows m-4 .. m+4 (bwd: fine row pairs m-4 .. m+4); m is emitted when m+4 arrives
    g.begin = g.row_lo - 4;
    g.end = g.row_hi + 4;
    g.nsteps = (g.end - g.begin + Cfg::RA - 1) / Cfg::RA;
    return g;
}

// ================================================================================================
// forward
// ================================================================================================
struct FwdState {
    float wLL[9], wLH[9], wHL[9], wHH[9];
    float abs_acc;  // sum |yh| over the detail coefficients this thread owns (wavelet L1 regulariser by-product)
};

TNL_HD void fwd_state_init(FwdState& st) {
    for (int i = 0; i < 9; ++i) st.wLL[i] = st.wLH[i] = st.wHL[i] = st.wHH[i] = 0.f;
    st.abs_acc = 0.f;
}

// stage the RA coarse rows of step ss (rows begin + ss*RA ..) for this thread's column
template <typename Cfg>
TNL_HD void fwd_issue_stage(const IdwtGeom& g, float* stage, const float* x, const float* yh, int tid, int ss) {
    const size_t plane_px = (size_t)g.n * g.n;
    const int mcol = g.m0 - 4 + g.col;
    const bool col_ok = mcol >= 0 && mcol < g.n;
    const int j0 = g.begin + ss * Cfg::RA;
#pragma unroll
    for (int rr = 0; rr < Cfg::RA; ++rr) {
        const int j = j0 + rr;
        const bool ok = col_ok && j >= 0 && j < g.n && j < g.end;
        const size_t px = ok ? ((size_t)j * g.n + mcol) : 0;
        cp_async4(stage + (rr * 4 + 0) * Cfg::NT + tid, x + ((size_t)g.plane * plane_px + px) * g.C + g.chan, ok);
        cp_async4(stage + (rr * 4 + 1) * Cfg::NT + tid, yh + ((size_t)(g.plane * 3 + 0) * plane_px + px) * g.C + g.chan, ok);
        cp_async4(stage + (rr * 4 + 2) * Cfg::NT + tid, yh + ((size_t)(g.plane * 3 + 1) * plane_px + px) * g.C + g.chan, ok);
        cp_async4(stage + (rr * 4 + 3) * Cfg::NT + tid, yh + ((size_t)(g.plane * 3 + 2) * plane_px + px) * g.C + g.chan, ok);
    }
    cp_async_commit();
}

// H-axis synthesis of one (lo, hi) pair; newest sample at ring slot SLOT:
//   W[d] = w[(SLOT + 1 + d) % 9], d = 0 (row m-4) .. 8 (row m+4);  even = sum W[d] g[16-2d], odd = sum W[d] g[17-2d]
template <int SLOT, bool DOUBLE_LO>
TNL_HD void synth_pair(const float (&lo)[9], const float (&hi)[9], float& even, float& odd) {
    float e = 0.f, o = 0.f;
#pragma unroll
    for (int d = 0; d < 9; ++d) {
        const float wl = lo[(SLOT + 1 + d) % 9], wh = hi[(SLOT + 1 + d) % 9];
        const float s = DOUBLE_LO ? 2.0f : 1.0f;  // yl = 2*x folded into the low-pass taps (exact scaling)
        const float ge0 = s * rec_lo(16 - 2 * d), go0 = s * rec_lo(17 - 2 * d);
        const float ge1 = rec_hi(16 - 2 * d), go1 = rec_hi(17 - 2 * d);
        if (ge0 != 0.f) e = fmaf(ge0, wl, e);
        if (go0 != 0.f) o = fmaf(go0, wl, o);
        if (ge1 != 0.f) e = fmaf(ge1, wh, e);
        if (go1 != 0.f) o = fmaf(go1, wh, o);
    }
    even = e;
    odd = o;
}

template <int SLOT>
TNL_HD void fwd_row(FwdState& st, const float* stage, float* mid, int tid, int rr, int NT, int RB, int RS, bool own) {
    st.wLL[SLOT] = stage[(rr * 4 + 0) * NT + tid];
    st.wLH[SLOT] = stage[(rr * 4 + 1) * NT + tid];
    st.wHL[SLOT] = stage[(rr * 4 + 2) * NT + tid];
    st.wHH[SLOT] = stage[(rr * 4 + 3) * NT + tid];
    if (own) st.abs_acc += fabsf(st.wLH[SLOT]) + fabsf(st.wHL[SLOT]) + fabsf(st.wHH[SLOT]);
    float e0, o0, e1, o1;
    synth_pair<SLOT, true>(st.wLL, st.wLH, e0, o0);
    synth_pair<SLOT, false>(st.wHL, st.wHH, e1, o1);
    mid[(0 * RB + 2 * rr) * RS + tid] = e0;
    mid[(0 * RB + 2 * rr + 1) * RS + tid] = o0;
    mid[(1 * RB + 2 * rr) * RS + tid] = e1;
    mid[(1 * RB + 2 * rr + 1) * RS + tid] = o1;
}

// phase A of step phase PH (= step index mod 3): RA rows enter the windows at ring slots 3*PH .. 3*PH+2
template <typename Cfg, int PH>
TNL_HD void fwd_phase_a(const IdwtGeom& g, FwdState& st, const float* stage, float* mid, int tid, int ss) {
    // a coefficient is "owned" by exactly one CTA: its column lies in the strip proper (not the halo) and its row in the chunk
    const bool col_own = g.col >= 4 && g.col < 4 + Cfg::TM;
    const int j0 = g.begin + ss * Cfg::RA;
    fwd_row<(PH * 3 + 0) % 9>(st, stage, mid, tid, 0, Cfg::NT, Cfg::RB, Cfg::RS_F, col_own && j0 + 0 >= g.row_lo && j0 + 0 < g.row_hi);
    fwd_row<(PH * 3 + 1) % 9>(st, stage, mid, tid, 1, Cfg::NT, Cfg::RB, Cfg::RS_F, col_own && j0 + 1 >= g.row_lo && j0 + 1 < g.row_hi);
    fwd_row<(PH * 3 + 2) % 9>(st, stage, mid, tid, 2, Cfg::NT, Cfg::RB, Cfg::RS_F, col_own && j0 + 2 >= g.row_lo && j0 + 2 < g.row_hi);
}

// phase B: W-axis synthesis. item = (mid row r, strip sb of 8 output columns, channel cb)
template <typename Cfg>
TNL_HD void fwd_phase_b(const IdwtGeom& g, const float* mid, float* out, int tid, int ss) {
    const int cb = tid % Cfg::CG;
    const int rs = tid / Cfg::CG;
    const int r = rs % Cfg::RB;
    const int sb = rs / Cfg::RB;
    const int m = g.begin + ss * Cfg::RA + (r >> 1) - 4;  // coarse row whose synthesis produced mid row r
    if (m < g.row_lo || m >= g.row_hi) return;
    const float* mlo = mid + (0 * Cfg::RB + r) * Cfg::RS_F + (4 * sb) * Cfg::CG + cb;
    const float* mhi = mid + (1 * Cfg::RB + r) * Cfg::RS_F + (4 * sb) * Cfg::CG + cb;
    float lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        hi[i] = mhi[i * Cfg::CG];
        lo[i] = (i >= 2 && i <= 10) ? mlo[i * Cfg::CG] : 0.f;  // rec_lo only has taps at d = 2..7
    }
    const int Y = 2 * m + (r & 1);
    const int n2 = 2 * g.n;
    float* orow = out + (((size_t)g.plane * n2 + Y) * n2) * g.C + g.c0 + cb;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float ev = 0.f, od = 0.f;
#pragma unroll
        for (int d = 0; d < 9; ++d) {
            const float ge0 = rec_lo(16 - 2 * d), go0 = rec_lo(17 - 2 * d);
            const float ge1 = rec_hi(16 - 2 * d), go1 = rec_hi(17 - 2 * d);
            if (ge0 != 0.f) ev = fmaf(ge0, lo[e + d], ev);
            if (go0 != 0.f) od = fmaf(go0, lo[e + d], od);
            if (ge1 != 0.f) ev = fmaf(ge1, hi[e + d], ev);
            if (go1 != 0.f) od = fmaf(go1, hi[e + d], od);
        }
        const int X = 2 * (g.m0 + 4 * sb + e);
        if (X < n2) {
            orow[(size_t)X * g.C] = ev;
            orow[(size_t)(X + 1) * g.C] = od;
        }
    }
}

// ================================================================================================
// backward (adjoint).  g_out [3][2n][2n][C] -> g_x [3][n][n][C] (includes the factor 2), g_yh [3][3][n][n][C]
//   a[m][X] = sum_k g0[k] G[2m-8+k][X]         b[m][X] = sum_k g1[k] G[2m-8+k][X]        (H-axis adjoint)
//   gLL[m][w] = 2 sum_k g0[k] a[m][2w-8+k]      gHL[m][w] = sum_k g1[k] a[m][2w-8+k]      (W-axis adjoint)
//   gLH[m][w] =   sum_k g0[k] b[m][2w-8+k]      gHH[m][w] = sum_k g1[k] b[m][2w-8+k]
// thread <-> two fine columns (Xa = 2*m0-8+col, Xb = Xa + WIN) x channel, 18-deep register windows.
// ================================================================================================
struct BwdState {
    float wa[18], wb[18];
};

TNL_HD void bwd_state_init(BwdState& st) {
    for (int i = 0; i < 18; ++i) st.wa[i] = st.wb[i] = 0.f;
}

template <typename Cfg>
TNL_HD void bwd_issue_stage(const IdwtGeom& g, float* stage, const float* gout, int tid, int ss) {
    const int n2 = 2 * g.n;
    const int Xa = 2 * g.m0 - 8 + g.col, Xb = Xa + Cfg::WIN;
    const int y0 = 2 * (g.begin + ss * Cfg::RA), y_end = 2 * g.end;
#pragma unroll
    for (int rr = 0; rr < 2 * Cfg::RA; ++rr) {
        const int y = y0 + rr;
        const bool rok = y >= 0 && y < n2 && y < y_end;
        const bool oka = rok && Xa >= 0 && Xa < n2;
        const bool okb = rok && Xb >= 0 && Xb < n2;
        const float* pa = gout + (((size_t)g.plane * n2 + (oka ? y : 0)) * n2 + (oka ? Xa : 0)) * g.C + g.chan;
        const float* pb = gout + (((size_t)g.plane * n2 + (okb ? y : 0)) * n2 + (okb ? Xb : 0)) * g.C + g.chan;
        cp_async4(stage + (rr * 2 + 0) * Cfg::NT + tid, pa, oka);
        cp_async4(stage + (rr * 2 + 1) * Cfg::NT + tid, pb, okb);
    }
    cp_async_commit();
}

// newest sample (k = 17) at ring slot SLOT;  U[k] = w[(SLOT + 1 + k) % 18]
template <int SLOT>
TNL_HD void analyse(const float (&w)[18], float& a, float& b) {
    float ra = 0.f, rb = 0.f;
#pragma unroll
    for (int k = 0; k < 18; ++k) {
        const float u = w[(SLOT + 1 + k) % 18];
        const float g0 = rec_lo(k), g1 = rec_hi(k);
        if (g0 != 0.f) ra = fmaf(g0, u, ra);
        if (g1 != 0.f) rb = fmaf(g1, u, rb);
    }
    a = ra;
    b = rb;
}

template <int SLOT>  // SLOT = ring slot of the second (newest) row of the pair; odd
TNL_HD void bwd_row(BwdState& st, const float* stage, float* mid, int tid, int rr, int NT, int RA, int RS) {
    st.wa[(SLOT + 17) % 18] = stage[((2 * rr) * 2 + 0) * NT + tid];
    st.wb[(SLOT + 17) % 18] = stage[((2 * rr) * 2 + 1) * NT + tid];
    st.wa[SLOT] = stage[((2 * rr + 1) * 2 + 0) * NT + tid];
    st.wb[SLOT] = stage[((2 * rr + 1) * 2 + 1) * NT + tid];
    float aa, ba, ab, bb;
    analyse<SLOT>(st.wa, aa, ba);
    analyse<SLOT>(st.wb, ab, bb);
    // mid layout [band a/b][row rr][fine column 0..2*WIN)[CG]; this thread owns columns col and col + WIN
    mid[(0 * RA + rr) * RS + tid] = aa;
    mid[(0 * RA + rr) * RS + NT + tid] = ab;
    mid[(1 * RA + rr) * RS + tid] = ba;
    mid[(1 * RA + rr) * RS + NT + tid] = bb;
}

template <typename Cfg, int PH>
TNL_HD void bwd_phase_a(BwdState& st, const float* stage, float* mid, int tid) {
    bwd_row<(2 * (PH * 3 + 0) + 1) % 18>(st, stage, mid, tid, 0, Cfg::NT, Cfg::RA, Cfg::RS_B);
    bwd_row<(2 * (PH * 3 + 1) + 1) % 18>(st, stage, mid, tid, 1, Cfg::NT, Cfg::RA, Cfg::RS_B);
    bwd_row<(2 * (PH * 3 + 2) + 1) % 18>(st, stage, mid, tid, 2, Cfg::NT, Cfg::RA, Cfg::RS_B);
}

// phase B: W-axis adjoint. item = (row r, pair of coarse columns sb, channel cb)
TNL_HD float signf_(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }  // torch.sign

// yh / reg: optional fused gradient of the wavelet L1 regulariser, g_yh += reg * sign(yh)  (nerf/utils.py:640-655)
template <typename Cfg>
TNL_HD void bwd_phase_b(const IdwtGeom& g, const float* mid, float* g_x, float* g_yh, int tid, int ss, const float* yh,
                        float reg) {
    const int cb = tid % Cfg::CG;
    const int rs = tid / Cfg::CG;  // 0 .. 23 = 3 rows x 8 column pairs
    const int r = rs % Cfg::RA;
    const int sb = rs / Cfg::RA;   // coarse columns m0 + 2*sb, m0 + 2*sb + 1
    const int m = g.begin + ss * Cfg::RA + r - 4;
    if (m < g.row_lo || m >= g.row_hi) return;
    // fine column X = 2w - 8 + k ; buffer column = X - (2*m0 - 8) = 2*(w - m0) + k
    const float* ma = mid + (0 * Cfg::RA + r) * Cfg::RS_B + (4 * sb) * Cfg::CG + cb;
    const float* mb = mid + (1 * Cfg::RA + r) * Cfg::RS_B + (4 * sb) * Cfg::CG + cb;
    float va[20], vb[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        va[i] = ma[i * Cfg::CG];
        vb[i] = mb[i * Cfg::CG];
    }
    const size_t plane_px = (size_t)g.n * g.n;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        float ll = 0.f, hl = 0.f, lh = 0.f, hh = 0.f;
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            const float g0 = rec_lo(k), g1 = rec_hi(k);
            if (g0 != 0.f) {
                ll = fmaf(g0, va[2 * e + k], ll);
                lh = fmaf(g0, vb[2 * e + k], lh);
            }
            if (g1 != 0.f) {
                hl = fmaf(g1, va[2 * e + k], hl);
                hh = fmaf(g1, vb[2 * e + k], hh);
            }
        }
        const int w = g.m0 + 2 * sb + e;
        if (w < g.n) {
            const size_t px = (size_t)m * g.n + w;
            const size_t i0 = ((size_t)(g.plane * 3 + 0) * plane_px + px) * g.C + g.c0 + cb;
            const size_t i1 = ((size_t)(g.plane * 3 + 1) * plane_px + px) * g.C + g.c0 + cb;
            const size_t i2 = ((size_t)(g.plane * 3 + 2) * plane_px + px) * g.C + g.c0 + cb;
            if (yh != nullptr) {
                lh = fmaf(reg, signf_(yh[i0]), lh);
                hl = fmaf(reg, signf_(yh[i1]), hl);
                hh = fmaf(reg, signf_(yh[i2]), hh);
            }
            g_x[((size_t)g.plane * plane_px + px) * g.C + g.c0 + cb] = 2.0f * ll;
            g_yh[i0] = lh;
            g_yh[i1] = hl;
            g_yh[i2] = hh;
        }
    }
}

// launch geometry shared by the kernel launcher and the emulator
template <typename Cfg>
inline void idwt_grid(unsigned n, unsigned C, unsigned num_sm, unsigned& gx, unsigned& gy, unsigned& gz, unsigned& rows) {
    gx = (n + Cfg::TM - 1) / Cfg::TM;
    gz = 3 * (C / Cfg::CG);
    rows = 96;  // halo overhead 8/rows; shrink the chunk until the grid covers the machine a few times
    while (rows > 24 && gx * ((n + rows - 1) / rows) * gz < 4 * num_sm) rows /= 2;
    if (rows > n) rows = n;
    gy = (n + rows - 1) / rows;
}

}  // namespace tnl
