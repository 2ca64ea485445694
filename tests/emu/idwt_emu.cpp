// TEST-ONLY host emulator of the IDWT kernels: runs the per-thread phase functions of
// trinerflet_b200/csrc/idwt_core.cuh in lock-step (all threads phase A, barrier, all threads phase B),
// so the CPU test-suite can check the kernels' index arithmetic against the oracle without a GPU.
// Never linked into the product library.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../trinerflet_b200/csrc/idwt_core.cuh"

using namespace tnl;

// one CTA of the forward kernel; `geom(t)` yields thread t's geometry (dense grid mapping or work-list item)
template <typename Cfg, typename Geom>
static void emu_fwd_cta(const float* x, const float* yh, float* out, float* abs_sum, Geom geom) {
    std::vector<float> smem(2 * Cfg::MID_F + 3 * Cfg::STAGE);
    std::vector<FwdState> st(Cfg::NT);
    std::vector<IdwtGeom> geo(Cfg::NT);
    {
        float* mid0 = smem.data();
        float* stage0 = smem.data() + 2 * Cfg::MID_F;
        for (int t = 0; t < Cfg::NT; ++t) {
            geo[t] = geom(t);
            fwd_state_init<Cfg>(st[t], geo[t], x, yh, t);
            fwd_issue_stage<Cfg>(geo[t], st[t], stage0, t);
            fwd_issue_stage<Cfg>(geo[t], st[t], stage0 + Cfg::STAGE, t);
        }
        const int nsteps = geo[0].nsteps;
        for (int ss = 0; ss < nsteps; ++ss) {
            float* mid = mid0 + (ss & 1) * Cfg::MID_F;
            const float* stage = stage0 + (ss % 3) * Cfg::STAGE;
            for (int t = 0; t < Cfg::NT; ++t) {
                fwd_issue_stage<Cfg>(geo[t], st[t], stage0 + ((ss + 2) % 3) * Cfg::STAGE, t);
                switch (ss % 3) {
                    case 0: fwd_phase_a<Cfg, 0>(geo[t], st[t], stage, mid, t, ss); break;
                    case 1: fwd_phase_a<Cfg, 1>(geo[t], st[t], stage, mid, t, ss); break;
                    default: fwd_phase_a<Cfg, 2>(geo[t], st[t], stage, mid, t, ss); break;
                }
            }
            for (int t = 0; t < Cfg::NT; ++t) fwd_phase_b<Cfg>(geo[t], st[t], mid, out, ss);
        }
        if (abs_sum) for (int t = 0; t < Cfg::NT; ++t) *abs_sum += st[t].abs_acc;
    }
}

template <typename Cfg>
static void emu_fwd(const float* x, const float* yh, float* out, unsigned n, unsigned C, float* abs_sum) {
    unsigned gx, gy, gz, rows;
    idwt_grid<Cfg>(n, C, 148, gx, gy, gz, rows);
    for (unsigned bz = 0; bz < gz; ++bz) for (unsigned by = 0; by < gy; ++by) for (unsigned bx = 0; bx < gx; ++bx)
        emu_fwd_cta<Cfg>(x, yh, out, abs_sum, [&](int t) {
            return idwt_geom<Cfg>(t, IdwtBlock{(int)bx, (int)by, (int)bz}, (int)n, (int)C, (int)rows);
        });
}

// work-list mode (k_idwt_fwd with items != nullptr): blockIdx.x = item * chunks + chunk
template <typename Cfg>
static void emu_fwd_items(const float* x, const float* yh, float* out, unsigned n, unsigned C, float* abs_sum, const int32_t* items,
                          int n_items) {
    const int chunks = C / Cfg::CG;
    for (int b = 0; b < n_items * chunks; ++b) {
        const int32_t* it = items + 4 * (b / chunks);
        emu_fwd_cta<Cfg>(x, yh, out, abs_sum, [&](int t) {
            return idwt_geom_item<Cfg>(t, b % chunks, IdwtItem{it[0], it[1], it[2], it[3]}, (int)n, (int)C);
        });
    }
}

template <typename Cfg, typename Geom>
static void emu_bwd_cta(const float* g, float* g_x, float* g_yh, const float* yh, float reg, Geom geom) {
    std::vector<float> smem(2 * Cfg::MID_B + 3 * Cfg::STAGE);
    std::vector<BwdState> st(Cfg::NT);
    std::vector<IdwtGeom> geo(Cfg::NT);
    {
        float* mid0 = smem.data();
        float* stage0 = smem.data() + 2 * Cfg::MID_B;
        for (int t = 0; t < Cfg::NT; ++t) {
            geo[t] = geom(t);
            bwd_state_init<Cfg>(st[t], geo[t], g);
            bwd_issue_stage<Cfg>(geo[t], st[t], stage0, t);
            bwd_issue_stage<Cfg>(geo[t], st[t], stage0 + Cfg::STAGE, t);
        }
        const int nsteps = geo[0].nsteps;
        for (int ss = 0; ss < nsteps; ++ss) {
            float* mid = mid0 + (ss & 1) * Cfg::MID_B;
            const float* stage = stage0 + (ss % 3) * Cfg::STAGE;
            for (int t = 0; t < Cfg::NT; ++t) {
                bwd_issue_stage<Cfg>(geo[t], st[t], stage0 + ((ss + 2) % 3) * Cfg::STAGE, t);
                switch (ss % 3) {
                    case 0: bwd_phase_a<Cfg, 0>(st[t], stage, mid, t); break;
                    case 1: bwd_phase_a<Cfg, 1>(st[t], stage, mid, t); break;
                    default: bwd_phase_a<Cfg, 2>(st[t], stage, mid, t); break;
                }
            }
            for (int t = 0; t < Cfg::NT; ++t) bwd_phase_b<Cfg>(geo[t], mid, g_x, g_yh, t, ss, yh, reg);
        }
    }
}

template <typename Cfg>
static void emu_bwd(const float* g, float* g_x, float* g_yh, unsigned n, unsigned C, const float* yh, float reg) {
    unsigned gx, gy, gz, rows;
    idwt_grid<Cfg>(n, C, 148, gx, gy, gz, rows);
    for (unsigned bz = 0; bz < gz; ++bz) for (unsigned by = 0; by < gy; ++by) for (unsigned bx = 0; bx < gx; ++bx)
        emu_bwd_cta<Cfg>(g, g_x, g_yh, yh, reg, [&](int t) {
            return idwt_geom<Cfg>(t, IdwtBlock{(int)bx, (int)by, (int)bz}, (int)n, (int)C, (int)rows);
        });
}

template <typename Cfg>
static void emu_bwd_items(const float* g, float* g_x, float* g_yh, unsigned n, unsigned C, const float* yh, float reg,
                          const int32_t* items, int n_items) {
    const int chunks = C / Cfg::CG;
    for (int b = 0; b < n_items * chunks; ++b) {
        const int32_t* it = items + 4 * (b / chunks);
        emu_bwd_cta<Cfg>(g, g_x, g_yh, yh, reg, [&](int t) {
            return idwt_geom_item<Cfg>(t, b % chunks, IdwtItem{it[0], it[1], it[2], it[3]}, (int)n, (int)C);
        });
    }
}

extern "C" {
int emu_idwt_level_forward(const float* x, const float* yh, float* out, uint32_t n, uint32_t C, float* abs_sum) {
    if (C % 32 == 0) emu_fwd<IdwtCfg<32, 32>>(x, yh, out, n, C, abs_sum);
    else if (C % 24 == 0) emu_fwd<IdwtCfg<24, 32>>(x, yh, out, n, C, abs_sum);
    else if (C % 16 == 0) emu_fwd<IdwtCfg<16, 32>>(x, yh, out, n, C, abs_sum);
    else emu_fwd<IdwtCfg<8, 32>>(x, yh, out, n, C, abs_sum);
    return 0;
}
// work-list variants: the channel-chunk choice mirrors tnl_idwt_level_*_sparse (24 / 16 / 8 channels per CTA)
int emu_idwt_level_forward_items(const float* x, const float* yh, float* out, uint32_t n, uint32_t C, float* abs_sum,
                                 const int32_t* items, int32_t n_items) {
    if (C % 24 == 0) emu_fwd_items<IdwtCfg<24, 32>>(x, yh, out, n, C, abs_sum, items, n_items);
    else if (C % 16 == 0) emu_fwd_items<IdwtCfg<16, 32>>(x, yh, out, n, C, abs_sum, items, n_items);
    else emu_fwd_items<IdwtCfg<8, 32>>(x, yh, out, n, C, abs_sum, items, n_items);
    return 0;
}
int emu_idwt_level_backward_items(const float* g, float* g_x, float* g_yh, uint32_t n, uint32_t C, const float* yh, float reg,
                                  const int32_t* items, int32_t n_items) {
    if (C % 24 == 0) emu_bwd_items<IdwtCfg<24, 32>>(g, g_x, g_yh, n, C, yh, reg, items, n_items);
    else if (C % 16 == 0) emu_bwd_items<IdwtCfg<16, 32>>(g, g_x, g_yh, n, C, yh, reg, items, n_items);
    else emu_bwd_items<IdwtCfg<8, 32>>(g, g_x, g_yh, n, C, yh, reg, items, n_items);
    return 0;
}
int emu_idwt_level_backward(const float* g, float* g_x, float* g_yh, uint32_t n, uint32_t C, const float* yh, float reg) {
    if (C % 32 == 0) emu_bwd<IdwtCfg<32, 32>>(g, g_x, g_yh, n, C, yh, reg);
    else if (C % 24 == 0) emu_bwd<IdwtCfg<24, 32>>(g, g_x, g_yh, n, C, yh, reg);
    else if (C % 16 == 0) emu_bwd<IdwtCfg<16, 32>>(g, g_x, g_yh, n, C, yh, reg);
    else emu_bwd<IdwtCfg<8, 32>>(g, g_x, g_yh, n, C, yh, reg);
    return 0;
}
}
