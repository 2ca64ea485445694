/* TEST INFRASTRUCTURE ONLY (oracle) -- plain-C CPU restatement of the reference's ray-marching
 * operators, one scalar loop per ray, fp32 arithmetic with the SAME rounding sequence the
 * reference's nvcc build produces (IEEE division, default -fmad=true contraction written out as
 * explicit fmaf() at the sites the reference's SASS contracts; see oracle/README.md).
 *
 * Restates /root/reference/aux_libs/raymarching/src/raymarching.cu:
 *   orc_near_far_from_aabb      :91-145      orc_sph_from_ray          :162-198
 *   orc_morton3d / _invert      :56-81,214-254   orc_packbits          :267-289
 *   orc_march_rays_train        :311-480  (helpers mip_from_pos :42-47, mip_from_dt :49-54)
 *   orc_composite_train_fwd     :500-577     orc_composite_train_bwd   :601-682
 *   orc_march_rays              :700-805     orc_composite_rays        :818-905
 * Differences by construction: rays are visited in index order, so sample slots are allocated in
 * ray order (the reference allocates with atomicAdd in a race-dependent order, :405-406); parity is
 * therefore defined after canonicalising by ray id.  expf() stands in for the GPU's __expf
 * (ex2.approx), so composited floats carry a tolerance; every integer/bit/position result is exact.
 *
 * Pinned by tests/golden/raymarch_ref.npz = outputs of the reference kernels themselves
 * (oracle/_ref, compiled unmodified from /root/reference) run on a B200.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -o oracle/_build/liboracle_raymarch.so oracle/raymarch.c -lm
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#define SQRT3F 1.7320508075688772f
#define RPIF 0.3183098861837907f

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float sgnf(float x) { return copysignf(1.0f, x); }

static inline uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton_enc(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
static inline uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

void orc_morton3d(const int32_t *coords, uint32_t N, int32_t *indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int32_t)morton_enc((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}

void orc_morton3d_invert(const int32_t *indices, uint32_t N, int32_t *coords) {
    for (uint32_t n = 0; n < N; ++n) {
        /* the reference shifts a signed int (arithmetic shift), then masks */
        int32_t v = indices[n];
        coords[3 * n + 0] = (int32_t)compact3((uint32_t)(v >> 0));
        coords[3 * n + 1] = (int32_t)compact3((uint32_t)(v >> 1));
        coords[3 * n + 2] = (int32_t)compact3((uint32_t)(v >> 2));
    }
}

void orc_packbits(const float *grid, uint32_t N, float thresh, uint8_t *bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t b = 0;
        for (int i = 0; i < 8; ++i)
            if (grid[8 * (size_t)n + i] > thresh) b |= (uint8_t)(1u << i);
        bitfield[n] = b;
    }
}

void orc_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                            float min_near, float *nears, float *fars) {
    for (uint32_t n = 0; n < N; ++n) {
        const float *o = rays_o + 3 * (size_t)n, *d = rays_d + 3 * (size_t)n;
        float tn = -FLT_MAX, tf = FLT_MAX;
        int miss = 0;
        for (int a = 0; a < 3 && !miss; ++a) {
            const float r = 1.0f / d[a];
            float t0 = (aabb[a] - o[a]) * r, t1 = (aabb[a + 3] - o[a]) * r;
            if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
            if (a == 0) { tn = t0; tf = t1; continue; }
            if (tn > t1 || t0 > tf) { miss = 1; break; }
            if (t0 > tn) tn = t0;
            if (t1 < tf) tf = t1;
        }
        if (miss) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (tn < min_near) tn = min_near;
        nears[n] = tn;
        fars[n] = tf;
    }
}

void orc_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const float *o = rays_o + 3 * (size_t)n, *d = rays_d + 3 * (size_t)n;
        const float A = fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0]));
        const float B = fmaf(o[2], d[2], fmaf(o[1], d[1], o[0] * d[0]));
        const float C = fmaf(o[2], o[2], fmaf(o[1], o[1], o[0] * o[0])) - radius * radius;
        const float t = (-B + sqrtf(B * B - A * C)) / A;
        const float x = fmaf(t, d[0], o[0]), y = fmaf(t, d[1], o[1]), z = fmaf(t, d[2], o[2]);
        const float theta = atan2f(sqrtf(fmaf(z, z, x * x)), y);
        const float phi = atan2f(z, x);
        coords[2 * n] = 2 * theta * RPIF - 1;
        coords[2 * n + 1] = phi * RPIF;
    }
}

/* ---- one marching decision at parameter t; shared by train and inference marching ---- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH;
    uint32_t C, H;
    float H3;
    const uint8_t *grid;
} march_ctx;

static void ctx_init(march_ctx *c, const float *o, const float *d, float bound, float dt_gamma,
                     uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid) {
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2];
    c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1.0f / d[0]; c->rdy = 1.0f / d[1]; c->rdz = 1.0f / d[2];
    c->bound = bound; c->dt_gamma = dt_gamma;
    c->dt_min = 2 * SQRT3F / (float)max_steps;
    c->dt_max = 2 * SQRT3F * (float)(1 << (C - 1)) / (float)H;
    c->rH = 1.0f / (float)H;
    c->C = C; c->H = H;
    c->H3 = (float)(H * H * H);
    c->grid = grid;
}

static inline int level_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static inline int level_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)(dt * H * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* returns 1 if the cell at t is occupied (sample emitted: xyz, dt valid), else 0 and *t advanced past the cell */
static inline int march_probe(const march_ctx *c, float *t, float *xyz, float *dt_out) {
    const float tt0 = *t;
    const float x = clampf(fmaf(tt0, c->dx, c->ox), -c->bound, c->bound);
    const float y = clampf(fmaf(tt0, c->dy, c->oy), -c->bound, c->bound);
    const float z = clampf(fmaf(tt0, c->dz, c->oz), -c->bound, c->bound);
    const float dt = clampf(tt0 * c->dt_gamma, c->dt_min, c->dt_max);
    const int la = level_from_pos(x, y, z, (float)c->C), lb = level_from_dt(dt, (float)c->H, (float)c->C);
    const int level = la > lb ? la : lb;
    const float mip_bound = fminf(scalbnf(1.0f, level), c->bound);
    const float mip_rbound = 1.0f / mip_bound;
    const float Hm1 = (float)(c->H - 1);
    /* 0.5 is a double literal in the reference: float fma, widen, scale (exact), narrow, truncate */
    const int nx = (int)clampf((float)(0.5 * fmaf(x, mip_rbound, 1.0f) * c->H), 0.0f, Hm1);
    const int ny = (int)clampf((float)(0.5 * fmaf(y, mip_rbound, 1.0f) * c->H), 0.0f, Hm1);
    const int nz = (int)clampf((float)(0.5 * fmaf(z, mip_rbound, 1.0f) * c->H), 0.0f, Hm1);
    const uint32_t index = (uint32_t)(level * c->H3 + (float)morton_enc((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = c->grid[index / 8] & (1 << (index % 8));
    if (occ) {
        xyz[0] = x; xyz[1] = y; xyz[2] = z;
        *dt_out = dt;
        return 1;
    }
    const float tx = (fmaf(fmaf((float)nx + 0.5f + 0.5f * sgnf(c->dx), c->rH * 2, 0.0f) - 1, mip_bound, -x)) * c->rdx;
    const float ty = (fmaf(fmaf((float)ny + 0.5f + 0.5f * sgnf(c->dy), c->rH * 2, 0.0f) - 1, mip_bound, -y)) * c->rdy;
    const float tz = (fmaf(fmaf((float)nz + 0.5f + 0.5f * sgnf(c->dz), c->rH * 2, 0.0f) - 1, mip_bound, -z)) * c->rdz;
    const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tc = tt0;
    do {
        tc += clampf(tc * c->dt_gamma, c->dt_min, c->dt_max);
    } while (tc < tt);
    *t = tc;
    return 0;
}

/* rays: [N,3] = (ray id, offset, count) in ray order; counter[0] += total samples, counter[1] += N */
void orc_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                          float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                          const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                          int32_t *rays, int32_t *counter, const float *noises) {
    for (uint32_t n = 0; n < N; ++n) {
        march_ctx c;
        ctx_init(&c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[n];
        float t0 = nears[n];
        t0 = fmaf(clampf(t0 * dt_gamma, c.dt_min, c.dt_max), noises[n], t0);
        /* pass 1: count */
        float t = t0, xyz[3], dt;
        uint32_t num = 0;
        while (t < far && num < max_steps) {
            if (march_probe(&c, &t, xyz, &dt)) { num++; t += dt; }
        }
        const uint32_t off = (uint32_t)counter[0];
        counter[0] += (int32_t)num;
        const uint32_t slot = (uint32_t)counter[1];
        counter[1] += 1;
        rays[3 * slot] = (int32_t)n; rays[3 * slot + 1] = (int32_t)off; rays[3 * slot + 2] = (int32_t)num;
        if (num == 0 || off + num > M) continue;
        /* pass 2: write */
        float *px = xyzs + 3 * (size_t)off, *pd = dirs + 3 * (size_t)off, *pl = deltas + 2 * (size_t)off;
        t = t0;
        float last_t = t;
        uint32_t step = 0;
        while (t < far && step < num) {
            if (march_probe(&c, &t, xyz, &dt)) {
                px[0] = xyz[0]; px[1] = xyz[1]; px[2] = xyz[2];
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

void orc_composite_train_fwd(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                             uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num = (uint32_t)rays[3 * n + 2];
        float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0, T = 1.0f;
        if (num != 0 && offset + num <= M) {
            const float *s = sigmas + offset, *c = rgbs + 3 * (size_t)offset, *dl = deltas + 2 * (size_t)offset;
            for (uint32_t k = 0; k < num; ++k, ++s, c += 3, dl += 2) {
                const float alpha = 1.0f - expf(-s[0] * dl[0]);
                const float w = alpha * T;
                r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
                t += dl[1];
                d = fmaf(w, t, d);
                ws += w;
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[index] = ws; depth[index] = d;
        image[3 * index] = r; image[3 * index + 1] = g; image[3 * index + 2] = b;
    }
}

void orc_composite_train_bwd(const float *grad_ws, const float *grad_image, const float *sigmas, const float *rgbs,
                             const float *deltas, const int32_t *rays, const float *weights_sum, const float *image,
                             uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas, float *grad_rgbs) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num = (uint32_t)rays[3 * n + 2];
        if (num == 0 || offset + num > M) continue;
        const float *gi = grad_image + 3 * (size_t)index;
        const float gws = grad_ws[index], wsf = weights_sum[index];
        const float rf = image[3 * index], gf = image[3 * index + 1], bf = image[3 * index + 2];
        const float *s = sigmas + offset, *c = rgbs + 3 * (size_t)offset, *dl = deltas + 2 * (size_t)offset;
        float *gs = grad_sigmas + offset, *gc = grad_rgbs + 3 * (size_t)offset;
        float r = 0, g = 0, b = 0, ws = 0, T = 1.0f;
        for (uint32_t k = 0; k < num; ++k, ++s, c += 3, dl += 2, ++gs, gc += 3) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float w = alpha * T;
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            ws += w;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * w; gc[1] = gi[1] * w; gc[2] = gi[2] * w;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (rf - r)) + gi[1] * (T * c[1] - (gf - g)) +
                             gi[2] * (T * c[2] - (bf - b)) + gws * (1 - wsf));
            if (T < T_thresh) break;
        }
    }
}

void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                    const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                    uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                    float *xyzs, float *dirs, float *deltas, const float *noises) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        march_ctx c;
        ctx_init(&c, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H, grid);
        float *px = xyzs + 3 * (size_t)n * n_step, *pd = dirs + 3 * (size_t)n * n_step, *pl = deltas + 2 * (size_t)n * n_step;
        const float far = fars[index];
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t);
        float last_t = t, xyz[3], dt;
        uint32_t step = 0;
        while (t < far && step < n_step) {
            if (march_probe(&c, &t, xyz, &dt)) {
                px[0] = xyz[0]; px[1] = xyz[1]; px[2] = xyz[2];
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                        const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum,
                        float *depth, float *image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const float *s = sigmas + (size_t)n * n_step, *c = rgbs + 3 * (size_t)n * n_step, *dl = deltas + 2 * (size_t)n * n_step;
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[3 * index], g = image[3 * index + 1], b = image[3 * index + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            t += dl[1];
            d = fmaf(w, t, d);
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            if (T < T_thresh) break;
            ++s; c += 3; dl += 2; ++step;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = ws; depth[index] = d;
        image[3 * index] = r; image[3 * index + 1] = g; image[3 * index + 2] = b;
    }
}
