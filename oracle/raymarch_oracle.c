/*
 * raymarch_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the reference's ray-marching / compositing kernels
 * (raymarching/src/raymarching.cu; each function cites the lines it follows).  It is the parity oracle
 * for tests/ and bench.py's cpu_baseline; it is never linked into or called from raw_ngp_b200/.
 *
 * GPU arithmetic that matters for bit-exact sample counts is emulated explicitly:
 *   - nvcc's default -fmad=true contracts a*b+c inside one expression into one FFMA: written as fmaf();
 *   - the voxel index goes through double in the reference (0.5 * ...): written the same way;
 *   - division and sqrt are IEEE on both sides; build with -O2 -ffp-contract=off (see Makefile).
 * __expf (ex2.approx based) is not reproducible on the CPU: compositing is compared with a tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf1(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:56-81 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3_inv(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:42-54 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* raymarching.cu:91-145 */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                            float* nears, float* fars) {
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { t = near_y; near_y = far_y; far_y = t; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = 3.402823466e+38f; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { t = near_z; near_z = far_z; far_z = t; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = 3.402823466e+38f; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:162-198 */
void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float B = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
        const float C = fmaf(-radius, radius, fmaf(oz, oz, fmaf(oy, oy, ox * ox)));
        const float t = (-B + sqrtf(fmaf(B, B, -(A * C)))) / A;
        const float x = fmaf(t, dx, ox), y = fmaf(t, dy, oy), z = fmaf(t, dz, oz);
        const float theta = atan2f(sqrtf(fmaf(z, z, x * x)), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = fmaf(2 * theta, RPI, -1.0f);
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* raymarching.cu:214-226, 237-254 */
void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; n++) indices[n] = (int32_t)morton3(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}
void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];
        coords[n * 3] = (int32_t)morton3_inv((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3_inv((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3_inv((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:267-289 */
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:303-319 */
void orc_flatten_rays(const int32_t* rays, uint32_t N, uint32_t M, int32_t* res) {
    (void)M;
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t off = (uint32_t)rays[n * 2], num = (uint32_t)rays[n * 2 + 1];
        for (uint32_t i = 0; i < num; i++) res[off + i] = (int32_t)n;
    }
}

/* One probe of the marching loop, raymarching.cu:407-437 (train) == :778-808 (infer). */
typedef struct { float cx, cy, cz, dt, mip_bound; int nx, ny, nz; int keep; } probe_t;

static inline probe_t probe_at(float t, float ox, float oy, float oz, float dx, float dy, float dz, const uint8_t* grid,
                               float bound, int contract, float dt_gamma, float dt_min, float dt_max, uint32_t C, uint32_t H,
                               float H3) {
    probe_t q;
    const float x = clampf(fmaf(t, dx, ox), -bound, bound);
    const float y = clampf(fmaf(t, dy, oy), -bound, bound);
    const float z = clampf(fmaf(t, dz, oz), -bound, bound);
    q.dt = clampf(t * dt_gamma, dt_min, dt_max);
    const int a = mip_from_pos(x, y, z, (float)C), b = mip_from_dt(q.dt, (float)H, (float)C);
    const int level = a > b ? a : b;
    q.mip_bound = fminf(scalbnf(1.0f, level), bound);
    const float mip_rbound = 1 / q.mip_bound;
    q.cx = x; q.cy = y; q.cz = z;
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int outer = contract && mag > 1;
    if (outer) {
        const float s = (2 - 1 / mag) / mag;
        q.cx *= s; q.cy *= s; q.cz *= s;
    }
    q.nx = (int)clampf((float)(0.5 * (double)fmaf(q.cx, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
    q.ny = (int)clampf((float)(0.5 * (double)fmaf(q.cy, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
    q.nz = (int)clampf((float)(0.5 * (double)fmaf(q.cz, mip_rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
    const uint32_t index = (uint32_t)fmaf((float)level, H3, (float)morton3((uint32_t)q.nx, (uint32_t)q.ny, (uint32_t)q.nz));
    const int occ = grid[index / 8] & (1 << (index % 8));
    q.keep = occ || outer;
    return q;
}

/* raymarching.cu:468-480 */
static inline float skip_voxel(float t, const probe_t* q, float dx, float dy, float dz, float rdx, float rdy, float rdz,
                               float rH, float dt_gamma, float dt_min, float dt_max) {
    const float tx = fmaf(fmaf((q->nx + 0.5f + 0.5f * signf1(dx)) * rH, 2.0f, -1.0f), q->mip_bound, -q->cx) * rdx;
    const float ty = fmaf(fmaf((q->ny + 0.5f + 0.5f * signf1(dy)) * rH, 2.0f, -1.0f), q->mip_bound, -q->cy) * rdy;
    const float tz = fmaf(fmaf((q->nz + 0.5f + 0.5f * signf1(dz)) * rH, 2.0f, -1.0f), q->mip_bound, -q->cz) * rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        const float dt = clampf(t * dt_gamma, dt_min, dt_max);
        t += dt;
    } while (t < tt);
    return t;
}

/*
 * raymarching.cu:337-491, both passes.  rays[n] = (offset, count) with ray-ordered offsets (a legal outcome of
 * the reference's atomicAdd, the one its backward assumes).  If xyzs == NULL only counts/offsets are produced.
 * Returns M.
 */
uint32_t orc_march_rays_train(const float* rays_o, const float* rays_d, const float* rays_ldir, const uint8_t* grid,
                              float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                              uint32_t H, const float* nears, const float* fars, const float* noises, int32_t* rays,
                              float* xyzs, float* dirs, float* ts, float* ldirs) {
    const float SQRT3 = 1.7320508075688772f;
    const float rH = 1 / (float)H;
    const float H3 = (float)(H * H * H);
    const float dt_min = 2 * SQRT3 / (float)max_steps;
    const float dt_max = 2 * SQRT3 * bound / (float)H;
    uint32_t total = 0;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        const float far = fars[n];
        float t = nears[n];
        t = fmaf(clampf(t * dt_gamma, dt_min, dt_max), noises[n], t);
        uint32_t step = 0;
        const uint32_t offset = total;
        while (t < far && step < max_steps) {
            const probe_t q = probe_at(t, ox, oy, oz, dx, dy, dz, grid, bound, contract, dt_gamma, dt_min, dt_max, C, H, H3);
            if (q.keep) {
                t += q.dt;
                if (xyzs) {
                    const size_t i = (size_t)offset + step;
                    xyzs[i * 3] = q.cx; xyzs[i * 3 + 1] = q.cy; xyzs[i * 3 + 2] = q.cz;
                    dirs[i * 3] = dx; dirs[i * 3 + 1] = dy; dirs[i * 3 + 2] = dz;
                    ts[i * 2] = t; ts[i * 2 + 1] = q.dt;
                    if (rays_ldir && ldirs) {
                        ldirs[i * 3] = rays_ldir[n * 3]; ldirs[i * 3 + 1] = rays_ldir[n * 3 + 1]; ldirs[i * 3 + 2] = rays_ldir[n * 3 + 2];
                    }
                }
                step++;
            } else {
                t = skip_voxel(t, &q, dx, dy, dz, rdx, rdy, rdz, rH, dt_gamma, dt_min, dt_max);
            }
        }
        rays[n * 2] = (int32_t)offset;
        rays[n * 2 + 1] = (int32_t)step;
        total += step;
    }
    return total;
}

/* raymarching.cu:519-597 */
void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                                      uint32_t M, uint32_t N, float T_thresh, float* weights, float* weights_sum,
                                      float* depth, float* image) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t offset = (uint32_t)rays[n * 2], num_steps = (uint32_t)rays[n * 2 + 1];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        if (!(num_steps == 0 || offset + num_steps > M)) {
            for (uint32_t step = 0; step < num_steps; step++) {
                const size_t i = (size_t)offset + step;
                const float alpha = 1.0f - expf(-sigmas[i] * ts[i * 2 + 1]);
                const float weight = alpha * T;
                weights[i] = weight;
                r = fmaf(weight, rgbs[i * 3], r);
                g = fmaf(weight, rgbs[i * 3 + 1], g);
                b = fmaf(weight, rgbs[i * 3 + 2], b);
                ws += weight;
                d = fmaf(weight, ts[i * 2], d);
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[n] = ws; depth[n] = d;
        image[n * 3] = r; image[n * 3 + 1] = g; image[n * 3 + 2] = b;
    }
}

/* raymarching.cu:623-712 */
void orc_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum, const float* grad_depth,
                                       const float* grad_image, const float* sigmas, const float* rgbs, const float* ts,
                                       const int32_t* rays, const float* weights_sum, const float* depth,
                                       const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                       float* grad_rgbs) {
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t offset = (uint32_t)rays[n * 2], num_steps = (uint32_t)rays[n * 2 + 1];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float* gi = grad_image + n * 3;
        const float r_final = image[n * 3], g_final = image[n * 3 + 1], b_final = image[n * 3 + 2];
        const float ws_final = weights_sum[n], d_final = depth[n];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
        for (uint32_t step = 0; step < num_steps; step++) {
            const size_t i = (size_t)offset + step;
            const float alpha = 1.0f - expf(-sigmas[i] * ts[i * 2 + 1]);
            const float weight = alpha * T;
            r = fmaf(weight, rgbs[i * 3], r);
            g = fmaf(weight, rgbs[i * 3 + 1], g);
            b = fmaf(weight, rgbs[i * 3 + 2], b);
            ws += weight;
            d = fmaf(weight, ts[i * 2], d);
            T *= 1.0f - alpha;
            grad_rgbs[i * 3] = gi[0] * weight;
            grad_rgbs[i * 3 + 1] = gi[1] * weight;
            grad_rgbs[i * 3 + 2] = gi[2] * weight;
            grad_sigmas[i] = ts[i * 2 + 1] * (gi[0] * (T * rgbs[i * 3] - (r_final - r)) + gi[1] * (T * rgbs[i * 3 + 1] - (g_final - g)) +
                                              gi[2] * (T * rgbs[i * 3 + 2] - (b_final - b)) +
                                              (grad_weights_sum[n] + grad_weights[i]) * (T - (ws_final - ws)) +
                                              grad_depth[n] * (T * ts[i * 2] - (d_final - d)));
            if (T < T_thresh) break;
        }
    }
}

/* raymarching.cu:730-846; outputs must be zero-filled by the caller like the reference wrapper does. */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                    const float* rays_d, float bound, int contract, float dt_gamma, uint32_t max_steps, uint32_t C,
                    uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs,
                    float* ts, const float* noises) {
    (void)nears;
    const float SQRT3 = 1.7320508075688772f;
    const float rH = 1 / (float)H;
    const float H3 = (float)(H * H * H);
    const float dt_min = 2 * SQRT3 / (float)max_steps;
    const float dt_max = 2 * SQRT3 * bound / (float)H;
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float ox = rays_o[index * 3], oy = rays_o[index * 3 + 1], oz = rays_o[index * 3 + 2];
        const float dx = rays_d[index * 3], dy = rays_d[index * 3 + 1], dz = rays_d[index * 3 + 2];
        const float rdx = 1 / (dx + 1e-10f), rdy = 1 / (dy + 1e-10f), rdz = 1 / (dz + 1e-10f);
        const float far = fars[index];
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, dt_min, dt_max), noises[n], t);
        uint32_t step = 0;
        while (t < far && step < n_step) {
            const probe_t q = probe_at(t, ox, oy, oz, dx, dy, dz, grid, bound, contract, dt_gamma, dt_min, dt_max, C, H, H3);
            if (q.keep) {
                const size_t i = (size_t)n * n_step + step;
                xyzs[i * 3] = q.cx; xyzs[i * 3 + 1] = q.cy; xyzs[i * 3 + 2] = q.cz;
                dirs[i * 3] = dx; dirs[i * 3 + 1] = dy; dirs[i * 3 + 2] = dz;
                t += q.dt;
                ts[i * 2] = t; ts[i * 2 + 1] = q.dt;
                step++;
            } else {
                t = skip_voxel(t, &q, dx, dy, dz, rdx, rdy, rdz, rH, dt_gamma, dt_min, dt_max);
            }
        }
    }
}

/* raymarching.cu:859-941 */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                        const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth,
                        float* image) {
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        float t = 0.f, d = depth[index], r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        float weight_sum = weights_sum[index];
        uint32_t step = 0;
        while (step < n_step) {
            const size_t i = (size_t)n * n_step + step;
            if (ts[i * 2] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[i] * ts[i * 2 + 1]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t = ts[i * 2];
            d = fmaf(weight, t, d);
            r = fmaf(weight, rgbs[i * 3], r);
            g = fmaf(weight, rgbs[i * 3 + 1], g);
            b = fmaf(weight, rgbs[i * 3 + 2], b);
            if (T < T_thresh) break;
            step++;
        }
        if (step < n_step) rays_alive[n] = -1;
        else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}
