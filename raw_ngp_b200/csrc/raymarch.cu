// raymarch.cu -- occupancy-grid ray marching and volume compositing for sm_100a.
//
// Operator semantics follow raymarching/src/raymarching.cu of the reference (kernel line ranges are cited
// at each function).  What is different:
//   * the occupancy bitfield (256 KiB per cascade at H=128) is read through ld.global.nc so it lives in
//     L1/L2; the cascade is selected in registers;
//   * the voxel index needs no FP64: 0.5*v is exact in fp32 and (0.5*v)*H rounds the exact product once,
//     which is the value the reference obtains through its double detour (raymarching.cu:432-434);
//   * sample offsets are an exclusive prefix sum in ray order (deterministic) computed by the last block
//     of the counting kernel, not an atomicAdd ticket in arrival order (raymarching.cu:486-490);
//   * compositing is one warp per ray: coalesced 32-sample chunks, transmittance by a shuffle product scan,
//     early termination by ballot;
//   * inference marching zero-fills its own tail, there is no memset per iteration.
// Do NOT build this file with --use_fast_math: sample counts are compared bit-exactly with the reference.
#include "common.cuh"

namespace ngp {
namespace {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float sign1(float x) { return copysignf(1.0f, x); }

// 10-bit-per-axis Morton code.  raymarching.cu:56-81.
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t v) {
    v &= 0x49249249u;
    v = (v | (v >> 2)) & 0xc30c30c3u;
    v = (v | (v >> 4)) & 0x0f00f00fu;
    v = (v | (v >> 8)) & 0xff0000ffu;
    v = (v | (v >> 16)) & 0x0000ffffu;
    return v;
}

// cascade level from position / from step size.  raymarching.cu:42-54.
__device__ __forceinline__ int cascade_from_pos(float x, float y, float z, float n_cascades) {
    const float m = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(m, &e);
    return fminf(n_cascades - 1, fmaxf(0, e));
}
__device__ __forceinline__ int cascade_from_dt(float dt, float H, float n_cascades) {
    const float m = dt * H * 0.5f;
    int e;
    frexpf(m, &e);
    return fminf(n_cascades - 1, fmaxf(0, e));
}

// Per-ray constants shared by the training and inference marchers.
struct MarchParams {
    const uint8_t* __restrict__ grid;
    float bound, dt_gamma, dt_min, dt_max, rH, H3, Hf, Cf;
    uint32_t H;
    bool contract;
    // one cascade and no contraction (bound <= 1): the cascade level is always 0 and mip_bound a constant, so probe() skips
    // the two frexpf, the scalbnf, the division and the contraction test -- about half of its instructions, same results
    bool simple;
    float mip_bound0, mip_rbound0;
};

__device__ __forceinline__ MarchParams make_params(const uint8_t* grid, float bound, bool contract, float dt_gamma,
                                                   uint32_t max_steps, uint32_t C, uint32_t H) {
    MarchParams p;
    p.grid = grid;
    p.bound = bound;
    p.contract = contract;
    p.dt_gamma = dt_gamma;
    p.dt_min = 2 * 1.7320508075688772f / max_steps;      // raymarching.cu:396
    p.dt_max = 2 * 1.7320508075688772f * bound / H;      // raymarching.cu:397
    p.rH = 1 / (float)H;
    p.H3 = H * H * H;
    p.Hf = (float)H;
    p.Cf = (float)C;
    p.H = H;
    p.simple = (C == 1) && !contract;
    p.mip_bound0 = fminf(1.0f, bound);            // fminf(scalbnf(1.0f, 0), bound)
    p.mip_rbound0 = 1 / p.mip_bound0;
    return p;
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

// One probe of the marching loop at parameter t (raymarching.cu:407-437 / 778-808): clamps the position,
// picks the cascade, contracts, finds the voxel and tests its bit.  Returns true when the sample is kept.
struct Probe {
    float cx, cy, cz, dt, mip_bound;
    int nx, ny, nz;
};
__device__ __forceinline__ bool probe(const MarchParams& p, const Ray& r, float t, Probe& q) {
    const float x = clampf(r.ox + t * r.dx, -p.bound, p.bound);
    const float y = clampf(r.oy + t * r.dy, -p.bound, p.bound);
    const float z = clampf(r.oz + t * r.dz, -p.bound, p.bound);
    q.dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);

    int level = 0;
    float mip_rbound = p.mip_rbound0;
    bool outer = false;
    q.mip_bound = p.mip_bound0;
    q.cx = x; q.cy = y; q.cz = z;
    if (!p.simple) {        // uniform over the launch
        level = max(cascade_from_pos(x, y, z, p.Cf), cascade_from_dt(q.dt, p.Hf, p.Cf));
        q.mip_bound = fminf(scalbnf(1.0f, level), p.bound);
        mip_rbound = 1 / q.mip_bound;
        const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        outer = p.contract && mag > 1;
        if (outer) {  // L-inf contraction, all axes scaled (raymarching.cu:423-429)
            const float s = (2 - 1 / mag) / mag;
            q.cx *= s; q.cy *= s; q.cz *= s;
        }
    }
    // 0.5*(c/mip_bound + 1)*H, clamped to [0, H-1], truncated
    q.nx = (int)clampf((0.5f * (q.cx * mip_rbound + 1)) * p.Hf, 0.0f, (float)(p.H - 1));
    q.ny = (int)clampf((0.5f * (q.cy * mip_rbound + 1)) * p.Hf, 0.0f, (float)(p.H - 1));
    q.nz = (int)clampf((0.5f * (q.cz * mip_rbound + 1)) * p.Hf, 0.0f, (float)(p.H - 1));

    // bit index is evaluated in fp32 like the reference (raymarching.cu:436)
    const uint32_t index = level * p.H3 + morton_encode(q.nx, q.ny, q.nz);
    const bool occ = __ldg(p.grid + index / 8) & (1 << (index % 8));
    return occ || outer;
}

// Advance t past the current (empty) voxel in dt-sized steps.  raymarching.cu:468-480.
__device__ __forceinline__ float skip_voxel(const MarchParams& p, const Ray& r, float t, const Probe& q) {
    const float tx = (((q.nx + 0.5f + 0.5f * sign1(r.dx)) * p.rH * 2 - 1) * q.mip_bound - q.cx) * r.rdx;
    const float ty = (((q.ny + 0.5f + 0.5f * sign1(r.dy)) * p.rH * 2 - 1) * q.mip_bound - q.cy) * r.rdy;
    const float tz = (((q.nz + 0.5f + 0.5f * sign1(r.dz)) * p.rH * 2 - 1) * q.mip_bound - q.cz) * r.rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do {
        const float dt = clampf(t * p.dt_gamma, p.dt_min, p.dt_max);
        t += dt;
    } while (t < tt);
    return t;
}

// ---------------------------------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------------------------------

// raymarching.cu:91-145
__global__ void near_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                const float* __restrict__ aabb, uint32_t N, float min_near, float* __restrict__ nears,
                                float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float lo[3] = {__ldg(aabb), __ldg(aabb + 1), __ldg(aabb + 2)};
    const float hi[3] = {__ldg(aabb + 3), __ldg(aabb + 4), __ldg(aabb + 5)};
    float tn = 0.f, tf = 0.f;
    bool miss = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float o = __ldg(rays_o + (size_t)n * 3 + a);
        const float rd = 1 / __ldg(rays_d + (size_t)n * 3 + a);
        float t0 = (lo[a] - o) * rd, t1 = (hi[a] - o) * rd;
        if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
        if (a == 0) { tn = t0; tf = t1; }
        else if (!miss) {
            if (tn > t1 || t0 > tf) miss = true;
            else { if (t0 > tn) tn = t0; if (t1 < tf) tf = t1; }
        }
    }
    if (miss) { nears[n] = fars[n] = 3.402823466e+38f; return; }  // numeric_limits<float>::max()
    if (tn < min_near) tn = min_near;
    nears[n] = tn;
    fars[n] = tf;
}

// raymarching.cu:162-198
__global__ void sph_from_ray_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                                    uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[n * 2] = 2 * theta * 0.3183098861837907f - 1;
    coords[n * 2 + 1] = phi * 0.3183098861837907f;
}

// raymarching.cu:214-226
__global__ void morton3d_kernel(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int)morton_encode(__ldg(coords + (size_t)n * 3), __ldg(coords + (size_t)n * 3 + 1), __ldg(coords + (size_t)n * 3 + 2));
}

// raymarching.cu:237-254
__global__ void morton3d_invert_kernel(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ind = __ldg(indices + n);   // arithmetic shifts of a signed value, like the reference
    coords[(size_t)n * 3] = (int)compact3(ind >> 0);
    coords[(size_t)n * 3 + 1] = (int)compact3(ind >> 1);
    coords[(size_t)n * 3 + 2] = (int)compact3(ind >> 2);
}

// raymarching.cu:267-289.  One thread packs 8 densities (two 16-byte loads) into one byte.
__global__ void packbits_kernel(const float* __restrict__ grid, uint32_t N, float density_thresh,
                                const float* __restrict__ thresh_dev, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float th = density_thresh;
    if (thresh_dev) th = fminf(__ldg(thresh_dev), density_thresh);
    const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + (size_t)n * 2);
    const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + (size_t)n * 2 + 1);
    uint32_t bits = 0;
    bits |= (a.x > th) ? 1u : 0u;   bits |= (a.y > th) ? 2u : 0u;
    bits |= (a.z > th) ? 4u : 0u;   bits |= (a.w > th) ? 8u : 0u;
    bits |= (b.x > th) ? 16u : 0u;  bits |= (b.y > th) ? 32u : 0u;
    bits |= (b.z > th) ? 64u : 0u;  bits |= (b.w > th) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// raymarching.cu:303-319; one warp per ray so the stores coalesce.
__global__ void flatten_rays_kernel(const int* __restrict__ rays, uint32_t N, uint32_t M, int* __restrict__ res) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t offset = __ldg(rays + (size_t)n * 2), num = __ldg(rays + (size_t)n * 2 + 1);
    for (uint32_t i = lane; i < num; i += 32)
        if (offset + i < M) res[offset + i] = (int)n;
}

// ---------------------------------------------------------------------------------------------------
// training march
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, size_t n, bool eps) {
    Ray r;
    r.ox = __ldg(rays_o + n * 3); r.oy = __ldg(rays_o + n * 3 + 1); r.oz = __ldg(rays_o + n * 3 + 2);
    r.dx = __ldg(rays_d + n * 3); r.dy = __ldg(rays_d + n * 3 + 1); r.dz = __ldg(rays_d + n * 3 + 2);
    if (eps) {  // inference marcher: raymarching.cu:762
        r.rdx = 1 / (r.dx + 1e-10f); r.rdy = 1 / (r.dy + 1e-10f); r.rdz = 1 / (r.dz + 1e-10f);
    } else {    // training marcher: raymarching.cu:388
        r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
    }
    return r;
}

constexpr uint32_t kMarchThreads = 64;

// Pass 1 (raymarching.cu:337-491 with xyzs == nullptr) + ray-ordered exclusive scan by the last block.
__global__ void __launch_bounds__(kMarchThreads)
march_train_count_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
                         float bound, bool contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                         const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
                         int* __restrict__ rays, int* __restrict__ counter) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        const MarchParams p = make_params(grid, bound, contract, dt_gamma, max_steps, C, H);
        const Ray r = load_ray(rays_o, rays_d, n, false);
        const float far = __ldg(fars + n);
        float t = __ldg(nears + n);
        t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * __ldg(noises + n);
        uint32_t step = 0;
        while (t < far && step < max_steps) {
            Probe q;
            if (probe(p, r, t, q)) { step++; t += q.dt; }
            else t = skip_voxel(p, r, t, q);
        }
        rays[(size_t)n * 2 + 1] = (int)step;
    }

    // last block to arrive turns the counts into ray-ordered offsets
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x >= 32) return;
    const uint32_t lane = threadIdx.x;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < N; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t cnt = (i < N) ? (uint32_t)__ldcg(rays + (size_t)i * 2 + 1) : 0u;
        uint32_t incl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= (uint32_t)s) incl += v;
        }
        if (i < N) rays[(size_t)i * 2] = (int)(carry + incl - cnt);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { counter[0] = (int)carry; counter[1] = 0; }
}

// Pass 2 (raymarching.cu:337-491 with output buffers): re-march and write the samples of each ray.
template <bool LDIR>
__global__ void __launch_bounds__(kMarchThreads)
march_train_write_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ rays_ldir,
                         const uint8_t* __restrict__ grid, float bound, bool contract, float dt_gamma, uint32_t max_steps,
                         uint32_t N, uint32_t C, uint32_t H, const float* __restrict__ nears, const float* __restrict__ fars,
                         const float* __restrict__ noises, const int* __restrict__ rays, uint32_t M, float* __restrict__ xyzs,
                         float* __restrict__ dirs, float* __restrict__ ts, float* __restrict__ ldirs) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2);
    const uint32_t num_steps = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
    if (num_steps == 0 || offset + num_steps > M) return;

    const MarchParams p = make_params(grid, bound, contract, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o, rays_d, n, false);
    float lx = 0, ly = 0, lz = 0;
    if (LDIR) { lx = __ldg(rays_ldir + (size_t)n * 3); ly = __ldg(rays_ldir + (size_t)n * 3 + 1); lz = __ldg(rays_ldir + (size_t)n * 3 + 2); }
    const float far = __ldg(fars + n);
    float t = __ldg(nears + n);
    t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * __ldg(noises + n);

    float* px = xyzs + (size_t)offset * 3;
    float* pd = dirs + (size_t)offset * 3;
    float* pt = ts + (size_t)offset * 2;
    float* pl = LDIR ? ldirs + (size_t)offset * 3 : nullptr;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
        Probe q;
        if (probe(p, r, t, q)) {
            step++;
            t += q.dt;
            px[0] = q.cx; px[1] = q.cy; px[2] = q.cz;   // contracted coordinates (raymarching.cu:446)
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            *reinterpret_cast<float2*>(pt) = make_float2(t, q.dt);  // t AFTER the step (raymarching.cu:458)
            if (LDIR) { pl[0] = lx; pl[1] = ly; pl[2] = lz; pl += 3; }
            px += 3; pd += 3; pt += 2;
        } else {
            t = skip_voxel(p, r, t, q);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// warp-cooperative training march
// ---------------------------------------------------------------------------------------------------
// The reference loop is serial per ray, but the sequence of t values it can ever visit is a fixed lattice
// u_0 = t0, u_{i+1} = u_i + clamp(u_i * dt_gamma, dt_min, dt_max): a kept sample advances by one lattice step and
// the empty-voxel skip `do { t += dt } while (t < tt)` advances to the first lattice point >= tt.  So one warp per ray
//   (1) builds a window of kWin lattice points (the only serial fp32 chain, identical adds to the reference),
//   (2) probes all of them in parallel (32 per instruction): keep flag, or the jump target found by binary search,
//   (3) lets one lane chase the next[] pointers (no memory traffic beyond shared memory),
// and repeats from the t where the chase left the window.  The t of every kept sample is stored so that pass 2
// writes all samples of a ray in parallel with coalesced stores and never re-marches.  Results are bit-identical
// to the serial loop: the same fp32 operations produce every t, and keep/skip decisions are taken on those t.
// Window of 512 lattice points = 12 KB of shared memory per block.  With 1024 (24 KB x 7 resident blocks per SM) the kernel took
// the 196 KB shared-memory carve-out and left the occupancy-bitfield probes ~60 KB of L1 while the optimizer of the previous step
// streams 366 MB through L2 beside it: the training step is 10 us faster with 512 (699 -> 689 us; 256: 689, 128: 691, 64: 694).
#ifndef NGP_MARCH_WINDOW
#define NGP_MARCH_WINDOW 512
#endif
constexpr uint32_t kWin = NGP_MARCH_WINDOW;
#ifndef NGP_MARCH_JUMP_ROUNDS
#define NGP_MARCH_JUMP_ROUNDS 2
#endif
constexpr uint32_t kCoopWarps = 4;

// tt of an empty probe (raymarching.cu:470-474)
__device__ __forceinline__ float voxel_exit(const MarchParams& p, const Ray& r, float t, const Probe& q) {
    const float tx = (((q.nx + 0.5f + 0.5f * sign1(r.dx)) * p.rH * 2 - 1) * q.mip_bound - q.cx) * r.rdx;
    const float ty = (((q.ny + 0.5f + 0.5f * sign1(r.dy)) * p.rH * 2 - 1) * q.mip_bound - q.cy) * r.rdy;
    const float tz = (((q.nz + 0.5f + 0.5f * sign1(r.dz)) * p.rH * 2 - 1) * q.mip_bound - q.cz) * r.rdz;
    return t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
}

// The differentiable slab test the renderer actually uses (nerf/renderer.py:139-158, torch ops in fp32):
// (aabb - o) / (d + 1e-15), near = max over axes of the smaller root, far = min of the larger, miss -> 1e9 for both,
// near clamped to min_near.
__device__ __forceinline__ void near_far_torch(const Ray& r, const float* __restrict__ aabb, float min_near, float& near, float& far) {
    const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
    near = -INFINITY; far = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float den = __fadd_rn(d[a], 1e-15f);
        const float t0 = __fdiv_rn(__fsub_rn(__ldg(aabb + a), o[a]), den);
        const float t1 = __fdiv_rn(__fsub_rn(__ldg(aabb + 3 + a), o[a]), den);
        const float lo = (t0 < t1) ? t0 : t1, hi = (t0 > t1) ? t0 : t1;
        near = fmaxf(near, lo);
        far = fminf(far, hi);
    }
    if (far < near) { near = 1e9f; far = 1e9f; }
    near = fmaxf(near, min_near);
}

// NF: near/far are computed here from the AABB (and written to nears/fars when those are non-null) instead of read.
// cap != 0: counter[2] receives the number of samples of the longest ray-ordered prefix that fits in `cap` rows.
template <bool NF>
__global__ void __launch_bounds__(kCoopWarps * 32)
march_train_count_coop_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
                              float bound, bool contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                              const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
                              int* __restrict__ rays, int* __restrict__ counter, float* __restrict__ t_scratch,
                              const float* __restrict__ aabb, float min_near, float* __restrict__ nears_out,
                              float* __restrict__ fars_out, uint32_t cap, const float* __restrict__ cam_near_far,
                              const int* __restrict__ n_rays_dev) {
    __shared__ float s_u[kCoopWarps][kWin];
    __shared__ uint16_t s_next[kCoopWarps][kWin];   // bit 15: keep, low bits: next lattice index (kWin = leaves the window)
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.x * kCoopWarps + warp;
    if (n < N) {
        const MarchParams p = make_params(grid, bound, contract, dt_gamma, max_steps, C, H);
        const Ray r = load_ray(rays_o, rays_d, n, false);
        float far, t;
        if (NF) {
            near_far_torch(r, aabb, min_near, t, far);
            if (cam_near_far) {      // renderer.py:529-533: nears = max(nears, cam_near), fars = min(fars, cam_far)
                t = fmaxf(t, __ldg(cam_near_far + (size_t)n * 2));
                far = fminf(far, __ldg(cam_near_far + (size_t)n * 2 + 1));
            }
            if (lane == 0 && nears_out) { nears_out[n] = t; fars_out[n] = far; }
        } else {
            far = __ldg(fars + n);
            t = __ldg(nears + n);
        }
        // adaptive ray count (train_utils.py:563-564): only the first *n_rays_dev rays of the batch are live, the others get
        // no samples (and no loss, see composite_train_mse_kernel)
        if (n_rays_dev && n >= (uint32_t)__ldg(n_rays_dev)) far = -1.0f;
        t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * __ldg(noises + n);
        float* u = s_u[warp];
        uint16_t* nx = s_next[warp];
        float* tout = t_scratch + (size_t)n * max_steps;
        uint32_t step = 0;
        while (t < far && step < max_steps) {       // warp-uniform
            // (1) lattice window starting at t
            uint32_t cnt = 0;
            const float dt_c = clampf(0.f, p.dt_min, p.dt_max);      // the step when dt_gamma == 0 (raymarching.cu:412)
            if (p.dt_gamma == 0.f && t >= dt_c && dt_c > 0.f) {
                // Constant step (the reference default, --dt_gamma 0): u_{i+1} = fl(u_i + dt).  Inside one binade
                // [2^e, 2^(e+1)) every u is a multiple of ulp = 2^(e-23), so fl(u + dt) = u + D * ulp with D = dt / ulp
                // rounded to nearest -- the same D for the whole binade unless dt / ulp ends in exactly .5 (a tie, resolved
                // by the parity of u; then the serial chain below is used).  Lanes fill 32 lattice points per step with
                // integer arithmetic on the significand; the one add that crosses into the next binade is a real fp32 add.
                const uint32_t dbits = __float_as_uint(dt_c), Md = (dbits & 0x7FFFFFu) | 0x800000u;
                const int ed = (int)(dbits >> 23);            // biased exponent of dt (dt is a normal number)
                float tt = t;
                bool fallback = false;
                while (cnt < kWin && tt < far) {
                    const uint32_t tb = __float_as_uint(tt);
                    const int sh = (int)(tb >> 23) - ed;       // ulp(tt) / ulp(dt) = 2^sh, sh >= 0 because tt >= dt
                    uint32_t D;
                    if (sh == 0) D = Md;
                    else if (sh > 24) D = 0;
                    else {
                        const uint32_t rem = Md & ((1u << sh) - 1u), halfway = 1u << (sh - 1);
                        if (rem == halfway) { fallback = true; break; }
                        D = (Md >> sh) + (rem > halfway ? 1u : 0u);
                    }
                    if (D == 0) { fallback = true; break; }   // dt below half an ulp: t would stop advancing (degenerate)
                    const uint32_t U0 = (tb & 0x7FFFFFu) | 0x800000u;
                    uint32_t n_in = (0xFFFFFFu - U0) / D + 1u;           // lattice points of this binade, starting at tt
                    n_in = min(n_in, kWin - cnt);
                    uint32_t stored = 0;
                    for (uint32_t k0 = 0; k0 < n_in; k0 += 32) {
                        const uint32_t k = k0 + lane;
                        const float uk = __uint_as_float((tb & 0xFF800000u) | ((U0 + k * D) & 0x7FFFFFu));
                        const bool ok = k < n_in && uk < far;
                        if (ok) u[cnt + k] = uk;
                        const uint32_t m = __ballot_sync(0xffffffffu, ok);
                        stored += __popc(m);
                        if (m != 0xffffffffu) break;
                    }
                    cnt += stored;
                    if (stored < n_in) break;                  // reached `far` (or the window is full: stored == n_in then)
                    // last point of the binade + dt: the add that changes the exponent
                    const float ulast = __uint_as_float((tb & 0xFF800000u) | ((U0 + (n_in - 1) * D) & 0x7FFFFFu));
                    tt = ulast + dt_c;
                }
                if (fallback) cnt = 0;
            }
            if (cnt == 0) {
                // general case: every lane runs the same serial chain, lane i%32 keeps u_i
                float tt = t, mine = 0.f;
                while (cnt < kWin && tt < far) {
                    if ((cnt & 31u) == lane) mine = tt;
                    tt += clampf(tt * p.dt_gamma, p.dt_min, p.dt_max);
                    cnt++;
                    if ((cnt & 31u) == 0) u[cnt - 32 + lane] = mine;
                }
                if ((cnt & 31u) != 0 && lane < (cnt & 31u)) u[(cnt & ~31u) + lane] = mine;
            }
            __syncwarp();
            // (2) parallel probes
            for (uint32_t i = lane; i < cnt; i += 32) {
                Probe q;
                const float ti = u[i];
                uint32_t code;
                if (probe(p, r, ti, q)) {
                    code = 0x8000u | (i + 1);
                } else {
                    const float tt = voxel_exit(p, r, ti, q);
                    uint32_t lo = i + 1, hi = cnt;          // first j in (i, cnt) with u[j] >= tt, else cnt
                    if (p.dt_gamma == 0.f) {
                        // constant step: the target is near i + (tt - t_i) / dt; walk the few remaining entries
                        const float est = (tt - ti) / clampf(0.f, p.dt_min, p.dt_max);
                        uint32_t j = (est < (float)(cnt - i)) ? i + (uint32_t)fmaxf(est, 1.f) : cnt;
                        j = min(max(j, lo), hi);
                        while (j > lo && u[j - 1] >= tt) j--;
                        while (j < hi && u[j] < tt) j++;
                        lo = j;
                    } else {
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (u[mid] < tt) lo = mid + 1; else hi = mid;
                        }
                    }
                    code = lo;
                }
                nx[i] = (uint16_t)code;
            }
            __syncwarp();
            // (2b) pointer jumping: in free space the chase below follows ~80 skips per window, every hop a dependent shared-
            // memory read executed by all 32 lanes -- a third of the kernel's instructions.  Two rounds of "skip over the next
            // empty entry" (in parallel over the window) shorten those chains fourfold.  An entry is only skipped if it is empty
            // AND its own target stays inside the window: a keep entry must be visited, and the last hop of a chain that leaves
            // the window has to be seen by the chase (exit_skip).  Lanes may read an entry another lane is updating: either value
            // is a valid target further down the same chain.
#pragma unroll 1
            for (uint32_t round = 0; round < NGP_MARCH_JUMP_ROUNDS; round++) {
                for (uint32_t i = lane; i < cnt; i += 32) {
                    const uint32_t c = nx[i];
                    if (!(c & 0x8000u) && c < cnt) {
                        const uint32_t c2 = nx[c];
                        if (!(c2 & 0x8000u) && c2 < cnt) nx[i] = (uint16_t)c2;
                    }
                }
                __syncwarp();
            }
            // (3) pointer chase (uniform over the warp).  A run of consecutive kept lattice points is consumed in one
            // step: its length comes from a ballot over the 32 entries around i and its t's are stored by parallel lanes.
            uint32_t i = 0;
            int exit_skip = -1;                     // lattice index of a skip whose target lies beyond the window
            while (i < cnt && step < max_steps) {
                const uint32_t code = nx[i];                   // same address in every lane: one broadcast read
                if (!(code & 0x8000u)) {                        // empty voxel: follow the skip (the common case in free space)
                    if (code >= cnt) exit_skip = (int)i;
                    i = code;
                    continue;
                }
                const uint32_t c0 = i & ~31u;
                const uint32_t code_l = (c0 + lane < cnt) ? (uint32_t)nx[c0 + lane] : 0u;
                const uint32_t keep = __ballot_sync(0xffffffffu, (code_l & 0x8000u) != 0u) >> (i & 31u);
                uint32_t run = (keep == 0xffffffffu) ? 32u : (uint32_t)__ffs(~keep) - 1u;   // bits past the chunk end are 0
                run = min(run, max_steps - step);
                if (lane < run) tout[step + lane] = u[i + lane];
                step += run;
                i += run;
            }
            if (step >= max_steps) break;
            if (cnt < kWin) break;                  // the window reached `far`: the ray is finished
            // leave the window exactly as the reference loop would continue
            if (exit_skip >= 0) {
                // finish the skip serially: every u in the window is < tt, keep stepping from the last one
                Probe q;
                const float ts = u[exit_skip];
                probe(p, r, ts, q);
                const float tt = voxel_exit(p, r, ts, q);
                float tc = u[cnt - 1];
                do { tc += clampf(tc * p.dt_gamma, p.dt_min, p.dt_max); } while (tc < tt);
                t = tc;
            } else {
                // the last point of the window was kept: one more lattice step
                const float tl = u[cnt - 1];
                t = tl + clampf(tl * p.dt_gamma, p.dt_min, p.dt_max);
            }
            __syncwarp();
        }
        if (lane == 0) rays[(size_t)n * 2 + 1] = (int)step;
    }

    // last block to arrive turns the counts into ray-ordered offsets
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // Block-wide exclusive scan of the N counts: thread t owns the contiguous chunk [t * per, (t + 1) * per) so that all loads
    // of the block are in flight together (a single warp walking the array pays one L2 round trip per 32 rays).
    __shared__ uint32_t s_part[kCoopWarps * 32];
    const uint32_t T = kCoopWarps * 32, per = (N + T - 1) / T;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, N);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += (uint32_t)__ldcg(rays + (size_t)i * 2 + 1);
    s_part[threadIdx.x] = sum;
    __syncthreads();
    uint32_t carry = 0;
    for (uint32_t t = 0; t < threadIdx.x; t++) carry += s_part[t];        // T = 128 shared-memory reads
    uint32_t fit = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = (uint32_t)__ldcg(rays + (size_t)i * 2 + 1);
        rays[(size_t)i * 2] = (int)carry;
        carry += c;
        if (carry <= cap) fit = max(fit, carry);
    }
    __shared__ uint32_t s_fit[kCoopWarps];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) fit = max(fit, __shfl_xor_sync(0xffffffffu, fit, sft));
    if (lane == 0) s_fit[warp] = fit;
    __syncthreads();
    if (threadIdx.x == T - 1) {
        uint32_t f = 0;
        for (uint32_t w = 0; w < kCoopWarps; w++) f = max(f, s_fit[w]);
        counter[0] = (int)carry; counter[1] = 0;       // the last thread's running sum is the total
        if (cap) counter[2] = (int)f;
    }
}

// Pass 2 from the stored sample t's: one warp per ray, samples written in parallel (coalesced).
template <bool LDIR>
__global__ void __launch_bounds__(kCoopWarps * 32)
march_train_write_coop_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ rays_ldir,
                              const uint8_t* __restrict__ grid, float bound, bool contract, float dt_gamma, uint32_t max_steps,
                              uint32_t N, uint32_t C, uint32_t H, const int* __restrict__ rays, uint32_t M,
                              const int* __restrict__ m_dev, const float* __restrict__ t_scratch, float* __restrict__ xyzs,
                              float* __restrict__ dirs, float* __restrict__ ts, float* __restrict__ ldirs) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n = blockIdx.x * kCoopWarps + warp;
    if (n >= N) return;
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2), count = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
    if (count == 0 || offset + count > M) return;
    const MarchParams p = make_params(grid, bound, contract, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o, rays_d, n, false);
    float lx = 0, ly = 0, lz = 0;
    if (LDIR) { lx = __ldg(rays_ldir + (size_t)n * 3); ly = __ldg(rays_ldir + (size_t)n * 3 + 1); lz = __ldg(rays_ldir + (size_t)n * 3 + 2); }
    const float* tin = t_scratch + (size_t)n * max_steps;
    for (uint32_t k = lane; k < count; k += 32) {
        const float t = __ldg(tin + k);
        Probe q;
        probe(p, r, t, q);
        const size_t i = (size_t)offset + k;
        xyzs[i * 3] = q.cx; xyzs[i * 3 + 1] = q.cy; xyzs[i * 3 + 2] = q.cz;
        dirs[i * 3] = r.dx; dirs[i * 3 + 1] = r.dy; dirs[i * 3 + 2] = r.dz;
        *reinterpret_cast<float2*>(ts + i * 2) = make_float2(t + q.dt, q.dt);
        if (LDIR) { ldirs[i * 3] = lx; ldirs[i * 3 + 1] = ly; ldirs[i * 3 + 2] = lz; }
    }
}

// ---------------------------------------------------------------------------------------------------
// training composite: one warp per ray
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kCompThreads = 128;  // 4 rays per block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, uint32_t lane) {
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= (uint32_t)s) v += u;
    }
    return v;
}
__device__ __forceinline__ float warp_incl_prod(float v, uint32_t lane) {
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= (uint32_t)s) v *= u;
    }
    return v;
}

// raymarching.cu:519-597
__global__ void __launch_bounds__(kCompThreads)
composite_train_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                           const int* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh, float* __restrict__ weights,
                           float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2), count = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
    float r = 0, g = 0, b = 0, ws = 0, d = 0;
    if (count != 0 && offset + count <= M) {
        float T = 1.0f;  // transmittance entering the chunk
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t k = base + lane;
            const bool valid = k < count;
            const size_t i = (size_t)offset + k;
            float alpha = 0.f, tk = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
            if (valid) {
                const float2 tt = __ldg(reinterpret_cast<const float2*>(ts) + i);
                tk = tt.x;
                alpha = 1.0f - __expf(-__ldg(sigmas + i) * tt.y);
                cr = __ldg(rgbs + i * 3); cg = __ldg(rgbs + i * 3 + 1); cb = __ldg(rgbs + i * 3 + 2);
            }
            const float incl = warp_incl_prod(1.0f - alpha, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T_before = T * excl, T_after = T * incl;
            // the reference stops AFTER accumulating the first sample whose outgoing T < T_thresh
            const uint32_t dead = __ballot_sync(0xffffffffu, valid && T_after < T_thresh);
            const uint32_t last_live = dead ? (uint32_t)(__ffs(dead) - 1) : 31u;
            if (valid && lane <= last_live) {
                const float w = alpha * T_before;
                weights[i] = w;
                r += w * cr; g += w * cg; b += w * cb; ws += w; d += w * tk;
            }
            if (dead) break;
            T = __shfl_sync(0xffffffffu, T_after, 31);
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    }
    if (lane == 0) {
        weights_sum[n] = ws;
        depth[n] = d;
        image[(size_t)n * 3] = r; image[(size_t)n * 3 + 1] = g; image[(size_t)n * 3 + 2] = b;
    }
}

// raymarching.cu:623-712
__global__ void __launch_bounds__(kCompThreads)
composite_train_bwd_kernel(const float* __restrict__ grad_weights, const float* __restrict__ grad_weights_sum,
                           const float* __restrict__ grad_depth, const float* __restrict__ grad_image,
                           const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                           const int* __restrict__ rays, const float* __restrict__ weights_sum, const float* __restrict__ depth,
                           const float* __restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                           float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2), count = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
    if (count == 0 || offset + count > M) return;

    const float gi_r = __ldg(grad_image + (size_t)n * 3), gi_g = __ldg(grad_image + (size_t)n * 3 + 1), gi_b = __ldg(grad_image + (size_t)n * 3 + 2);
    const float g_ws = __ldg(grad_weights_sum + n), g_d = __ldg(grad_depth + n);
    const float r_fin = __ldg(image + (size_t)n * 3), g_fin = __ldg(image + (size_t)n * 3 + 1), b_fin = __ldg(image + (size_t)n * 3 + 2);
    const float ws_fin = __ldg(weights_sum + n), d_fin = __ldg(depth + n);

    float T = 1.0f, r0 = 0, g0 = 0, b0 = 0, ws0 = 0, d0 = 0;  // running prefixes entering the chunk
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t k = base + lane;
        const bool valid = k < count;
        const size_t i = (size_t)offset + k;
        float alpha = 0.f, tk = 0.f, dtk = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, gw = 0.f;
        if (valid) {
            const float2 tt = __ldg(reinterpret_cast<const float2*>(ts) + i);
            tk = tt.x; dtk = tt.y;
            alpha = 1.0f - __expf(-__ldg(sigmas + i) * dtk);
            cr = __ldg(rgbs + i * 3); cg = __ldg(rgbs + i * 3 + 1); cb = __ldg(rgbs + i * 3 + 2);
            gw = __ldg(grad_weights + i);
        }
        const float incl = warp_incl_prod(1.0f - alpha, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T_before = T * excl, T_after = T * incl;
        const uint32_t dead = __ballot_sync(0xffffffffu, valid && T_after < T_thresh);
        const uint32_t last_live = dead ? (uint32_t)(__ffs(dead) - 1) : 31u;
        const bool live = valid && lane <= last_live;
        const float w = live ? alpha * T_before : 0.f;
        const float pr = r0 + warp_incl_sum(w * cr, lane);
        const float pg = g0 + warp_incl_sum(w * cg, lane);
        const float pb = b0 + warp_incl_sum(w * cb, lane);
        const float pw = ws0 + warp_incl_sum(w, lane);
        const float pd = d0 + warp_incl_sum(w * tk, lane);
        if (live) {
            grad_rgbs[i * 3] = gi_r * w; grad_rgbs[i * 3 + 1] = gi_g * w; grad_rgbs[i * 3 + 2] = gi_b * w;
            grad_sigmas[i] = dtk * (gi_r * (T_after * cr - (r_fin - pr)) + gi_g * (T_after * cg - (g_fin - pg)) +
                                    gi_b * (T_after * cb - (b_fin - pb)) + (g_ws + gw) * (T_after - (ws_fin - pw)) +
                                    g_d * (T_after * tk - (d_fin - pd)));
        }
        if (dead) break;
        T = __shfl_sync(0xffffffffu, T_after, 31);
        r0 = __shfl_sync(0xffffffffu, pr, 31); g0 = __shfl_sync(0xffffffffu, pg, 31); b0 = __shfl_sync(0xffffffffu, pb, 31);
        ws0 = __shfl_sync(0xffffffffu, pw, 31); d0 = __shfl_sync(0xffffffffu, pd, 31);
    }
}


// Training composite + MSE loss + its backward in one pass per ray (one warp per ray):
//   image = composite + (1 - weights_sum) * bg            (renderer.py:553,672)
//   loss  = mean_rays mean_c (image - target)^2           (train_utils.py:540-541)
//   d image = loss_scale * 2 (image - target) / (3 N),  d weights_sum = -bg * sum_c d image_c
// followed by the backward recurrences of raymarching.cu:623-712 on the same samples (still in L1/L2).
// Every sample row < M of a ray that fits gets a gradient (zero past the early stop), so no memset is needed.
// The last block sums the per-ray losses in a fixed order (deterministic) into loss_out[0].
__global__ void __launch_bounds__(kCompThreads)
composite_train_mse_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                           const int* __restrict__ rays, uint32_t M, const int* __restrict__ m_dev, uint32_t N, float T_thresh,
                           float bg, const float* __restrict__ target, float loss_scale, float* __restrict__ image_out,
                           float* __restrict__ ray_loss, float* __restrict__ loss_out, int* __restrict__ ticket,
                           float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, int loss_mode,
                           const float* __restrict__ exposure, const ngp_loss_opts x) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    if (x.loss_scale_dev) loss_scale = __ldg(x.loss_scale_dev);      // dynamic GradScaler scale, updated on the device
    // live rays of the batch (adaptive ray count): rays at or beyond it carry no samples, no loss and no gradient
    const uint32_t n_live = x.n_rays_dev ? min(N, (uint32_t)max(__ldg(x.n_rays_dev), 1)) : N;
    if (n < N && n >= n_live) {
        if (lane == 0) {
            ray_loss[n] = 0.f;
            if (x.entropy_ray) x.entropy_ray[n] = 0.f;
            if (image_out) { image_out[(size_t)n * 3] = 0.f; image_out[(size_t)n * 3 + 1] = 0.f; image_out[(size_t)n * 3 + 2] = 0.f; }
            if (x.weights_sum_out) x.weights_sum_out[n] = 0.f;
            if (x.depth_out) x.depth_out[n] = 0.f;
        }
    } else if (n < N) {
        const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2), count = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
        const bool has = count != 0 && offset + count <= M;
        float r = 0, g = 0, b = 0, ws = 0, d = 0;
        // The kernel lasts as long as its longest ray (a warp walks its ray in chunks of 32 samples, one scan after the other), so
        // the loads of chunk c + 1 are issued before chunk c is scanned: the L2 round trip leaves the serial chain.
        struct Chunk { float2 tt; float sg, cr, cg, cb; };
        auto fetch = [&](uint32_t base) {
            Chunk c = {make_float2(0.f, 0.f), 0.f, 0.f, 0.f, 0.f};
            const uint32_t k = base + lane;
            if (k < count) {
                const size_t i = (size_t)offset + k;
                c.tt = __ldg(reinterpret_cast<const float2*>(ts) + i);
                c.sg = __ldg(sigmas + i);
                c.cr = __ldg(rgbs + i * 3); c.cg = __ldg(rgbs + i * 3 + 1); c.cb = __ldg(rgbs + i * 3 + 2);
            }
            return c;
        };
        if (has) {
            float T = 1.0f;
            Chunk nxt = fetch(0);
            for (uint32_t base = 0; base < count; base += 32) {
                const uint32_t k = base + lane;
                const bool valid = k < count;
                const Chunk cur = nxt;
                if (base + 32 < count) nxt = fetch(base + 32);
                float alpha = 0.f, tk = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
                if (valid) {
                    tk = cur.tt.x;
                    alpha = 1.0f - __expf(-cur.sg * cur.tt.y);
                    cr = cur.cr; cg = cur.cg; cb = cur.cb;
                }
                const float incl = warp_incl_prod(1.0f - alpha, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.0f;
                const float T_before = T * excl, T_after = T * incl;
                const uint32_t dead = __ballot_sync(0xffffffffu, valid && T_after < T_thresh);
                const uint32_t last_live = dead ? (uint32_t)(__ffs(dead) - 1) : 31u;
                if (valid && lane <= last_live) {
                    const float w = alpha * T_before;
                    r += w * cr; g += w * cg; b += w * cb; ws += w; d += w * tk;
                }
                if (dead) break;
                T = __shfl_sync(0xffffffffu, T_after, 31);
            }
            r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
        }
        // background: one scalar, or per ray (train_utils.py:495-496, background == 'random')
        float bgr = bg, bgg = bg, bgb = bg;
        if (x.bg_rays) { bgr = __ldg(x.bg_rays + (size_t)n * 3); bgg = __ldg(x.bg_rays + (size_t)n * 3 + 1); bgb = __ldg(x.bg_rays + (size_t)n * 3 + 2); }
        const float ir = r + (1.0f - ws) * bgr, ig = g + (1.0f - ws) * bgg, ib = b + (1.0f - ws) * bgb;
        // loss_mode 0: MSE (train_utils.py:540-541).  loss_mode 1: the clipped, tone-curve weighted MSE of the raw/HDR path
        // (train_utils.py:529-536): c = min(1, pred * exposure), loss = mean (c - gt)^2 / (1e-3 + stop_grad(c))^2
        float tr = __ldg(target + (size_t)n * 3), tgn = __ldg(target + (size_t)n * 3 + 1), tb = __ldg(target + (size_t)n * 3 + 2);
        if (x.target_alpha) {      // RGBA images: gt = rgb * a + bg * (1 - a)   (train_utils.py:504-505)
            const float a = __ldg(x.target_alpha + n);
            tr = tr * a + bgr * (1.f - a); tgn = tgn * a + bgg * (1.f - a); tb = tb * a + bgb * (1.f - a);
        }
        // per-channel weights of the loss: lossmult (Bayer mask of mosaiced raw data) and loss_weight (train_utils.py:515-536);
        // the normaliser is sum(lossmult) (there: lossmult_tensor.sum()), handed in as its reciprocal, else 3 * live rays
        float wr = 1.f, wgn = 1.f, wb = 1.f;
        if (x.lossmult) { wr = __ldg(x.lossmult + (size_t)n * 3); wgn = __ldg(x.lossmult + (size_t)n * 3 + 1); wb = __ldg(x.lossmult + (size_t)n * 3 + 2); }
        if (x.loss_weight) { wr *= __ldg(x.loss_weight + (size_t)n * 3); wgn *= __ldg(x.loss_weight + (size_t)n * 3 + 1); wb *= __ldg(x.loss_weight + (size_t)n * 3 + 2); }
        const float inv_norm = x.inv_norm_dev ? __ldg(x.inv_norm_dev) : 1.0f / (3.0f * (float)n_live);
        float er, eg, eb, dr = 1.f, dg = 1.f, db = 1.f;     // residuals and d loss_c / d image_c = 2 * w_c * e_c * d_c * inv_norm
        if (loss_mode == 1) {
            const float ex = exposure ? __ldg(exposure + n) : 1.f;
            const float pr_ = ir * ex, pg_ = ig * ex, pb_ = ib * ex;
            const float cr_ = fminf(1.f, pr_), cg_ = fminf(1.f, pg_), cb_ = fminf(1.f, pb_);
            const float sr = 1.f / (1e-3f + cr_), sg_ = 1.f / (1e-3f + cg_), sb = 1.f / (1e-3f + cb_);
            er = (cr_ - tr) * sr; eg = (cg_ - tgn) * sg_; eb = (cb_ - tb) * sb;
            // d min(1, p) / d p: 1 below the clip, 0 above, 1/2 exactly on it (torch.minimum splits the gradient at a tie; a ray
            // that sees only a white background with exposure 1 sits there)
            dr = (pr_ < 1.f) ? ex * sr : (pr_ == 1.f ? 0.5f * ex * sr : 0.f);
            dg = (pg_ < 1.f) ? ex * sg_ : (pg_ == 1.f ? 0.5f * ex * sg_ : 0.f);
            db = (pb_ < 1.f) ? ex * sb : (pb_ == 1.f ? 0.5f * ex * sb : 0.f);
        } else {
            er = ir - tr; eg = ig - tgn; eb = ib - tb;
        }
        // entropy regulariser on the opacity (train_utils.py:553-556): lambda * mean_n H2(clamp(ws, 1e-5, 1 - 1e-5))
        float ent = 0.f, g_ent = 0.f;
        if (x.lambda_entropy > 0.f) {
            const float wc = fminf(fmaxf(ws, 1e-5f), 1.0f - 1e-5f);
            ent = -wc * log2f(wc) - (1.f - wc) * log2f(1.f - wc);
            if (ws > 1e-5f && ws < 1.0f - 1e-5f) g_ent = x.lambda_entropy * (log2f(1.f - wc) - log2f(wc)) / (float)n_live;
        }
        if (lane == 0) {
            if (image_out) { image_out[(size_t)n * 3] = ir; image_out[(size_t)n * 3 + 1] = ig; image_out[(size_t)n * 3 + 2] = ib; }
            ray_loss[n] = (wr * er * er + wgn * eg * eg + wb * eb * eb) * inv_norm;
            if (x.entropy_ray) x.entropy_ray[n] = ent;
            if (x.weights_sum_out) x.weights_sum_out[n] = ws;
            if (x.depth_out) x.depth_out[n] = d;
        }
        if (has) {
            const float gs = loss_scale * 2.0f * inv_norm;
            const float gi_r = gs * wr * er * dr, gi_g = gs * wgn * eg * dg, gi_b = gs * wb * eb * db;
            const float g_ws = -(bgr * gi_r + bgg * gi_g + bgb * gi_b) + loss_scale * g_ent;
            float T = 1.0f, r0 = 0, g0 = 0, b0 = 0, ws0 = 0;
            uint32_t base = 0;
            Chunk nxt = fetch(0);
            for (; base < count; base += 32) {
                const uint32_t k = base + lane;
                const bool valid = k < count;
                const size_t i = (size_t)offset + k;
                const Chunk cur = nxt;
                if (base + 32 < count) nxt = fetch(base + 32);
                float alpha = 0.f, dtk = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
                if (valid) {
                    dtk = cur.tt.y;
                    alpha = 1.0f - __expf(-cur.sg * dtk);
                    cr = cur.cr; cg = cur.cg; cb = cur.cb;
                }
                const float incl = warp_incl_prod(1.0f - alpha, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.0f;
                const float T_before = T * excl, T_after = T * incl;
                const uint32_t dead = __ballot_sync(0xffffffffu, valid && T_after < T_thresh);
                const uint32_t last_live = dead ? (uint32_t)(__ffs(dead) - 1) : 31u;
                const bool live = valid && lane <= last_live;
                const float w = live ? alpha * T_before : 0.f;
                const float pr = r0 + warp_incl_sum(w * cr, lane);
                const float pg = g0 + warp_incl_sum(w * cg, lane);
                const float pb = b0 + warp_incl_sum(w * cb, lane);
                const float pw = ws0 + warp_incl_sum(w, lane);
                if (valid) {
                    float gsig = 0.f;
                    if (live)
                        gsig = dtk * (gi_r * (T_after * cr - (r - pr)) + gi_g * (T_after * cg - (g - pg)) +
                                      gi_b * (T_after * cb - (b - pb)) + g_ws * (T_after - (ws - pw)));
                    grad_rgbs[i * 3] = gi_r * w; grad_rgbs[i * 3 + 1] = gi_g * w; grad_rgbs[i * 3 + 2] = gi_b * w;
                    grad_sigmas[i] = gsig;
                }
                if (dead) { base += 32; break; }
                T = __shfl_sync(0xffffffffu, T_after, 31);
                r0 = __shfl_sync(0xffffffffu, pr, 31); g0 = __shfl_sync(0xffffffffu, pg, 31); b0 = __shfl_sync(0xffffffffu, pb, 31);
                ws0 = __shfl_sync(0xffffffffu, pw, 31);
            }
            for (uint32_t k = base + lane; k < count; k += 32) {   // samples after the early stop carry no gradient
                const size_t i = (size_t)offset + k;
                grad_rgbs[i * 3] = 0.f; grad_rgbs[i * 3 + 1] = 0.f; grad_rgbs[i * 3 + 2] = 0.f;
                grad_sigmas[i] = 0.f;
            }
        }
    }
    // last block: deterministic mean of the per-ray losses
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    float acc = 0.f, acc_e = 0.f;
    for (uint32_t i = lane; i < N; i += 32) acc += __ldcg(ray_loss + i);
    if (x.entropy_ray) for (uint32_t i = lane; i < N; i += 32) acc_e += __ldcg(x.entropy_ray + i);
    acc = warp_sum(acc);
    acc_e = warp_sum(acc_e);
    if (lane == 0) {
        const float e = x.lambda_entropy * acc_e / (float)n_live;
        loss_out[0] = acc + e;                 // ray_loss already carries the normaliser
        if (x.parts_out) { x.parts_out[0] = acc; x.parts_out[1] = e; }
        *ticket = 0;
    }
}

// Segmented sums of _march_rays_train.backward (raymarching/raymarching.py:319-329); one warp per ray.
__global__ void __launch_bounds__(kCompThreads)
march_train_bwd_kernel(const float* __restrict__ dL_dxyzs, const float* __restrict__ dL_ddirs, const float* __restrict__ ts,
                       const int* __restrict__ rays, uint32_t N, uint32_t M, float* __restrict__ dL_do, float* __restrict__ dL_dd) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 2);
    uint32_t count = (uint32_t)__ldg(rays + (size_t)n * 2 + 1);
    // a ray whose samples do not fit in the buffers was not rendered (composite_rays_train early-out, raymarching.cu:540-547;
    // only possible with the fixed-capacity buffers of FusedTrainStep): no gradient
    if ((uint64_t)offset + count > M) count = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
    for (uint32_t k = lane; k < count; k += 32) {
        const size_t i = (size_t)offset + k;
        const float gx = __ldg(dL_dxyzs + i * 3), gy = __ldg(dL_dxyzs + i * 3 + 1), gz = __ldg(dL_dxyzs + i * 3 + 2);
        const float t = __ldg(ts + i * 2);
        ox += gx; oy += gy; oz += gz;
        dx += gx * t; dy += gy * t; dz += gz * t;
        if (dL_ddirs) { dx += __ldg(dL_ddirs + i * 3); dy += __ldg(dL_ddirs + i * 3 + 1); dz += __ldg(dL_ddirs + i * 3 + 2); }
    }
    ox = warp_sum(ox); oy = warp_sum(oy); oz = warp_sum(oz);
    dx = warp_sum(dx); dy = warp_sum(dy); dz = warp_sum(dz);
    if (lane == 0) {
        dL_do[(size_t)n * 3] = ox; dL_do[(size_t)n * 3 + 1] = oy; dL_do[(size_t)n * 3 + 2] = oz;
        dL_dd[(size_t)n * 3] = dx; dL_dd[(size_t)n * 3 + 1] = dy; dL_dd[(size_t)n * 3 + 2] = dz;
    }
}

// ---------------------------------------------------------------------------------------------------
// inference march / composite
// ---------------------------------------------------------------------------------------------------

// raymarching.cu:730-846.  One thread per alive ray marches up to n_step samples; the samples of a warp's 32 rays form one
// contiguous block of each output ([n_alive, n_step, 3|3|2]), so they are staged in shared memory and written with
// coalesced stores -- per-thread stores at a stride of n_step samples cost one L2 transaction per 4-byte word, which is
// what bounded this kernel (8x the sector writes).
__global__ void __launch_bounds__(128)
march_infer_kernel(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, const float* __restrict__ rays_t,
                   const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, bool contract, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ nears,
                   const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ ts,
                   const float* __restrict__ noises, bool staged, const int* __restrict__ ctl) {
    extern __shared__ float s_stage[];
    // ctl (device-driven loop, ngp_march_rays_dev): the live ray count and the steps per ray of this iteration come from
    // device memory, the launch was sized for an upper bound of n_alive
    if (ctl) { n_alive = (uint32_t)__ldg(ctl); n_step = (uint32_t)__ldg(ctl + 1); }
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n0 = n - lane;                           // first ray of the warp
    if (n0 >= n_alive) return;                              // whole warp
    const bool active = n < n_alive;
    // staged: odd per-lane strides in shared memory (no bank conflicts); otherwise (n_step too large for shared memory) every
    // thread writes its own rows of the outputs directly
    const uint32_t sx_stride = staged ? 3 * n_step + 1 : 3 * n_step, st_stride = staged ? 2 * n_step + 1 : 2 * n_step;
    float* sx = staged ? s_stage + wid * 32 * (sx_stride + st_stride) : xyzs + (size_t)n0 * sx_stride;
    float* st = staged ? sx + 32 * sx_stride : ts + (size_t)n0 * st_stride;
    (void)nears;
    uint32_t step = 0;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (active) {
        const int index = __ldg(rays_alive + n);
        const MarchParams p = make_params(grid, bound, contract, dt_gamma, max_steps, C, H);
        const Ray r = load_ray(rays_o, rays_d, (size_t)index, true);
        dx = r.dx; dy = r.dy; dz = r.dz;
        const float far = __ldg(fars + index);
        float t = __ldg(rays_t + index);
        t += clampf(t * p.dt_gamma, p.dt_min, p.dt_max) * __ldg(noises + n);
        float* px = sx + lane * sx_stride;
        float* pt = st + lane * st_stride;
        while (t < far && step < n_step) {
            Probe q;
            if (probe(p, r, t, q)) {
                px[0] = q.cx; px[1] = q.cy; px[2] = q.cz;
                t += q.dt;
                pt[0] = t; pt[1] = q.dt;
                px += 3; pt += 2;
                step++;
            } else {
                t = skip_voxel(p, r, t, q);
            }
        }
        // unwritten tail: ts[0] == 0 is the "ray finished" sentinel read by composite_rays (raymarching.cu:893-894)
        for (uint32_t k = step; k < n_step; k++) {
            px[0] = px[1] = px[2] = 0.f;
            pt[0] = pt[1] = 0.f;
            px += 3; pt += 2;
        }
    }
    if (!staged) {
        if (active) {
            float* pd = dirs + (size_t)n * 3 * n_step;
            for (uint32_t k = 0; k < n_step; k++) {
                pd[3 * k] = k < step ? dx : 0.f; pd[3 * k + 1] = k < step ? dy : 0.f; pd[3 * k + 2] = k < step ? dz : 0.f;
            }
        }
        return;
    }
    __syncwarp();
    const uint32_t rays_here = min(32u, n_alive - n0);
    {   // xyzs and dirs: 3 * n_step words per ray
        const uint32_t per = 3 * n_step, total = rays_here * per;
        float* gx = xyzs + (size_t)n0 * per;
        float* gd = dirs + (size_t)n0 * per;
        for (uint32_t i = lane; i < ((total + 31u) & ~31u); i += 32) {
            const uint32_t ray = min(i / per, 31u), rem = i - (i / per) * per;
            const uint32_t k = rem / 3, c = rem - k * 3;
            const uint32_t cnt = __shfl_sync(0xffffffffu, step, ray);
            const float d0 = __shfl_sync(0xffffffffu, dx, ray), d1 = __shfl_sync(0xffffffffu, dy, ray), d2 = __shfl_sync(0xffffffffu, dz, ray);
            if (i < total) {
                gx[i] = sx[ray * sx_stride + rem];
                gd[i] = (k < cnt) ? (c == 0 ? d0 : (c == 1 ? d1 : d2)) : 0.f;
            }
        }
    }
    {   // ts: 2 * n_step words per ray
        const uint32_t per = 2 * n_step, total = rays_here * per;
        float* gt = ts + (size_t)n0 * per;
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t ray = i / per, rem = i - ray * per;
            gt[i] = st[ray * st_stride + rem];
        }
    }
}

// raymarching.cu:859-941
__global__ void __launch_bounds__(128)
composite_infer_kernel(uint32_t n_alive, uint32_t n_step, float T_thresh, int* __restrict__ rays_alive, float* __restrict__ rays_t,
                       const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ ts,
                       float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
                       const int* __restrict__ ctl) {
    if (ctl) { n_alive = (uint32_t)__ldg(ctl); n_step = (uint32_t)__ldg(ctl + 1); }
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float* sg = sigmas + (size_t)n * n_step;
    const float* cl = rgbs + (size_t)n * n_step * 3;
    const float* tp = ts + (size_t)n * n_step * 2;
    float t = 0.f;
    float d = depth[index], r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    float wsum = weights_sum[index];
    uint32_t step = 0;
    while (step < n_step) {
        const float2 tt = __ldg(reinterpret_cast<const float2*>(tp) + step);
        if (tt.x == 0) break;  // finished ray
        const float alpha = 1.0f - __expf(-__ldg(sg + step) * tt.y);
        const float T = 1 - wsum;
        const float w = alpha * T;
        wsum += w;
        t = tt.x;
        d += w * t;
        r += w * __ldg(cl + step * 3); g += w * __ldg(cl + step * 3 + 1); b += w * __ldg(cl + step * 3 + 2);
        if (T < T_thresh) break;
        step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = wsum;
    depth[index] = d;
    image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
}

// Ordered compaction of the surviving ray ids by a single block (inference batches are <= a few million ids:
// 8 MB read, one pass).  Replaces `rays_alive[rays_alive >= 0]` (nerf/renderer.py:612).
__global__ void __launch_bounds__(1024)
compact_alive_kernel(const int* __restrict__ rays_alive, uint32_t n_alive, int* __restrict__ out, int* __restrict__ n_out) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_alive; base += 1024) {
        const uint32_t i = base + tid;
        const int v = (i < n_alive) ? __ldg(rays_alive + i) : -1;
        const bool keep = v >= 0;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        uint32_t wsum = s_warp[lane];   // warp 0..31 totals
        uint32_t incl = wsum;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= (uint32_t)s) incl += u;
        }
        const uint32_t warp_excl = __shfl_sync(0xffffffffu, incl - wsum, wid);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t carry = s_carry;
        if (keep) out[carry + warp_excl + __popc(m & ((1u << lane) - 1))] = v;
        __syncthreads();
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (tid == 0) n_out[0] = (int)s_carry;
}

// The same over many blocks (a 1080p frame starts with 2 M ids): every block owns 4096 consecutive ids, 16 per thread.
// Pass 1 counts the survivors per block; pass 2 sums the counts of the blocks before it (<= 512 values), scans its own
// threads and writes -- order preserved, two launches, no spinning.
constexpr uint32_t kCompactThreads = 256, kCompactPer = 16, kCompactTile = kCompactThreads * kCompactPer;

__device__ __forceinline__ uint32_t compact_load(const int* __restrict__ rays_alive, uint32_t n_alive, uint32_t first, int (&v)[kCompactPer]) {
    uint32_t cnt = 0;
    if (first + kCompactPer <= n_alive) {
#pragma unroll
        for (uint32_t q = 0; q < kCompactPer / 4; q++) {
            const int4 u = __ldg(reinterpret_cast<const int4*>(rays_alive + first) + q);
            v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
        }
    } else {
#pragma unroll
        for (uint32_t k = 0; k < kCompactPer; k++) v[k] = (first + k < n_alive) ? __ldg(rays_alive + first + k) : -1;
    }
#pragma unroll
    for (uint32_t k = 0; k < kCompactPer; k++) cnt += v[k] >= 0;
    return cnt;
}

// exclusive prefix of `cnt` over the block's threads; total in *block_total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t cnt, uint32_t* s_warp, uint32_t* block_total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= (uint32_t)s) incl += u;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < kCompactThreads / 32; w++) {
        const uint32_t c = s_warp[w];
        if (w < wid) before += c;
        total += c;
    }
    *block_total = total;
    return before + incl - cnt;
}

__global__ void __launch_bounds__(kCompactThreads)
compact_count_kernel(const int* __restrict__ rays_alive, uint32_t n_alive, int* __restrict__ block_counts, const int* __restrict__ ctl) {
    __shared__ uint32_t s_warp[kCompactThreads / 32];
    if (ctl) n_alive = (uint32_t)__ldg(ctl);
    int v[kCompactPer];
    const uint32_t cnt = compact_load(rays_alive, n_alive, blockIdx.x * kCompactTile + threadIdx.x * kCompactPer, v);
    uint32_t total;
    block_exclusive_scan(cnt, s_warp, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (int)total;
}

__global__ void __launch_bounds__(kCompactThreads)
compact_write_kernel(const int* __restrict__ rays_alive, uint32_t n_alive, const int* __restrict__ block_counts, int* __restrict__ out,
                     int* __restrict__ n_out, const int* __restrict__ ctl, int* __restrict__ ctl_next, uint32_t row_cap, uint32_t step_cap,
                     uint32_t max_steps) {
    __shared__ uint32_t s_warp[kCompactThreads / 32];
    __shared__ uint32_t s_red[kCompactThreads / 32];
    if (ctl) n_alive = (uint32_t)__ldg(ctl);
    // survivors in the blocks before this one
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kCompactThreads) part += (uint32_t)__ldg(block_counts + b);
    part = __reduce_add_sync(0xffffffffu, part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    int v[kCompactPer];
    const uint32_t cnt = compact_load(rays_alive, n_alive, blockIdx.x * kCompactTile + threadIdx.x * kCompactPer, v);
    uint32_t total;
    uint32_t pos = block_exclusive_scan(cnt, s_warp, &total);      // (its __syncthreads also publishes s_red)
    uint32_t base = 0;
#pragma unroll
    for (uint32_t w = 0; w < kCompactThreads / 32; w++) base += s_red[w];
    pos += base;
#pragma unroll
    for (uint32_t k = 0; k < kCompactPer; k++)
        if (v[k] >= 0) out[pos++] = v[k];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const uint32_t alive = base + total;
        n_out[0] = (int)alive;
        if (ctl_next) {
            // the loop header of run_cuda (renderer.py:588-598, 616) for the NEXT iteration: step += n_step; stop at max_steps or
            // when no ray is left; n_step = max(min(row_cap // n_alive, step_cap), 1) (the reference: row_cap = N, step_cap = 8).
            // Written to the other control block: the blocks of this launch still read the current one.
            const uint32_t step = (uint32_t)__ldg(ctl + 3) + (uint32_t)__ldg(ctl + 1);
            const uint32_t n_next = (step >= max_steps) ? 0u : alive;
            const uint32_t ns = n_next ? max(min(row_cap / n_next, step_cap), 1u) : 1u;
            ctl_next[0] = (int)n_next; ctl_next[1] = (int)ns; ctl_next[2] = (int)(n_next * ns); ctl_next[3] = (int)step;
        }
    }
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                      float min_near, float* nears, float* fars, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays_o || !rays_d || !aabb || !nears || !fars) return NGP_ERR_NULL;
    near_far_kernel<<<div_up(N, 128u), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return finish_launch();
}

extern "C" int ngp_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                                ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays_o || !rays_d || !coords) return NGP_ERR_NULL;
    sph_from_ray_kernel<<<div_up(N, 128u), 128, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    return finish_launch();
}

extern "C" int ngp_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!coords || !indices) return NGP_ERR_NULL;
    morton3d_kernel<<<div_up(N, 256u), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    return finish_launch();
}

extern "C" int ngp_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!coords || !indices) return NGP_ERR_NULL;
    morton3d_invert_kernel<<<div_up(N, 256u), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    return finish_launch();
}

extern "C" int ngp_packbits(const float* grid, uint32_t N, float density_thresh, const float* thresh_dev,
                            uint8_t* bitfield, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!grid || !bitfield) return NGP_ERR_NULL;
    if (!aligned(grid, 16)) return NGP_ERR_ALIGN;
    packbits_kernel<<<div_up(N, 256u), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, thresh_dev, bitfield);
    return finish_launch();
}

extern "C" int ngp_flatten_rays(const int32_t* rays, uint32_t N, uint32_t M, int32_t* res, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays || !res) return NGP_ERR_NULL;
    flatten_rays_kernel<<<div_up(N * 32u, 128u), 128, 0, (cudaStream_t)stream>>>(rays, N, M, res);
    return finish_launch();
}

extern "C" int ngp_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                          int contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                                          uint32_t H, const float* nears, const float* fars, const float* noises,
                                          int32_t* rays, int32_t* counter, float* t_scratch, ngp_stream_t stream) {
    if (!counter) return NGP_ERR_NULL;
    if (N == 0) return NGP_OK;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !noises || !rays) return NGP_ERR_NULL;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024) return NGP_ERR_BAD_ARG;
    if (t_scratch)
        march_train_count_coop_kernel<false><<<div_up(N, kCoopWarps), kCoopWarps * 32, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, t_scratch,
            nullptr, 0.f, nullptr, nullptr, 0u, nullptr, nullptr);
    else
        march_train_count_kernel<<<div_up(N, kMarchThreads), kMarchThreads, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter);
    return finish_launch();
}

extern "C" int ngp_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                               const uint8_t* grid, float bound, int contract, float dt_gamma,
                                               uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* noises,
                                               uint32_t cap, float* nears_out, float* fars_out, int32_t* rays,
                                               int32_t* counter, float* t_scratch, ngp_stream_t stream) {
    if (!counter) return NGP_ERR_NULL;
    if (N == 0) return NGP_OK;
    if (!rays_o || !rays_d || !aabb || !grid || !noises || !rays || !t_scratch) return NGP_ERR_NULL;
    if ((nears_out != nullptr) != (fars_out != nullptr)) return NGP_ERR_NULL;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024 || cap == 0) return NGP_ERR_BAD_ARG;
    march_train_count_coop_kernel<true><<<div_up(N, kCoopWarps), kCoopWarps * 32, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nullptr, nullptr, noises, rays, counter,
        t_scratch, aabb, min_near, nears_out, fars_out, cap, nullptr, nullptr);
    return finish_launch();
}

extern "C" int ngp_march_rays_train_count_ex(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                             const float* cam_near_far, const int32_t* n_rays_dev, const uint8_t* grid, float bound,
                                             int contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                             const float* noises, uint32_t cap, float* nears_out, float* fars_out, int32_t* rays,
                                             int32_t* counter, float* t_scratch, ngp_stream_t stream) {
    if (!counter) return NGP_ERR_NULL;
    if (N == 0) return NGP_OK;
    if (!rays_o || !rays_d || !aabb || !grid || !noises || !rays || !t_scratch) return NGP_ERR_NULL;
    if ((nears_out != nullptr) != (fars_out != nullptr)) return NGP_ERR_NULL;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024 || cap == 0) return NGP_ERR_BAD_ARG;
    if (cam_near_far && !aligned(cam_near_far, 4)) return NGP_ERR_ALIGN;
    march_train_count_coop_kernel<true><<<div_up(N, kCoopWarps), kCoopWarps * 32, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nullptr, nullptr, noises, rays, counter,
        t_scratch, aabb, min_near, nears_out, fars_out, cap, cam_near_far, n_rays_dev);
    return finish_launch();
}

extern "C" int ngp_march_rays_train_write(const float* rays_o, const float* rays_d, const float* rays_ldir,
                                          const uint8_t* grid, float bound, int contract, float dt_gamma,
                                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears,
                                          const float* fars, const float* noises, const int32_t* rays, uint32_t M,
                                          const int32_t* m_dev, const float* t_scratch, float* xyzs, float* dirs, float* ts,
                                          float* ldirs, ngp_stream_t stream) {
    if (N == 0 || M == 0) return NGP_OK;
    if (!rays_o || !rays_d || !grid || !rays || !xyzs || !dirs || !ts) return NGP_ERR_NULL;
    if (!t_scratch && (!nears || !fars || !noises || m_dev)) return NGP_ERR_NULL;
    if ((rays_ldir != nullptr) != (ldirs != nullptr)) return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024) return NGP_ERR_BAD_ARG;
    if (t_scratch) {
        const uint32_t cb = div_up(N, kCoopWarps);
        if (rays_ldir)
            march_train_write_coop_kernel<true><<<cb, kCoopWarps * 32, 0, (cudaStream_t)stream>>>(
                rays_o, rays_d, rays_ldir, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, rays, M, m_dev, t_scratch, xyzs, dirs, ts, ldirs);
        else
            march_train_write_coop_kernel<false><<<cb, kCoopWarps * 32, 0, (cudaStream_t)stream>>>(
                rays_o, rays_d, rays_ldir, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, rays, M, m_dev, t_scratch, xyzs, dirs, ts, ldirs);
        return finish_launch();
    }
    const uint32_t blocks = div_up(N, kMarchThreads);
    if (rays_ldir)
        march_train_write_kernel<true><<<blocks, kMarchThreads, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, rays_ldir, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, M, xyzs, dirs, ts, ldirs);
    else
        march_train_write_kernel<false><<<blocks, kMarchThreads, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, rays_ldir, grid, bound, contract != 0, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, M, xyzs, dirs, ts, ldirs);
    return finish_launch();
}

extern "C" int ngp_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts,
                                                const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                                float* weights, float* weights_sum, float* depth, float* image,
                                                ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays || !weights_sum || !depth || !image) return NGP_ERR_NULL;
    if (M > 0 && (!sigmas || !rgbs || !ts || !weights)) return NGP_ERR_NULL;
    if (M > 0 && !aligned(ts, 8)) return NGP_ERR_ALIGN;
    composite_train_fwd_kernel<<<div_up(N * 32u, kCompThreads), kCompThreads, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, ts, rays, M, N, T_thresh, weights, weights_sum, depth, image);
    return finish_launch();
}

extern "C" int ngp_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum,
                                                 const float* grad_depth, const float* grad_image,
                                                 const float* sigmas, const float* rgbs, const float* ts,
                                                 const int32_t* rays, const float* weights_sum, const float* depth,
                                                 const float* image, uint32_t M, uint32_t N, float T_thresh,
                                                 float* grad_sigmas, float* grad_rgbs, ngp_stream_t stream) {
    if (N == 0 || M == 0) return NGP_OK;
    if (!grad_weights || !grad_weights_sum || !grad_depth || !grad_image || !sigmas || !rgbs || !ts || !rays ||
        !weights_sum || !depth || !image || !grad_sigmas || !grad_rgbs)
        return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    composite_train_bwd_kernel<<<div_up(N * 32u, kCompThreads), kCompThreads, 0, (cudaStream_t)stream>>>(
        grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays, weights_sum, depth, image, M, N,
        T_thresh, grad_sigmas, grad_rgbs);
    return finish_launch();
}

extern "C" int ngp_composite_train_mse(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays, uint32_t M,
                                       const int32_t* m_dev, uint32_t N, float T_thresh, float bg_color, const float* target,
                                       float loss_scale, float* image_out, float* ray_loss, float* loss_out, int32_t* ticket,
                                       float* grad_sigmas, float* grad_rgbs, int loss_mode, const float* exposure,
                                       ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays || !target || !ray_loss || !loss_out || !ticket) return NGP_ERR_NULL;
    if (M > 0 && (!sigmas || !rgbs || !ts || !grad_sigmas || !grad_rgbs)) return NGP_ERR_NULL;
    if (M > 0 && !aligned(ts, 8)) return NGP_ERR_ALIGN;
    if (loss_mode < 0 || loss_mode > 1) return NGP_ERR_BAD_ARG;
    ngp_loss_opts none = {};
    composite_train_mse_kernel<<<div_up(N * 32u, kCompThreads), kCompThreads, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, ts, rays, M, m_dev, N, T_thresh, bg_color, target, loss_scale, image_out, ray_loss, loss_out, ticket,
        grad_sigmas, grad_rgbs, loss_mode, exposure, none);
    return finish_launch();
}

extern "C" int ngp_composite_train_loss(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays, uint32_t M,
                                        const int32_t* m_dev, uint32_t N, float T_thresh, float bg_color, const float* target,
                                        float loss_scale, float* image_out, float* ray_loss, float* loss_out, int32_t* ticket,
                                        float* grad_sigmas, float* grad_rgbs, int loss_mode, const float* exposure,
                                        const ngp_loss_opts* opts, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays || !target || !ray_loss || !loss_out || !ticket) return NGP_ERR_NULL;
    if (M > 0 && (!sigmas || !rgbs || !ts || !grad_sigmas || !grad_rgbs)) return NGP_ERR_NULL;
    if (M > 0 && !aligned(ts, 8)) return NGP_ERR_ALIGN;
    if (loss_mode < 0 || loss_mode > 1) return NGP_ERR_BAD_ARG;
    ngp_loss_opts x = {};
    if (opts) x = *opts;
    if (x.lambda_entropy < 0.f || (x.lambda_entropy > 0.f && !x.entropy_ray)) return NGP_ERR_BAD_ARG;
    composite_train_mse_kernel<<<div_up(N * 32u, kCompThreads), kCompThreads, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, ts, rays, M, m_dev, N, T_thresh, bg_color, target, loss_scale, image_out, ray_loss, loss_out, ticket,
        grad_sigmas, grad_rgbs, loss_mode, exposure, x);
    return finish_launch();
}

extern "C" int ngp_march_rays_train_backward(const float* dL_dxyzs, const float* dL_ddirs, const float* ts,
                                             const int32_t* rays, uint32_t N, uint32_t M, float* dL_drays_o,
                                             float* dL_drays_d, ngp_stream_t stream) {
    if (N == 0) return NGP_OK;
    if (!rays || !dL_drays_o || !dL_drays_d) return NGP_ERR_NULL;
    if (M > 0 && (!dL_dxyzs || !ts)) return NGP_ERR_NULL;
    march_train_bwd_kernel<<<div_up(N * 32u, kCompThreads), kCompThreads, 0, (cudaStream_t)stream>>>(
        dL_dxyzs, dL_ddirs, ts, rays, N, M, dL_drays_o, dL_drays_d);
    return finish_launch();
}

extern "C" int ngp_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                              const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                              uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                              const float* fars, float* xyzs, float* dirs, float* ts, const float* noises,
                              ngp_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return NGP_OK;
    if (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !fars || !xyzs || !dirs || !ts || !noises) return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024) return NGP_ERR_BAD_ARG;
    uint32_t stage_bytes = 4u * 32u * (5u * n_step + 2u) * (uint32_t)sizeof(float);      // 4 warps per block
    const bool staged = stage_bytes <= 48u * 1024u;                                    // n_step <= 19 (the renderer uses <= 8)
    if (!staged) stage_bytes = 0;
    march_infer_kernel<<<div_up(n_alive, 128u), 128, stage_bytes, (cudaStream_t)stream>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, contract != 0, dt_gamma, max_steps, C, H, grid, nears, fars,
        xyzs, dirs, ts, noises, staged, nullptr);
    return finish_launch();
}

extern "C" int ngp_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                  const float* sigmas, const float* rgbs, const float* ts, float* weights_sum,
                                  float* depth, float* image, ngp_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return NGP_OK;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !ts || !weights_sum || !depth || !image) return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    composite_infer_kernel<<<div_up(n_alive, 128u), 128, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, nullptr);
    return finish_launch();
}

extern "C" int ngp_compact_rays_alive(const int32_t* rays_alive, uint32_t n_alive, int32_t* alive_out, int32_t* n_out,
                                      int32_t* workspace, ngp_stream_t stream) {
    if (!n_out) return NGP_ERR_NULL;
    if (n_alive > 0 && (!rays_alive || !alive_out)) return NGP_ERR_NULL;
    if (!workspace || n_alive <= kCompactTile || !aligned(rays_alive, 16)) {
        compact_alive_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rays_alive, n_alive, alive_out, n_out);
        return finish_launch();
    }
    const uint32_t blocks = div_up(n_alive, kCompactTile);
    compact_count_kernel<<<blocks, kCompactThreads, 0, (cudaStream_t)stream>>>(rays_alive, n_alive, workspace, nullptr);
    compact_write_kernel<<<blocks, kCompactThreads, 0, (cudaStream_t)stream>>>(rays_alive, n_alive, workspace, alive_out, n_out, nullptr, nullptr, 0, 0, 0);
    return finish_launch();
}

constexpr uint32_t kMaxDevStep = 16;     // largest n_step of the device-driven loop (shared-memory staging of the marcher)

// ---- the alive-ray loop of run_cuda (renderer.py:588-616) driven from the device ----------------------------------------------
// ctl = int32[4] {n_alive, n_step, n_alive * n_step, step}: every kernel of an iteration reads its sizes from ctl, the compaction
// writes the next iteration's block to ctl_next.  Launches are sized for `n_alive_bound` >= the true count (the host's last
// known value: the count never grows), so the host never has to wait for the device to size a launch.
extern "C" int ngp_march_rays_dev(const int32_t* ctl, uint32_t n_alive_bound, const int32_t* rays_alive, const float* rays_t,
                                  const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                                  uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* fars, float* xyzs,
                                  float* dirs, float* ts, const float* noises, ngp_stream_t stream) {
    if (n_alive_bound == 0) return NGP_OK;
    if (!ctl || !rays_alive || !rays_t || !rays_o || !rays_d || !grid || !fars || !xyzs || !dirs || !ts || !noises) return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    if (max_steps == 0 || H == 0 || C == 0 || H > 1024) return NGP_ERR_BAD_ARG;
    const uint32_t stage_bytes = 4u * 32u * (5u * kMaxDevStep + 2u) * (uint32_t)sizeof(float);      // n_step <= kMaxDevStep: 43 KB
    march_infer_kernel<<<div_up(n_alive_bound, 128u), 128, stage_bytes, (cudaStream_t)stream>>>(
        n_alive_bound, 1, rays_alive, rays_t, rays_o, rays_d, bound, contract != 0, dt_gamma, max_steps, C, H, grid, nullptr, fars,
        xyzs, dirs, ts, noises, true, ctl);
    return finish_launch();
}

extern "C" int ngp_composite_rays_dev(const int32_t* ctl, uint32_t n_alive_bound, float T_thresh, int32_t* rays_alive, float* rays_t,
                                      const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth,
                                      float* image, ngp_stream_t stream) {
    if (n_alive_bound == 0) return NGP_OK;
    if (!ctl || !rays_alive || !rays_t || !sigmas || !rgbs || !ts || !weights_sum || !depth || !image) return NGP_ERR_NULL;
    if (!aligned(ts, 8)) return NGP_ERR_ALIGN;
    composite_infer_kernel<<<div_up(n_alive_bound, 128u), 128, 0, (cudaStream_t)stream>>>(
        n_alive_bound, 1, T_thresh, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, ctl);
    return finish_launch();
}

extern "C" int ngp_compact_rays_alive_dev(const int32_t* ctl, int32_t* ctl_next, uint32_t n_alive_bound, uint32_t row_cap,
                                          uint32_t step_cap, uint32_t max_steps, const int32_t* rays_alive, int32_t* alive_out,
                                          int32_t* n_out, int32_t* workspace, ngp_stream_t stream) {
    if (!ctl || !ctl_next || !n_out || !workspace || !rays_alive || !alive_out) return NGP_ERR_NULL;
    if (!aligned(rays_alive, 16) || row_cap == 0 || max_steps == 0 || step_cap == 0 || step_cap > kMaxDevStep) return NGP_ERR_BAD_ARG;
    const uint32_t blocks = std::max(1u, div_up(n_alive_bound, kCompactTile));
    compact_count_kernel<<<blocks, kCompactThreads, 0, (cudaStream_t)stream>>>(rays_alive, n_alive_bound, workspace, ctl);
    compact_write_kernel<<<blocks, kCompactThreads, 0, (cudaStream_t)stream>>>(rays_alive, n_alive_bound, workspace, alive_out, n_out, ctl,
                                                                             ctl_next, row_cap, step_cap, max_steps);
    return finish_launch();
}

namespace ngp {
namespace {
__global__ void adaptive_num_rays_kernel(int* __restrict__ n_rays, const int* __restrict__ m_dev, uint32_t target_points, uint32_t n_max) {
    const int m = *m_dev, n = *n_rays;
    if (m <= 0) return;                                // no sample at all: keep the count (the reference would divide by zero)
    // int(round((num_points / outputs['num_points']) * num_rays)) in double like python, banker's rounding like round()
    const double v = ((double)target_points / (double)m) * (double)n;
    const long long r = llrint(v);
    *n_rays = (int)max(1ll, min((long long)n_max, r));
}
}  // namespace
// Uniform [0, 1) floats from a counter-based generator (splitmix64 of seed, call counter, element index; 24 random bits each).
// The per-ray jitter of the marcher and the random background colours are drawn inside the captured step: torch's own
// generator costs two eager int64 fills in front of every graph replay (seed / offset of its Philox state).
__global__ void __launch_bounds__(1024)
uniform_kernel(float* __restrict__ out, uint32_t n, uint64_t seed, int* __restrict__ counter, bool advance) {
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = (uint32_t)*counter;
    __syncthreads();
    const uint64_t base = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)s_c + 1ull);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t z = base + 0xD1B54A32D192ED03ull * ((uint64_t)i + 1ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        out[i] = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
    }
    if (advance && threadIdx.x == 0) *counter = (int)(s_c + 1u);      // single-block launches only (every thread has read s_c)
}
__global__ void counter_advance_kernel(int* __restrict__ counter) { *counter += 1; }

}  // namespace ngp

extern "C" int ngp_uniform(float* out, uint32_t n, uint64_t seed, int32_t* counter_dev, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!out || !counter_dev) return NGP_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= (1u << 16)) {
        ngp::uniform_kernel<<<1, 1024, 0, st>>>(out, n, seed, counter_dev, true);
    } else {
        ngp::uniform_kernel<<<std::min<uint32_t>(ngp::div_up(n, 1024u), 4u * ngp::kNumSMs), 1024, 0, st>>>(out, n, seed, counter_dev, false);
        ngp::counter_advance_kernel<<<1, 1, 0, st>>>(counter_dev);
    }
    return ngp::finish_launch();
}

extern "C" int ngp_adaptive_num_rays(int32_t* n_rays_dev, const int32_t* m_dev, uint32_t target_points, uint32_t n_max,
                                     ngp_stream_t stream) {
    if (!n_rays_dev || !m_dev) return NGP_ERR_NULL;
    if (n_max == 0 || target_points == 0) return NGP_ERR_BAD_ARG;
    ngp::adaptive_num_rays_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(n_rays_dev, m_dev, target_points, n_max);
    return ngp::finish_launch();
}
