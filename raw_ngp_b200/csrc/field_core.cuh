// field_core.cuh -- device helpers shared by the fused NeRF field kernels (field.cu, field_ws.cu): per-level index
// constants, corner rows / weights, the branch-free gather of one level, MMA plans.
#pragma once
#include "common.cuh"
#include "grid_core.cuh"
#include "mlp_core.cuh"
#include "tcgen05.cuh"

// measured at configs[1]: backward 317 us without, 362 us with the 16-byte form below (zero-padded slots are not free in L2)
#ifndef NGP_SCATTER_V4
#define NGP_SCATTER_V4 0
#endif

namespace ngp {
namespace fieldcore {

using namespace mlpcore;
using namespace gridcore;

constexpr uint32_t kMaxLevels = 32;

struct GridArgs {
    const __half* table;        // [sO, 2] fp16
    const int* offsets;         // [L+1]
    const float* feat_weights;  // [2L] or nullptr (BARF annealing window, network.py:99-109)
    float S, bound;
    uint32_t H, L, gridtype, interp;
    bool align_corners;
};

// Per-level constants, computed once per CTA into shared memory.  The index map of gridencoder.cu:61-79 (accumulate
// dense strides while stride <= hashmap_size, hash if the stride product exceeds the level's size, then wrap) is resolved
// per LEVEL here, so the per-corner work is two integer ops:
//   mode 0 (dense, never wraps):      row = x + y*cy + z*cz
//   mode 1 (hashed, power-of-two T):  row = (x ^ y*cy ^ z*cz) & mask         (cy, cz = the hash primes)
//   mode 2 (anything else: tiled grids that wrap, non power-of-two hash sizes): the generic entry_index()
struct LevelConst { uint32_t res, hashmap_size, offset, cy, cz, mask, mode, pad; };

__device__ __forceinline__ LevelConst make_level_const(uint32_t l, const GridArgs& g) {
    const uint32_t off = (uint32_t)__ldg(g.offsets + l);
    const uint32_t hs = (uint32_t)__ldg(g.offsets + l + 1) - off;
    const uint32_t res = level_resolution(l, g.S, g.H);
    LevelConst lv;
    lv.res = res; lv.offset = off; lv.hashmap_size = hs; lv.pad = 0;
    uint32_t stride = 1, c[3] = {0, 0, 0};
    bool all = true;
    for (uint32_t d = 0; d < 3; d++) {
        if (stride <= hs) { c[d] = stride; stride *= res; } else all = false;
    }
    const bool hashed = g.gridtype == 0 && stride > hs;
    if (hashed) {
        lv.cy = 2654435761u; lv.cz = 805459861u; lv.mask = hs - 1;
        lv.mode = ((hs & (hs - 1)) == 0) ? 1u : 2u;
    } else {
        lv.cy = c[1]; lv.cz = c[2]; lv.mask = 0xFFFFFFFFu;
        // all three strides accumulated and res^3 <= hs: the largest row is res^3 - 1 < hs, no wrap
        lv.mode = (all && stride <= hs) ? 0u : 2u;
    }
    return lv;
}

__device__ __forceinline__ void load_level_consts(LevelConst* s_lv, const GridArgs& g) {
    for (uint32_t l = threadIdx.x; l < g.L; l += blockDim.x) s_lv[l] = make_level_const(l, g);
}

// rows of the 8 corners of cell `base` (corner k: bit 0 = +x, bit 1 = +y, bit 2 = +z; +1 clamped to res-1); modes 0 and 1
__device__ __forceinline__ void corner_rows(const LevelConst& lv, const uint32_t (&base)[3], uint32_t (&rows)[8]) {
    const uint32_t x1 = min(base[0] + 1, lv.res - 1), y1 = min(base[1] + 1, lv.res - 1), z1 = min(base[2] + 1, lv.res - 1);
    const uint32_t ya = base[1] * lv.cy, yb = y1 * lv.cy, za = base[2] * lv.cz, zb = z1 * lv.cz;
    if (lv.mode == 0) {
        rows[0] = base[0] + ya + za; rows[1] = x1 + ya + za; rows[2] = base[0] + yb + za; rows[3] = x1 + yb + za;
        rows[4] = base[0] + ya + zb; rows[5] = x1 + ya + zb; rows[6] = base[0] + yb + zb; rows[7] = x1 + yb + zb;
    } else {
        rows[0] = (base[0] ^ ya ^ za) & lv.mask; rows[1] = (x1 ^ ya ^ za) & lv.mask;
        rows[2] = (base[0] ^ yb ^ za) & lv.mask; rows[3] = (x1 ^ yb ^ za) & lv.mask;
        rows[4] = (base[0] ^ ya ^ zb) & lv.mask; rows[5] = (x1 ^ ya ^ zb) & lv.mask;
        rows[6] = (base[0] ^ yb ^ zb) & lv.mask; rows[7] = (x1 ^ yb ^ zb) & lv.mask;
    }
}
// the same through the generic index map (mode 2 levels)
static __device__ __noinline__ void corner_rows_generic(uint32_t gridtype, uint32_t hashmap_size, uint32_t res, uint32_t bx, uint32_t by,
                                                 uint32_t bz, uint32_t* rows) {
    const uint32_t x1 = min(bx + 1, res - 1), y1 = min(by + 1, res - 1), z1 = min(bz + 1, res - 1);
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t q[3] = {(k & 1u) ? x1 : bx, (k & 2u) ? y1 : by, (k & 4u) ? z1 : bz};
        rows[k] = entry_index<3>(gridtype, hashmap_size, res, q);
    }
}

// trilinear weights in the reference's evaluation order ((1 * wx) * wy) * wz  (gridencoder.cu:172-186)
__device__ __forceinline__ void corner_weights(const float (&frac)[3], float (&w)[8]) {
    const float x0 = 1 - frac[0], y0 = 1 - frac[1], z0 = 1 - frac[2];
    const float xy[4] = {x0 * y0, frac[0] * y0, x0 * frac[1], frac[0] * frac[1]};
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) { w[k] = xy[k] * z0; w[k + 4] = xy[k] * frac[2]; }
}

// tcgen05.mma operand descriptors of one layer, built once per CTA (they do not change from tile to tile)
struct MmaStep { uint64_t a, b; };
constexpr uint32_t kMaxKSteps = 8;    // K <= 128
struct MmaPlan { MmaStep step[kMaxKSteps]; uint32_t idesc, n_steps, d_col, pad; };

__device__ __forceinline__ void issue_plan(uint32_t tmem, const MmaPlan& pl, bool accumulate_first) {
    for (uint32_t ks = 0; ks < pl.n_steps; ks++)
        tc::mma_f16_ss(tmem + pl.d_col, pl.step[ks].a, pl.step[ks].b, pl.idesc, (accumulate_first || ks > 0) ? 1u : 0u);
}

// position in [0,1]^3 the way GridEncoder.forward computes it: (x + bound) / (2*bound), the division by a host
// scalar being a multiplication by its fp32 reciprocal in torch (grid.py:160).
__device__ __forceinline__ void unit_cube(const float* __restrict__ xyz, float bound, float (&x)[3]) {
    const float inv = __fdiv_rn(1.0f, 2.0f * bound);
#pragma unroll
    for (int d = 0; d < 3; d++) x[d] = __fmul_rn(__fadd_rn(__ldg(xyz + d), bound), inv);
}

__device__ __forceinline__ bool locate3(const float (&x)[3], uint32_t res, bool align_corners, uint32_t interp,
                                        uint32_t (&base)[3], float (&frac)[3]) {
    if (x[0] < 0 || x[0] > 1 || x[1] < 0 || x[1] > 1 || x[2] < 0 || x[2] > 1) return false;
#pragma unroll
    for (uint32_t d = 0; d < 3; d++) {
        float p;
        if (align_corners) {
            p = x[d] * (float)(res - 1);
            base[d] = min((uint32_t)floorf(p), res - 2);
        } else {
            p = fminf(fmaxf(x[d] * (float)res - 0.5f, 0.0f), (float)(res - 1));
            base[d] = (uint32_t)floorf(p);
        }
        p -= (float)base[d];
        frac[d] = (interp == 1) ? smoothstep_f(p) : p;
    }
    return true;
}

__device__ __forceinline__ float half_round(float v) { return __half2float(__float2half_rn(v)); }


// One level of one sample: 8 gathers + half-precision interpolation (bit-identical to grid_forward_kernel with
// NGP_GRID_REF_ROUNDING).  Branch-free: a point outside [0,1]^3 (or a dead row) is clamped into the box so that its
// loads stay in bounds and the result is zeroed at the end; all gathers of a batch of levels are issued before the first
// one is consumed, and only the three interpolation fractions are kept live across the loads.
struct LevelGather {
    uint32_t v[8];      // raw half2 rows
    float frac[3];
};
__device__ __forceinline__ void gather_issue(LevelGather& q, const GridArgs& g, const LevelConst& lv, const float (&xc)[3]) {
    uint32_t base[3];
    locate3(xc, lv.res, g.align_corners, g.interp, base, q.frac);
    uint32_t rows[8];
    corner_rows(lv, base, rows);
    const uint32_t* __restrict__ lvl = reinterpret_cast<const uint32_t*>(g.table) + lv.offset;
#pragma unroll
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) q.v[k] = __ldg(lvl + rows[k]);
}
__device__ __forceinline__ __half2 gather_finish(const LevelGather& q, const GridArgs& g, uint32_t level, bool inside) {
    // Half-precision accumulation of the reference (at::Half results += w * value, gridencoder.cu:168,191): each product
    // is rounded to fp16 and added in fp16.  HADD2 rounds the exact sum once, which equals the reference's fp32 add +
    // fp16 rounding (24 >= 2 * 11 + 2 significand bits: no double rounding).
    float w[8];
    corner_weights(q.frac, w);
    __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q.v[k]));
        acc = __hadd2(acc, __floats2half2_rn(w[k] * f.x, w[k] * f.y));
    }
    if (g.feat_weights) {
        const float2 f = __half22float2(acc);
        acc = __floats2half2_rn(f.x * __ldg(g.feat_weights + 2 * level), f.y * __ldg(g.feat_weights + 2 * level + 1));
    }
    return inside ? acc : __floats2half2_rn(0.f, 0.f);
}

// a whole level through the generic index map (tiled grids that wrap, non power-of-two hash sizes): out of line, so
// that the common path above carries no call and keeps its gathers in registers
static __device__ __noinline__ uint32_t gather_level_generic(const __half* table, const float* feat_weights, uint32_t gridtype,
                                                      bool align_corners, uint32_t interp, uint32_t res, uint32_t hashmap_size,
                                                      uint32_t offset, float x0, float x1, float x2, uint32_t level, bool inside) {
    const float xc[3] = {x0, x1, x2};
    uint32_t base[3], rows[8];
    LevelGather q;
    locate3(xc, res, align_corners, interp, base, q.frac);
    corner_rows_generic(gridtype, hashmap_size, res, base[0], base[1], base[2], rows);
    const uint32_t* lvl = reinterpret_cast<const uint32_t*>(table) + offset;
    for (uint32_t k = 0; k < 8; k++) q.v[k] = __ldg(lvl + rows[k]);
    GridArgs g = {};
    g.feat_weights = feat_weights;
    const __half2 r = gather_finish(q, g, level, inside);
    return *reinterpret_cast<const uint32_t*>(&r);
}

// Two x-neighbour corners.  Rows that differ only in bit 0 share an aligned 8-byte slot of the fp16 F=2 table (dense
// levels: consecutive rows with an even first row; hashed levels: the first hash prime is 1, so x even -> rows h and h^1):
// one red.global.add.noftz.v2.f16x2 covers both.  Otherwise one packed reduction per corner.
__device__ __forceinline__ void scatter_pair_h2(__half* glvl, uint32_t row0, uint32_t row1, uint32_t p0, uint32_t p1) {
    if ((row0 ^ row1) == 1u) {
        const bool swap = row0 & 1u;
        red_add_v2_h2(glvl + (size_t)(row0 & ~1u) * 2, swap ? p1 : p0, swap ? p0 : p1);
    } else if (row0 == row1) {   // both corners clamp to the same row (level border)
        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&p0), *reinterpret_cast<const __half2*>(&p1));
        red_add_h2(glvl + (size_t)row0 * 2, *reinterpret_cast<const uint32_t*>(&sum));
    } else if (NGP_SCATTER_V4 && ((row0 ^ row1) >> 2) == 0u) {
        // same aligned group of four rows (x odd with x % 4 == 1: the +1 carries into bit 1 only): ONE 16-byte reduction with
        // +0.0 in the two other slots instead of two 4-byte ones -- the L2 applies a reduction per request, whatever its width
        const uint32_t a = row0 & 3u, b = row1 & 3u;
        red_add_v4_h2(glvl + (size_t)(row0 & ~3u) * 2, a == 0 ? p0 : (b == 0 ? p1 : 0u), a == 1 ? p0 : (b == 1 ? p1 : 0u),
                      a == 2 ? p0 : (b == 2 ? p1 : 0u), a == 3 ? p0 : (b == 3 ? p1 : 0u));
    } else {
        red_add_h2(glvl + (size_t)row0 * 2, p0);
        red_add_h2(glvl + (size_t)row1 * 2, p1);
    }
}

// A whole level of the scatter through the generic index map (mode 2 levels): out of line, reductions issued per corner.
static __device__ __noinline__ void scatter_corners_generic(__half* glvl, uint32_t gridtype, uint32_t hashmap_size, uint32_t res,
                                                            uint32_t bx, uint32_t by, uint32_t bz, uint32_t w0, uint32_t w1, uint32_t w2,
                                                            uint32_t w3, uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7) {
    const uint32_t wg[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
    const uint32_t x1 = min(bx + 1, res - 1), y1 = min(by + 1, res - 1), z1 = min(bz + 1, res - 1);
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t q[3] = {(k & 1u) ? x1 : bx, (k & 2u) ? y1 : by, (k & 4u) ? z1 : bz};
        red_add_h2(glvl + (size_t)entry_index<3>(gridtype, hashmap_size, res, q) * 2, wg[k]);
    }
}

// d (level features) / d x of one sample, for the gradient with respect to the POSITION (BARF: rays_o / rays_d receive
// gradients).  Like the reference's forward with calc_grad_inputs (gridencoder.cu:216-245) the derivative is produced where
// the 8 corner rows are already in registers -- the forward gather -- and saved in fp16 as dy_dx[level][dim][channel] (12 bytes
// per level); kernel_input_backward's contraction with the incoming gradient (gridencoder.cu:352-378) happens in the scatter
// warps of the backward kernel.
// out[d] = packed (channel 0, channel 1) of dimension d.
__device__ __forceinline__ void locate3_dfrac(const float (&x)[3], uint32_t res, bool align_corners, uint32_t interp, float (&dfrac)[3]) {
#pragma unroll
    for (uint32_t d = 0; d < 3; d++) {
        float p;
        if (align_corners) {
            p = x[d] * (float)(res - 1);
            p -= (float)min((uint32_t)floorf(p), res - 2);
        } else {
            p = fminf(fmaxf(x[d] * (float)res - 0.5f, 0.0f), (float)(res - 1));
            p -= floorf(p);
        }
        dfrac[d] = (interp == 1) ? smoothstep_df(p) : 1.0f;
    }
}
__device__ __forceinline__ void gather_finish_dydx(const LevelGather& q, const float (&dfrac)[3], float scale, bool inside,
                                                   uint32_t (&out)[3]) {
    // The derivative along one axis is the bilinear interpolation, over the two other axes, of the four corner differences
    // along that axis.  Evaluated in packed fp16 (both channels per instruction: HADD2 / HFMA2) -- the gather warps are issue
    // bound, and the reference accumulates these sums in fp16 as well (scalar_t results_grad, gridencoder.cu:222-243); the
    // scale factor res * d smoothstep is applied in fp32 at the end.
    const __half2* v = reinterpret_cast<const __half2*>(q.v);
    const __half2 f[3] = {__float2half2_rn(q.frac[0]), __float2half2_rn(q.frac[1]), __float2half2_rn(q.frac[2])};
    auto lerp = [](__half2 a, __half2 b, __half2 t) { return __hfma2(t, __hsub2(b, a), a); };
    __half2 r[3];
    // corner k: bit 0 = +x, bit 1 = +y, bit 2 = +z
    r[0] = lerp(lerp(__hsub2(v[1], v[0]), __hsub2(v[3], v[2]), f[1]), lerp(__hsub2(v[5], v[4]), __hsub2(v[7], v[6]), f[1]), f[2]);
    r[1] = lerp(lerp(__hsub2(v[2], v[0]), __hsub2(v[3], v[1]), f[0]), lerp(__hsub2(v[6], v[4]), __hsub2(v[7], v[5]), f[0]), f[2]);
    r[2] = lerp(lerp(__hsub2(v[4], v[0]), __hsub2(v[5], v[1]), f[0]), lerp(__hsub2(v[6], v[2]), __hsub2(v[7], v[3]), f[0]), f[1]);
#pragma unroll
    for (uint32_t d = 0; d < 3; d++) {
        const float2 c = __half22float2(r[d]);
        const float k = scale * dfrac[d];
        out[d] = inside ? pack_h2(k * c.x, k * c.y) : 0u;
    }
}

// the same for a level on the generic index map (mode 2): out of line like gather_level_generic
static __device__ __noinline__ void dydx_level_generic(const __half* table, uint32_t gridtype, bool align_corners, uint32_t interp, uint32_t res,
                                                       uint32_t hashmap_size, uint32_t offset, float x0, float x1, float x2, bool inside,
                                                       uint32_t* out) {
    const float xc[3] = {x0, x1, x2};
    uint32_t base[3], rows[8];
    LevelGather q;
    float dfrac[3];
    locate3(xc, res, align_corners, interp, base, q.frac);
    locate3_dfrac(xc, res, align_corners, interp, dfrac);
    corner_rows_generic(gridtype, hashmap_size, res, base[0], base[1], base[2], rows);
    const uint32_t* lvl = reinterpret_cast<const uint32_t*>(table) + offset;
    for (uint32_t k = 0; k < 8; k++) q.v[k] = __ldg(lvl + rows[k]);
    uint32_t o[3];
    gather_finish_dydx(q, dfrac, (float)(align_corners ? res - 1 : res), inside, o);
    out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}

// One level of one sample of the table-gradient scatter.  `gh` is d enc (2 features) as it comes out of the grid_mlp
// backward; it reaches the encoder as fp16 (autocast), optionally through the annealing window.  Must be called by all 32
// lanes of a warp whose lanes hold consecutive samples (the warp aggregation below uses full-warp shuffles).
__device__ __forceinline__ void scatter_level(const GridArgs& g, const LevelConst& lv, uint32_t level, const float (&x)[3], bool live,
                                              __half2 gh, __half* __restrict__ grad_table, uint32_t lane) {
    if (g.feat_weights) {
        const float2 gf = __half22float2(gh);
        gh = __floats2half2_rn(gf.x * __ldg(g.feat_weights + 2 * level), gf.y * __ldg(g.feat_weights + 2 * level + 1));
    }
    const float2 gf = __half22float2(gh);
    uint32_t base[3];
    float frac[3];
    const bool valid = live && locate3(x, lv.res, g.align_corners, g.interp, base, frac);
    // weighted contributions of the 8 corners, packed (channel 0, channel 1) in fp16: the table gradient is fp16 and is
    // accumulated by fp16 reductions either way (the reference issues one fp16x2 atomic per corner and sample,
    // gridencoder.cu:334-340)
    uint32_t wg[8];
    {
        float w[8];
        corner_weights(frac, w);
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) wg[k] = valid ? pack_h2(w[k] * gf.x, w[k] * gf.y) : 0u;
    }
    // warp aggregation: consecutive samples of a ray that sit in the same cell are summed with a segmented shuffle
    // reduction (as many rounds as the longest run needs) and only run heads issue reductions
    uint32_t key0 = 0xFFFFFFFFu, key1 = 0xFFFFFF00u | lane;
    if (valid) { key0 = base[0] | (base[1] << 16); key1 = base[2]; }
    const uint32_t pk0 = __shfl_up_sync(0xffffffffu, key0, 1), pk1 = __shfl_up_sync(0xffffffffu, key1, 1);
    const bool head = (lane == 0) || (pk0 != key0) || (pk1 != key1);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (heads != 0xffffffffu) {
        const uint32_t above = heads & ~((2u << lane) - 1u);
        const uint32_t end = (lane == 31 || above == 0) ? 31u : (uint32_t)__ffs(above) - 2u;
        const uint32_t longest = __reduce_max_sync(0xffffffffu, head ? end - lane + 1u : 0u);
        for (uint32_t d = 1; d < longest; d <<= 1) {
            const bool take = lane + d <= end;
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                uint32_t o = __shfl_down_sync(0xffffffffu, wg[k], d);
                o = take ? o : 0u;          // (+0.0, +0.0): adding it leaves the value unchanged
                const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&wg[k]), *reinterpret_cast<const __half2*>(&o));
                wg[k] = *reinterpret_cast<const uint32_t*>(&sum);
            }
        }
    }
    if (valid && head) {
        __half* glvl = grad_table + (size_t)lv.offset * 2;
        if (lv.mode == 2) {     // warp-uniform, rare
            scatter_corners_generic(glvl, g.gridtype, lv.hashmap_size, lv.res, base[0], base[1], base[2], wg[0], wg[1], wg[2], wg[3],
                                    wg[4], wg[5], wg[6], wg[7]);
        } else {
            uint32_t rows[8];
            corner_rows(lv, base, rows);
#pragma unroll
            for (uint32_t k = 0; k < 8; k += 2) scatter_pair_h2(glvl, rows[k], rows[k + 1], wg[k], wg[k + 1]);
        }
    }
}

}  // namespace fieldcore
}  // namespace ngp
