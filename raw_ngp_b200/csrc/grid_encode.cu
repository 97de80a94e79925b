// grid_encode.cu -- multiresolution hash / tiled grid encoder for sm_100a.
//
// Implements the semantics of the reference kernels kernel_grid, kernel_grid_backward,
// kernel_input_backward, kernel_grad_tv and kernel_grad_wd (gridencoder/src/gridencoder.cu:82-249,
// 252-349, 352-378, 525-631, 670-703) with a different data path:
//   * one thread per (point, level), blocks of one level are scheduled together (blockIdx.y = level)
//     so that a level's slice of the table (<= 2 MiB fp16) is what the SM's L1 and the 126 MB L2 see;
//   * every corner is ONE vector load of the whole C-wide row through ld.global.nc (half2 for the
//     fp16 F=2 table) and every gradient corner is ONE packed reduction (red.global.add.noftz.f16x2 /
//     red.global.add.v2.f32), never a per-element atomic;
//   * outputs and incoming gradients use the operator's own [B, L*C] layout, so the wrapper does no
//     permute copies (gridencoder/grid.py:63,81);
//   * the input gradient is recomputed from the table inside the backward kernel instead of being
//     streamed out as dy_dx [B, L*D*C] in forward and back in (gridencoder/grid.py:54, gridencoder.cu:352-378).
#include "common.cuh"
#include "grid_core.cuh"
#include "field_core.cuh"

namespace ngp {
namespace {

using namespace gridcore;

constexpr uint32_t kFwdThreads = 256;
constexpr uint32_t kBwdThreads = 256;

// Accumulators.  RefRound on fp16 tables reproduces the reference's at::Half arithmetic
// (product rounded to half, then a half+half add evaluated in fp32 and rounded; gridencoder.cu:168,191).
template <typename T, bool RefRound> struct Acc {
    float v;
    __device__ __forceinline__ Acc() : v(0.f) {}
    __device__ __forceinline__ void fma(float w, float g) { v += w * g; }
    __device__ __forceinline__ T get() const { return from_f32<T>(v); }
};
template <> struct Acc<__half, true> {
    __half v;
    __device__ __forceinline__ Acc() : v(__float2half_rn(0.f)) {}
    __device__ __forceinline__ void fma(float w, float g) {
        __half t = __float2half_rn(w * g);
        v = __float2half_rn(__half2float(v) + __half2float(t));
    }
    __device__ __forceinline__ __half get() const { return v; }
};

template <typename T, uint32_t D, uint32_t C, bool RefRound>
__global__ void __launch_bounds__(kFwdThreads)
grid_forward_kernel(const float* __restrict__ inputs, const T* __restrict__ table, const int* __restrict__ offsets,
                    T* __restrict__ outputs, T* __restrict__ dy_dx, uint32_t B, uint32_t L, uint32_t max_level,
                    float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;

    // The model's configuration (fp16 table, D = 3, F = 2, reference rounding, no dy_dx): per-level index constants, a
    // branch-free gather and native half2 accumulation (csrc/field_core.cuh) -- the code path of the fused field kernels.
    if constexpr (std::is_same<T, __half>::value && D == 3 && C == 2 && RefRound) {
        if (!dy_dx && level < max_level) {
            const fieldcore::GridArgs g = {table, offsets, nullptr, S, 1.f, H, L, gridtype, interp, align_corners};
            const fieldcore::LevelConst lv = fieldcore::make_level_const(level, g);
            const float x[3] = {__ldg(inputs + (size_t)b * 3), __ldg(inputs + (size_t)b * 3 + 1), __ldg(inputs + (size_t)b * 3 + 2)};
            const bool inside = x[0] >= 0 && x[0] <= 1 && x[1] >= 0 && x[1] <= 1 && x[2] >= 0 && x[2] <= 1;
            const float xc[3] = {fminf(fmaxf(x[0], 0.f), 1.f), fminf(fmaxf(x[1], 0.f), 1.f), fminf(fmaxf(x[2], 0.f), 1.f)};
            __half2 f;
            if (lv.mode == 2) {
                const uint32_t r = fieldcore::gather_level_generic(table, nullptr, gridtype, align_corners, interp, lv.res, lv.hashmap_size,
                                                                   lv.offset, xc[0], xc[1], xc[2], level, inside);
                f = *reinterpret_cast<const __half2*>(&r);
            } else {
                fieldcore::LevelGather q;
                fieldcore::gather_issue(q, g, lv, xc);
                f = fieldcore::gather_finish(q, g, level, inside);
            }
            *reinterpret_cast<__half2*>(outputs + (size_t)b * (L * C) + level * C) = f;
            return;
        }
    }

    T* out = outputs + (size_t)b * (L * C) + level * C;
    T* dout = dy_dx ? dy_dx + ((size_t)b * L + level) * (D * C) : nullptr;

    T zero[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) zero[c] = from_f32<T>(0.f);

    uint32_t base[D];
    float frac[D], dfrac[D];
    const uint32_t res = level_resolution(level, S, H);
    if (level >= max_level || !locate<D>(inputs + (size_t)b * D, res, align_corners, interp, base, frac, dfrac)) {
        store_row<T, C>(out, zero);
        if (dout) {
#pragma unroll
            for (uint32_t d = 0; d < D; d++) store_row<T, C>(dout + d * C, zero);
        }
        return;
    }

    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const T* __restrict__ lvl = table + (size_t)off * C;

    // gather the 2^D corner rows (independent loads, all in flight together)
    float val[1u << D][C];
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); k++) {
        uint32_t p[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) p[d] = (k & (1u << d)) ? min(base[d] + 1, res - 1) : base[d];
        load_row<T, C>(lvl + (size_t)entry_index<D>(gridtype, hashmap_size, res, p) * C, val[k]);
    }

    Acc<T, RefRound> acc[C];
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); k++) {
        float w = 1;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) w *= (k & (1u << d)) ? frac[d] : 1 - frac[d];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) acc[c].fma(w, val[k][c]);
    }
    T res_out[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) res_out[c] = acc[c].get();
    store_row<T, C>(out, res_out);

    if (dout) {
        // d out / d x_g = scale * dfrac_g * sum over the 2^(D-1) corner pairs along g (gridencoder.cu:205-247)
        const float scale = (float)(align_corners ? res - 1 : res);
#pragma unroll
        for (uint32_t g = 0; g < D; g++) {
            Acc<T, RefRound> dacc[C];
#pragma unroll
            for (uint32_t j = 0; j < (1u << (D - 1)); j++) {
                float w = scale;
                uint32_t lo = 0;  // corner id with bit g clear
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= g) ? nd + 1 : nd;
                    if (j & (1u << nd)) { w *= frac[d]; lo |= (1u << d); }
                    else w *= 1 - frac[d];
                }
                const uint32_t hi = lo | (1u << g);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) {
                    float diff = val[hi][c] - val[lo][c];
                    if (RefRound && std::is_same<T, __half>::value) diff = __half2float(__float2half_rn(diff));
                    dacc[c].fma(w * diff, dfrac[g]);
                }
            }
            T dres[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) dres[c] = dacc[c].get();
            store_row<T, C>(dout + g * C, dres);
        }
    }
}


// ---- forward, tile kernel (the model's shape: D = 3, C = 2, L % 8 == 0, no dy_dx) ---------------------------------
// The reference launches one thread per (point, level) with blockIdx.y = level: every level re-reads the input, writes 4
// (8) bytes per thread at a 64 (128) byte stride, and a thread has 8 loads in flight (gridencoder.cu:82-249, :467-490).
// Here a persistent CTA takes tiles of 128 points; thread (row, g) encodes levels g, g + 4, g + 8, ... of its point -- coarse
// dense levels (L1 hits) and fine hashed levels (L2 round trips) mixed in every thread, two levels = 16 table rows in flight
// per thread, branch-free with per-level index constants (field_core.cuh) -- and leaves the features in a shared-memory
// image of the tile's [128, L * C] output block, which then goes out as whole contiguous rows (4-byte words of consecutive
// threads are consecutive in global memory: every store instruction of a warp writes one full 128-byte line).  The input is
// read once per point (4 threads share the 12 bytes through L1).
constexpr uint32_t kTilePts = 128, kTileGroups = 4, kTileThreads = kTilePts * kTileGroups;

template <typename T> struct Row2;                       // one table row (C = 2) as a register value
template <> struct Row2<__half> {
    using raw = uint32_t;
    static constexpr uint32_t kWords = 1;
    static __device__ __forceinline__ float2 f2(raw v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
};
template <> struct Row2<__nv_bfloat16> {
    using raw = uint32_t;
    static constexpr uint32_t kWords = 1;
    static __device__ __forceinline__ float2 f2(raw v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
};
template <> struct Row2<float> {
    using raw = float2;
    static constexpr uint32_t kWords = 2;
    static __device__ __forceinline__ float2 f2(raw v) { return v; }
};

// interpolates the 8 gathered rows of one level and writes the level's two features into the tile image
template <typename T, bool RefRound>
__device__ __forceinline__ void tile_finish(const typename Row2<T>::raw (&v)[8], const float (&frac)[3], bool inside, uint32_t* dst) {
    float w[8];
    fieldcore::corner_weights(frac, w);
    if constexpr (std::is_same<T, __half>::value && RefRound) {
        // at::Half accumulation of the reference: product rounded to fp16, fp16 + fp16 rounded once (field_core.cuh)
        __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) {
            const float2 f = Row2<T>::f2(v[k]);
            acc = __hadd2(acc, __floats2half2_rn(w[k] * f.x, w[k] * f.y));
        }
        dst[0] = inside ? *reinterpret_cast<const uint32_t*>(&acc) : 0u;
    } else {
        float a0 = 0.f, a1 = 0.f;                       // fp32 accumulation in corner order (FFMA chain like the reference's)
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) {
            const float2 f = Row2<T>::f2(v[k]);
            a0 = fmaf(w[k], f.x, a0);
            a1 = fmaf(w[k], f.y, a1);
        }
        if (!inside) { a0 = 0.f; a1 = 0.f; }
        if constexpr (std::is_same<T, float>::value) { dst[0] = __float_as_uint(a0); dst[1] = __float_as_uint(a1); }
        else if constexpr (std::is_same<T, __half>::value) dst[0] = pack_h2(a0, a1);
        else dst[0] = pack_bf2(a0, a1);
    }
}

template <typename T>
__device__ __forceinline__ void tile_issue(typename Row2<T>::raw (&v)[8], float (&frac)[3], const T* __restrict__ table,
                                           const fieldcore::LevelConst& lv, bool align_corners, uint32_t interp, const float (&xc)[3]) {
    using raw = typename Row2<T>::raw;
    uint32_t base[3], rows[8];
    fieldcore::locate3(xc, lv.res, align_corners, interp, base, frac);
    fieldcore::corner_rows(lv, base, rows);
    const raw* __restrict__ lvl = reinterpret_cast<const raw*>(table) + lv.offset;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) v[k] = __ldg(lvl + rows[k]);
}

// a whole level through the generic index map (mode 2: tiled grids that wrap, hash sizes that are not powers of two): out of
// line, so that the common path keeps its corner rows in registers
template <typename T, bool RefRound>
static __device__ __noinline__ void tile_level_generic(const T* table, uint32_t gridtype, bool align_corners, uint32_t interp, uint32_t res,
                                                       uint32_t hashmap_size, uint32_t offset, float x0, float x1, float x2, bool inside,
                                                       uint32_t* dst) {
    using raw = typename Row2<T>::raw;
    const float xc[3] = {x0, x1, x2};
    uint32_t base[3], rows[8];
    float frac[3];
    raw v[8];
    fieldcore::locate3(xc, res, align_corners, interp, base, frac);
    fieldcore::corner_rows_generic(gridtype, hashmap_size, res, base[0], base[1], base[2], rows);
    const raw* lvl = reinterpret_cast<const raw*>(table) + offset;
    for (uint32_t k = 0; k < 8; k++) v[k] = __ldg(lvl + rows[k]);
    tile_finish<T, RefRound>(v, frac, inside, dst);
}

template <typename T, bool RefRound>
__global__ void __launch_bounds__(kTileThreads, 2)
grid_forward_tile_kernel(const float* __restrict__ inputs, const T* __restrict__ table, const int* __restrict__ offsets,
                         T* __restrict__ outputs, uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
                         uint32_t gridtype, bool align_corners, uint32_t interp) {
    using raw = typename Row2<T>::raw;
    constexpr uint32_t W = Row2<T>::kWords;
    extern __shared__ uint32_t s_tile[];                 // [128][L * W + 1] words: odd row stride, conflict-free both ways
    __shared__ fieldcore::LevelConst s_lv[fieldcore::kMaxLevels];
    {
        const fieldcore::GridArgs g = {nullptr, offsets, nullptr, S, 1.f, H, L, gridtype, interp, align_corners};
        fieldcore::load_level_consts(s_lv, g);
    }
    __syncthreads();
    const uint32_t r = threadIdx.x & (kTilePts - 1), grp = threadIdx.x / kTilePts;
    const uint32_t stride = L * W + 1, row_words = L * W;
    const uint32_t n_tiles = div_up(B, kTilePts);
    uint32_t* out_words = reinterpret_cast<uint32_t*>(outputs);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kTilePts + r;
        float x[3] = {2.f, 2.f, 2.f};
        if (row < B) { x[0] = __ldg(inputs + (size_t)row * 3); x[1] = __ldg(inputs + (size_t)row * 3 + 1); x[2] = __ldg(inputs + (size_t)row * 3 + 2); }
        const bool inside = x[0] >= 0 && x[0] <= 1 && x[1] >= 0 && x[1] <= 1 && x[2] >= 0 && x[2] <= 1;
        const float xc[3] = {fminf(fmaxf(x[0], 0.f), 1.f), fminf(fmaxf(x[1], 0.f), 1.f), fminf(fmaxf(x[2], 0.f), 1.f)};
        uint32_t* my = s_tile + r * stride;
        for (uint32_t level = grp; level < L; level += 2 * kTileGroups) {
            const uint32_t la = level, lb = level + kTileGroups;          // L % 8 == 0
            // levels at or above max_level are zero-filled (grid.py:41,52); clamping the level keeps the code branch-free
            const fieldcore::LevelConst& lva = s_lv[min(la, max_level - 1)];
            const fieldcore::LevelConst& lvb = s_lv[min(lb, max_level - 1)];
            if (lva.mode == 2 || lvb.mode == 2) {                        // warp-uniform, rare
                tile_level_generic<T, RefRound>(table, gridtype, align_corners, interp, lva.res, lva.hashmap_size, lva.offset, xc[0], xc[1],
                                                xc[2], inside && la < max_level, my + la * W);
                tile_level_generic<T, RefRound>(table, gridtype, align_corners, interp, lvb.res, lvb.hashmap_size, lvb.offset, xc[0], xc[1],
                                                xc[2], inside && lb < max_level, my + lb * W);
                continue;
            }
            raw va[8], vb[8];
            float fa[3], fb[3];
            tile_issue<T>(va, fa, table, lva, align_corners, interp, xc);
            tile_issue<T>(vb, fb, table, lvb, align_corners, interp, xc);
            tile_finish<T, RefRound>(va, fa, inside && la < max_level, my + la * W);
            tile_finish<T, RefRound>(vb, fb, inside && lb < max_level, my + lb * W);
        }
        __syncthreads();
        // the tile's output block is contiguous in global memory: word i of the block = row i / row_words, word i % row_words
        const uint32_t live_words = min(kTilePts, B - tile * kTilePts) * row_words;
        uint32_t* dst = out_words + (size_t)tile * kTilePts * row_words;
        for (uint32_t i = threadIdx.x; i < live_words; i += kTileThreads) {
            const uint32_t rr = i / row_words;
            dst[i] = s_tile[i + rr];
        }
        __syncthreads();
    }
}

// Backward: scatter w * grad into the table gradient and, if asked, recompute d out / d x from the table and reduce
// over the level's channels into grad_inputs.
//
// Consecutive samples of a ray fall into the same cell on the coarse levels (about 37 samples per cell at
// resolution 16 with dt = 2*sqrt(3)/1024), so every warp first merges runs of lanes that share a cell with a
// segmented shuffle reduction and only the head lane of each run issues reductions (warp-aggregated atomics);
// levels where no two neighbouring lanes share a cell skip the merge (one ballot of overhead).
template <typename T, uint32_t D, uint32_t C, bool InputGrad>
__global__ void __launch_bounds__(kBwdThreads)
grid_backward_kernel(const T* __restrict__ grad, const float* __restrict__ inputs, const T* __restrict__ table,
                     const int* __restrict__ offsets, T* __restrict__ grad_table, float* __restrict__ grad_inputs,
                     uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners,
                     uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t level = blockIdx.y;

    uint32_t base[D];
    float frac[D], dfrac[D];
    const uint32_t res = level_resolution(level, S, H);
    const bool valid = (b < B) && locate<D>(inputs + (size_t)(b < B ? b : 0) * D, res, align_corners, interp, base, frac, dfrac);

    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;

    float g[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) g[c] = 0.f;
    if (valid) load_row<T, C>(grad + (size_t)b * (L * C) + level * C, g);

    // weighted contributions of this sample to its 2^D corners
    float wg[1u << D][C];
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); k++) {
        float w = 1;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) w *= (k & (1u << d)) ? frac[d] : 1 - frac[d];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) wg[k][c] = valid ? w * g[c] : 0.f;
    }

    // ---- warp aggregation over runs of lanes in the same cell --------------------------------------------------
    uint32_t key0 = 0xFFFFFFFFu, key1 = 0xFFFFFF00u | lane;   // invalid lanes never match a neighbour
    if (valid) {
        key0 = base[0] | (base[1] << 16);                      // resolutions are < 65536
        key1 = (D > 2) ? base[D - 1] : 0u;
    }
    const uint32_t pk0 = __shfl_up_sync(0xffffffffu, key0, 1), pk1 = __shfl_up_sync(0xffffffffu, key1, 1);
    const bool head = (lane == 0) || (pk0 != key0) || (pk1 != key1);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (heads != 0xffffffffu) {
        const uint32_t above = heads & ~((2u << lane) - 1u);          // heads strictly above this lane
        const uint32_t end = (lane == 31 || above == 0) ? 31u : (uint32_t)__ffs(above) - 2u;  // last lane of my run
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); k++) {
#pragma unroll
                for (uint32_t c = 0; c < C; c++) {
                    const float o = __shfl_down_sync(0xffffffffu, wg[k][c], d);
                    if (lane + d <= end) wg[k][c] += o;
                }
            }
        }
    }

    uint32_t row[1u << D];
    if (valid) {
#pragma unroll
        for (uint32_t k = 0; k < (1u << D); k++) {
            uint32_t p[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) p[d] = (k & (1u << d)) ? min(base[d] + 1, res - 1) : base[d];
            row[k] = entry_index<D>(gridtype, hashmap_size, res, p);
        }
    }

    if (valid && head) {
        T* glvl = grad_table + (size_t)off * C;
#pragma unroll
        for (uint32_t k = 0; k < (1u << D); k += 2) scatter_pair<T, C>(glvl, row[k], row[k + 1], wg[k], wg[k + 1]);
    }

    if (InputGrad && valid) {
        float val[1u << D][C];
        const T* __restrict__ lvl = table + (size_t)off * C;
#pragma unroll
        for (uint32_t k = 0; k < (1u << D); k++) load_row<T, C>(lvl + (size_t)row[k] * C, val[k]);
        const float scale = (float)(align_corners ? res - 1 : res);
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            float s = 0.f;
#pragma unroll
            for (uint32_t j = 0; j < (1u << (D - 1)); j++) {
                float w = scale;
                uint32_t lo = 0;
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                    if (j & (1u << nd)) { w *= frac[d]; lo |= (1u << d); }
                    else w *= 1 - frac[d];
                }
                const uint32_t hi = lo | (1u << gd);
                float dot = 0.f;
#pragma unroll
                for (uint32_t c = 0; c < C; c++) dot += g[c] * (val[hi][c] - val[lo][c]);
                s += w * dot;
            }
            red_add_f32(grad_inputs + (size_t)b * D + gd, s * dfrac[gd]);
        }
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[b,l,c] * dy_dx[b,l,d,c]   (reference layout; gridencoder.cu:352-378)
template <typename T, uint32_t D, uint32_t C>
__global__ void grid_input_backward_kernel(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                           float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* gr = grad + (size_t)b * L * C;
    const T* dd = dy_dx + (size_t)b * L * D * C + d * C;
    float s = 0.f;
    for (uint32_t l = 0; l < L; l++) {
        float gv[C], dv[C];
        load_row<T, C>(gr + l * C, gv);
        load_row<T, C>(dd + (size_t)l * D * C, dv);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) s += gv[c] * dv[c];
    }
    grad_inputs[t] = s;
}

// Total-variation gradient at B sample positions per level.  gridencoder.cu:525-631.
template <typename T, uint32_t D, uint32_t C>
__global__ void grid_tv_kernel(const T* __restrict__ inputs, const T* __restrict__ table, T* __restrict__ grad,
                               const int* __restrict__ offsets, float weight, uint32_t B, uint32_t L, float S,
                               uint32_t H, uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = to_f32(inputs[(size_t)b * D + d]);
        if (x[d] < 0 || x[d] > 1) return;
    }
    const uint32_t off = (uint32_t)__ldg(offsets + level);
    const uint32_t hashmap_size = (uint32_t)__ldg(offsets + level + 1) - off;
    const uint32_t res = level_resolution(level, S, H);
    const T* __restrict__ lvl = table + (size_t)off * C;

    uint32_t cell[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (align_corners) cell[d] = min((uint32_t)floorf(x[d] * (float)(res - 1)), res - 2);
        else cell[d] = (uint32_t)floorf(fminf(fmaxf(x[d] * (float)res - 0.5f, 0.0f), (float)(res - 1)));
    }
    const uint32_t centre = entry_index<D>(gridtype, hashmap_size, res, cell);
    float vc[C];
    load_row<T, C>(lvl + (size_t)centre * C, vc);

    float sum[C], sq[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) sum[c] = sq[c] = 0.f;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = cell[d];
        float vn[C];
        if (cur < res) {  // always true, kept: the "+1" neighbour may be res (wraps through the index map)
            cell[d] = cur + 1;
            load_row<T, C>(lvl + (size_t)entry_index<D>(gridtype, hashmap_size, res, cell) * C, vn);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { float df = vc[c] - vn[c]; sum[c] += df; sq[c] += df * df; }
        }
        if (cur > 0) {
            cell[d] = cur - 1;
            load_row<T, C>(lvl + (size_t)entry_index<D>(gridtype, hashmap_size, res, cell) * C, vn);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { float df = vc[c] - vn[c]; sum[c] += df; sq[c] += df * df; }
        }
        cell[d] = cur;
    }
    const float w = weight / (2 * D);
    float upd[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) upd[c] = w * sum[c] * rsqrtf(sq[c] + 1e-9f);
    red_add_row<T, C>(grad + ((size_t)off + centre) * C, upd);
}

// Level-wise mean weight decay: grad += 2*w*table / hashmap_size(level).  gridencoder.cu:670-703.
template <typename T>
__global__ void grid_wd_kernel(const T* __restrict__ table, T* __restrict__ grad, const int* __restrict__ offsets,
                               float weight, uint32_t n_elems, uint32_t L, uint32_t C) {
    extern __shared__ int s_off[];
    for (uint32_t i = threadIdx.x; i <= L; i += blockDim.x) s_off[i] = offsets[i];
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
        const uint32_t n = e / C;
        uint32_t lo = 0, hi = L, level = 0;
        while (lo < hi) {
            const uint32_t m = (lo + hi) / 2;
            if ((uint32_t)s_off[m] <= n) { level = m; lo = m + 1; } else hi = m;
        }
        const uint32_t hashmap_size = (uint32_t)(s_off[level + 1] - s_off[level]);
        grad[e] = from_f32<T>(to_f32(grad[e]) + 2 * weight * to_f32(table[e]) / hashmap_size);
    }
}

__global__ void level_resolution_kernel(uint32_t L, float S, uint32_t H, uint32_t* __restrict__ out) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) out[l] = level_resolution(l, S, H);
}

template <typename T, uint32_t D>
int launch_forward(const float* inputs, const T* table, const int* offsets, T* outputs, T* dy_dx, uint32_t B,
                   uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H, uint32_t gridtype,
                   bool align_corners, uint32_t interp, bool ref_round, bool point_level, cudaStream_t st) {
    if constexpr (D == 3) {
        if (!point_level && C == 2 && !dy_dx && L % 8 == 0 && L <= fieldcore::kMaxLevels && max_level > 0) {
            const uint32_t smem = kTilePts * (L * Row2<T>::kWords + 1) * 4;
            const uint32_t blocks = std::min<uint32_t>(div_up(B, kTilePts), 2 * kNumSMs);
            if (ref_round) grid_forward_tile_kernel<T, true><<<blocks, kTileThreads, smem, st>>>(inputs, table, offsets, outputs, B, L, max_level, S, H, gridtype, align_corners, interp);
            else grid_forward_tile_kernel<T, false><<<blocks, kTileThreads, smem, st>>>(inputs, table, offsets, outputs, B, L, max_level, S, H, gridtype, align_corners, interp);
            return finish_launch();
        }
    }
    const dim3 grid(div_up(B, kFwdThreads), L, 1);
#define NGP_FWD(CC)                                                                                              \
    if (ref_round)                                                                                               \
        grid_forward_kernel<T, D, CC, true><<<grid, kFwdThreads, 0, st>>>(inputs, table, offsets, outputs, dy_dx, B, L, \
                                                                          max_level, S, H, gridtype, align_corners, interp); \
    else                                                                                                         \
        grid_forward_kernel<T, D, CC, false><<<grid, kFwdThreads, 0, st>>>(inputs, table, offsets, outputs, dy_dx, B, L, \
                                                                           max_level, S, H, gridtype, align_corners, interp);
    switch (C) {
        case 1: NGP_FWD(1); break;
        case 2: NGP_FWD(2); break;
        case 4: NGP_FWD(4); break;
        case 8: NGP_FWD(8); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
#undef NGP_FWD
    return finish_launch();
}

template <typename T, uint32_t D>
int launch_backward(const T* grad, const float* inputs, const T* table, const int* offsets, T* grad_table,
                    float* grad_inputs, uint32_t B, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                    uint32_t gridtype, bool align_corners, uint32_t interp, cudaStream_t st) {
    const dim3 grid(div_up(B, kBwdThreads), max_level, 1);
#define NGP_BWD(CC)                                                                                              \
    if (grad_inputs)                                                                                             \
        grid_backward_kernel<T, D, CC, true><<<grid, kBwdThreads, 0, st>>>(grad, inputs, table, offsets, grad_table, \
                                                                           grad_inputs, B, L, S, H, gridtype, align_corners, interp); \
    else                                                                                                         \
        grid_backward_kernel<T, D, CC, false><<<grid, kBwdThreads, 0, st>>>(grad, inputs, table, offsets, grad_table, \
                                                                            grad_inputs, B, L, S, H, gridtype, align_corners, interp);
    switch (C) {
        case 1: NGP_BWD(1); break;
        case 2: NGP_BWD(2); break;
        case 4: NGP_BWD(4); break;
        case 8: NGP_BWD(8); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
#undef NGP_BWD
    return finish_launch();
}

template <typename T, uint32_t D>
int launch_input_backward(const T* grad, const T* dy_dx, float* grad_inputs, uint32_t B, uint32_t C, uint32_t L,
                          cudaStream_t st) {
    const uint32_t blocks = div_up(B * D, 256u);
    switch (C) {
        case 1: grid_input_backward_kernel<T, D, 1><<<blocks, 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L); break;
        case 2: grid_input_backward_kernel<T, D, 2><<<blocks, 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L); break;
        case 4: grid_input_backward_kernel<T, D, 4><<<blocks, 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L); break;
        case 8: grid_input_backward_kernel<T, D, 8><<<blocks, 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
    return finish_launch();
}

template <typename T, uint32_t D>
int launch_tv(const T* inputs, const T* table, T* grad, const int* offsets, float weight, uint32_t B, uint32_t C,
              uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners, cudaStream_t st) {
    const dim3 grid(div_up(B, 256u), L, 1);
    switch (C) {
        case 1: grid_tv_kernel<T, D, 1><<<grid, 256, 0, st>>>(inputs, table, grad, offsets, weight, B, L, S, H, gridtype, align_corners); break;
        case 2: grid_tv_kernel<T, D, 2><<<grid, 256, 0, st>>>(inputs, table, grad, offsets, weight, B, L, S, H, gridtype, align_corners); break;
        case 4: grid_tv_kernel<T, D, 4><<<grid, 256, 0, st>>>(inputs, table, grad, offsets, weight, B, L, S, H, gridtype, align_corners); break;
        case 8: grid_tv_kernel<T, D, 8><<<grid, 256, 0, st>>>(inputs, table, grad, offsets, weight, B, L, S, H, gridtype, align_corners); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
    return finish_launch();
}

// table rows must be aligned to their own width for the vector loads / packed reductions
static inline bool row_aligned(const void* p, uint32_t C, size_t elem) {
    size_t w = C * elem;
    if (w > 16) w = 16;
    return aligned(p, w);
}

}  // namespace
}  // namespace ngp

using namespace ngp;

#define NGP_DISPATCH_DTYPE(dtype, ...)                                  \
    switch (dtype) {                                                    \
        case NGP_F32: { using T = float; __VA_ARGS__; } break;          \
        case NGP_F16: { using T = __half; __VA_ARGS__; } break;         \
        case NGP_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
        default: return NGP_ERR_BAD_DTYPE;                              \
    }

static inline size_t dtype_size(int dtype) { return dtype == NGP_F32 ? 4 : 2; }

extern "C" int ngp_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                                       void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                       uint32_t max_level, float S, uint32_t H, void* dy_dx, uint32_t gridtype,
                                       int align_corners, uint32_t interp, int dtype, uint32_t flags,
                                       ngp_stream_t stream) {
    if (dtype < NGP_F32 || dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (B == 0 || L == 0) return NGP_OK;
    if (!inputs || !embeddings || !offsets || !outputs) return NGP_ERR_NULL;
    if (gridtype > 1 || interp > 1 || max_level > L) return NGP_ERR_BAD_ARG;
    if (C != 1 && C != 2 && C != 4 && C != 8) return NGP_ERR_UNSUPPORTED;
    const size_t es = dtype_size(dtype);
    if (!row_aligned(embeddings, C, es) || !row_aligned(outputs, C, es) || (dy_dx && !row_aligned(dy_dx, C, es)) ||
        !aligned(inputs, 4))
        return NGP_ERR_ALIGN;
    const bool ref_round = (flags & NGP_GRID_REF_ROUNDING) && dtype == NGP_F16;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = NGP_ERR_UNSUPPORTED;
    NGP_DISPATCH_DTYPE(dtype, {
        if (D == 3) rc = launch_forward<T, 3>(inputs, (const T*)embeddings, offsets, (T*)outputs, (T*)dy_dx, B, C, L, max_level, S, H, gridtype, align_corners != 0, interp, ref_round, (flags & NGP_GRID_POINT_LEVEL_KERNELS) != 0, st);
        else if (D == 2) rc = launch_forward<T, 2>(inputs, (const T*)embeddings, offsets, (T*)outputs, (T*)dy_dx, B, C, L, max_level, S, H, gridtype, align_corners != 0, interp, ref_round, (flags & NGP_GRID_POINT_LEVEL_KERNELS) != 0, st);
    });
    return rc;
}

extern "C" int ngp_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                        const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                        uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                        float* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                                        int dtype, uint32_t flags, ngp_stream_t stream) {
    (void)flags;
    if (dtype < NGP_F32 || dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (B == 0 || L == 0 || max_level == 0) return NGP_OK;
    if (!grad || !inputs || !embeddings || !offsets || !grad_embeddings) return NGP_ERR_NULL;
    if (gridtype > 1 || interp > 1 || max_level > L) return NGP_ERR_BAD_ARG;
    if (C != 1 && C != 2 && C != 4 && C != 8) return NGP_ERR_UNSUPPORTED;
    const size_t es = dtype_size(dtype);
    if (!row_aligned(embeddings, C, es) || !row_aligned(grad, C, es) || !row_aligned(grad_embeddings, C, es) ||
        !aligned(inputs, 4) || (grad_inputs && !aligned(grad_inputs, 4)))
        return NGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = NGP_ERR_UNSUPPORTED;
    NGP_DISPATCH_DTYPE(dtype, {
        if (D == 3) rc = launch_backward<T, 3>((const T*)grad, inputs, (const T*)embeddings, offsets, (T*)grad_embeddings, grad_inputs, B, C, L, max_level, S, H, gridtype, align_corners != 0, interp, st);
        else if (D == 2) rc = launch_backward<T, 2>((const T*)grad, inputs, (const T*)embeddings, offsets, (T*)grad_embeddings, grad_inputs, B, C, L, max_level, S, H, gridtype, align_corners != 0, interp, st);
    });
    return rc;
}

extern "C" int ngp_grid_input_backward(const void* grad, const void* dy_dx, float* grad_inputs, uint32_t B,
                                       uint32_t D, uint32_t C, uint32_t L, int dtype, ngp_stream_t stream) {
    if (dtype < NGP_F32 || dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (B == 0) return NGP_OK;
    if (!grad || !dy_dx || !grad_inputs) return NGP_ERR_NULL;
    const size_t es = dtype_size(dtype);
    if (!row_aligned(grad, C, es) || !row_aligned(dy_dx, C, es)) return NGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = NGP_ERR_UNSUPPORTED;
    NGP_DISPATCH_DTYPE(dtype, {
        if (D == 3) rc = launch_input_backward<T, 3>((const T*)grad, (const T*)dy_dx, grad_inputs, B, C, L, st);
        else if (D == 2) rc = launch_input_backward<T, 2>((const T*)grad, (const T*)dy_dx, grad_inputs, B, C, L, st);
    });
    return rc;
}

extern "C" int ngp_grid_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                             const int32_t* offsets, float weight, uint32_t B, uint32_t D,
                                             uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                             int align_corners, int dtype, ngp_stream_t stream) {
    if (dtype < NGP_F32 || dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (B == 0 || L == 0) return NGP_OK;
    if (!inputs || !embeddings || !grad || !offsets) return NGP_ERR_NULL;
    if (gridtype > 1) return NGP_ERR_BAD_ARG;
    const size_t es = dtype_size(dtype);
    if (!row_aligned(embeddings, C, es) || !row_aligned(grad, C, es)) return NGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = NGP_ERR_UNSUPPORTED;
    NGP_DISPATCH_DTYPE(dtype, {
        if (D == 3) rc = launch_tv<T, 3>((const T*)inputs, (const T*)embeddings, (T*)grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st);
        else if (D == 2) rc = launch_tv<T, 2>((const T*)inputs, (const T*)embeddings, (T*)grad, offsets, weight, B, C, L, S, H, gridtype, align_corners != 0, st);
    });
    return rc;
}

extern "C" int ngp_grid_grad_weight_decay(const void* embeddings, void* grad, const int32_t* offsets, float weight,
                                          uint32_t n_entries, uint32_t C, uint32_t L, int dtype,
                                          ngp_stream_t stream) {
    if (dtype < NGP_F32 || dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (n_entries == 0) return NGP_OK;
    if (!embeddings || !grad || !offsets) return NGP_ERR_NULL;
    if (L == 0 || L > 1024) return NGP_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n = n_entries * C;
    const uint32_t blocks = min(div_up(n, 256u), (uint32_t)(kNumSMs * 8));
    NGP_DISPATCH_DTYPE(dtype, {
        grid_wd_kernel<T><<<blocks, 256, (L + 1) * sizeof(int), st>>>((const T*)embeddings, (T*)grad, offsets, weight, n, L, C);
    });
    return finish_launch();
}

extern "C" int ngp_grid_level_resolutions(uint32_t L, float S, uint32_t H, uint32_t* out_dev, ngp_stream_t stream) {
    if (L == 0) return NGP_OK;
    if (!out_dev) return NGP_ERR_NULL;
    level_resolution_kernel<<<div_up(L, 64u), 64, 0, (cudaStream_t)stream>>>(L, S, H, out_dev);
    return finish_launch();
}
