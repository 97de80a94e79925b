// grid_core.cuh -- device helpers of the multiresolution grid shared by grid_encode.cu and field.cu
// (index map, level geometry, cell location, paired vector reductions).
#pragma once
#include "common.cuh"

namespace ngp {
namespace gridcore {

// Spatial hash of tcnn / torch-ngp: xor of coordinate * prime (first prime is 1, which keeps x-neighbours
// in the same sector most of the time).  gridencoder.cu:45-58.
template <uint32_t D>
__device__ __forceinline__ uint32_t coherent_prime_hash(const uint32_t (&p)[D]) {
    constexpr uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t h = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) h ^= p[d] * kPrimes[d];
    return h;
}

// Entry index (row number inside the level) of an integer grid position.  gridencoder.cu:61-79:
// dense strides are accumulated only while stride <= hashmap_size; the level is hashed iff
// gridtype == hash and the (possibly truncated) stride product exceeds hashmap_size.
template <uint32_t D>
__device__ __forceinline__ uint32_t entry_index(uint32_t gridtype, uint32_t hashmap_size, uint32_t res,
                                                const uint32_t (&p)[D]) {
    uint32_t stride = 1, idx = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (stride <= hashmap_size) {
            idx += p[d] * stride;
            stride *= res;
        }
    }
    if (gridtype == 0 && stride > hashmap_size) idx = coherent_prime_hash<D>(p);
    // idx % hashmap_size; dense levels never wrap and hashed levels are powers of two in practice.
    if (idx >= hashmap_size) idx = ((hashmap_size & (hashmap_size - 1)) == 0) ? (idx & (hashmap_size - 1)) : (idx % hashmap_size);
    return idx;
}

// Per-level resolution exactly as the device code of the reference computes it (fp32, gridencoder.cu:133).
__device__ __forceinline__ uint32_t level_resolution(uint32_t level, float S, uint32_t H) {
    return (uint32_t)ceilf(exp2f(level * S) * H);
}

__device__ __forceinline__ float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float smoothstep_df(float v) { return 6 * v * (1.0f - v); }

// Position of the sample inside level `res`: integer base corner, fractional offset (after optional
// smoothstep) and d(frac)/d(pos).  gridencoder.cu:140-160.  Returns false if the point is outside [0,1]^D.
template <uint32_t D>
__device__ __forceinline__ bool locate(const float* __restrict__ x, uint32_t res, bool align_corners, uint32_t interp,
                                       uint32_t (&base)[D], float (&frac)[D], float (&dfrac)[D]) {
    float xin[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        xin[d] = __ldg(x + d);
        if (xin[d] < 0 || xin[d] > 1) oob = true;
    }
    if (oob) return false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        float p;
        if (align_corners) {
            p = xin[d] * (float)(res - 1);
            base[d] = min((uint32_t)floorf(p), res - 2);
        } else {
            p = fminf(fmaxf(xin[d] * (float)res - 0.5f, 0.0f), (float)(res - 1));
            base[d] = (uint32_t)floorf(p);
        }
        p -= (float)base[d];
        if (interp == 1) {
            dfrac[d] = smoothstep_df(p);
            frac[d] = smoothstep_f(p);
        } else {
            dfrac[d] = 1.0f;
            frac[d] = p;
        }
    }
    return true;
}

// ---- backward helpers ------------------------------------------------------------------------------------------
// Two corners that are x-neighbours usually live in the same 16-byte block of the table (dense levels: consecutive
// rows; hashed levels: the first hash prime is 1, so x and x+1 differ only in the low row bits unless x ends in
// ...11).  scatter_pair sends both contributions with ONE 16-byte vector reduction when they share a block
// (red.global.add.noftz.v4.f16x2 = 4 rows of an fp16 F=2 table, red.global.add.v4.f32 = 2 rows of an fp32 one)
// and falls back to one packed reduction per corner otherwise.
template <typename T, uint32_t C>
__device__ __forceinline__ void scatter_pair(T* glvl, uint32_t row0, uint32_t row1, const float (&v0)[C], const float (&v1)[C]) {
    if constexpr (C == 2 && sizeof(T) == 2) {
        if ((row0 >> 2) == (row1 >> 2)) {
            const uint32_t a = row0 & 3u, b = row1 & 3u;
            uint32_t p0, p1;
            if (a == b) {
                p0 = std::is_same<T, __half>::value ? pack_h2(v0[0] + v1[0], v0[1] + v1[1]) : pack_bf2(v0[0] + v1[0], v0[1] + v1[1]);
                p1 = 0u;
            } else {
                p0 = std::is_same<T, __half>::value ? pack_h2(v0[0], v0[1]) : pack_bf2(v0[0], v0[1]);
                p1 = std::is_same<T, __half>::value ? pack_h2(v1[0], v1[1]) : pack_bf2(v1[0], v1[1]);
            }
            uint32_t w[4];
#pragma unroll
            for (uint32_t s = 0; s < 4; s++) w[s] = (s == a ? p0 : 0u) | ((s == b && a != b) ? p1 : 0u);
            T* blk = glvl + (size_t)(row0 >> 2) * 8;
            if constexpr (std::is_same<T, __half>::value) red_add_v4_h2(blk, w[0], w[1], w[2], w[3]);
            else red_add_v4_bf2(blk, w[0], w[1], w[2], w[3]);
            return;
        }
    } else if constexpr (C == 2 && sizeof(T) == 4) {
        if ((row0 >> 1) == (row1 >> 1)) {
            const uint32_t a = row0 & 1u, b = row1 & 1u;
            float f[4];
            if (a == b) {
                f[2 * a] = v0[0] + v1[0]; f[2 * a + 1] = v0[1] + v1[1];
                f[2 * (a ^ 1)] = 0.f; f[2 * (a ^ 1) + 1] = 0.f;
            } else {
                f[2 * a] = v0[0]; f[2 * a + 1] = v0[1];
                f[2 * b] = v1[0]; f[2 * b + 1] = v1[1];
            }
            red_add_v4_f32(reinterpret_cast<float*>(glvl) + (size_t)(row0 >> 1) * 4, f[0], f[1], f[2], f[3]);
            return;
        }
    }
    red_add_row<T, C>(glvl + (size_t)row0 * C, v0);
    red_add_row<T, C>(glvl + (size_t)row1 * C, v1);
}


}  // namespace gridcore
}  // namespace ngp
