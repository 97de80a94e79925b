// mlp_core.cuh -- tile constants and shared-memory staging helpers shared by mlp.cu and field.cu.
#pragma once
#include "common.cuh"
#include "tcgen05.cuh"

namespace ngp {
namespace mlpcore {

constexpr uint32_t kTile = 128;          // samples per tile == UMMA M == threads per CTA
constexpr uint32_t kPanel = kTile * 16;  // bytes of one 8-column panel of a 128-row tile
constexpr uint32_t kMaxLayers = 4;

struct MlpArgs {
    const __half* w[kMaxLayers];   // [dims[l+1], dims[l]] row-major fp16
    __half* acts[kMaxLayers];      // forward: hidden activations out; backward: hidden activations in
    float* dw[kMaxLayers];         // backward: [dims[l+1], dims[l]] fp32, accumulated with atomics
    uint32_t dims[kMaxLayers + 1];
    uint32_t n_layers;
};

// weights -> shared memory in row-panel layout (R = N_l rows)
__device__ __forceinline__ void load_weight_tile(uint8_t* dst, const __half* __restrict__ w, uint32_t N, uint32_t K) {
    const uint32_t chunks = K / 8;
    for (uint32_t i = threadIdx.x; i < N * chunks; i += blockDim.x) {
        const uint32_t n = i / chunks, c = i - n * chunks;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + c * 8));
        *reinterpret_cast<uint4*>(dst + (size_t)c * (N * 16) + n * 16) = v;
    }
}

// one row of a [M, F] fp16 matrix -> this thread's row of a 128-row tile (cp.async, zero-fill past M)
__device__ __forceinline__ void load_row_tile(uint8_t* tile, const __half* __restrict__ src, uint32_t ld, uint32_t F,
                                              uint32_t row, uint32_t M) {
    const uint32_t t = threadIdx.x;
    if (row < M) {
        const __half* p = src + (size_t)row * ld;
        for (uint32_t c = 0; c < F / 8; c++) tc::cp_async16(tc::smem_u32(tile + c * kPanel + t * 16), p + c * 8);
    } else {
        for (uint32_t c = 0; c < F / 8; c++) *reinterpret_cast<uint4*>(tile + c * kPanel + t * 16) = make_uint4(0, 0, 0, 0);
    }
}

__device__ __forceinline__ void pack16(const float (&v)[16], uint4& lo, uint4& hi) {
    __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    lo = *reinterpret_cast<uint4*>(&h[0]);
    hi = *reinterpret_cast<uint4*>(&h[4]);
}


}  // namespace mlpcore
}  // namespace ngp
