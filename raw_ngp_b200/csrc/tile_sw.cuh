// tile_sw.cuh -- "swizzled row-major" shared-memory tiles for the warp-specialised field kernels.
//
// A tile holds R rows x W fp16 columns, W in {16, 32, 64}, row-major with rows of 2W = 32 / 64 / 128 bytes, and the
// 16-byte chunks of a row XOR-swizzled exactly like the UMMA canonical layouts SWIZZLE_32B / 64B / 128B
// (Swizzle<B,4,3>: address bits [4, 4+B) ^= bits [7, 7+B)):
//        chunk j of row r lives at  r * 2W + ((j ^ s(r)) << 4),   s(r) = (r >> (3 - B)) & (W / 8 - 1),  B = log2(W / 8)
// The same bytes are a K-major operand (MN = row, K = column: forward layers, dH) and an MN-major operand (MN = column,
// K = row: the transposed products of dW) -- the 8-line swizzle atom is symmetric in the two roles -- and the tensor core
// reads them without shared-memory bank conflicts (the unswizzled interleaved layout of mlp.cu costs ~4x per MMA).
// The tile image is also the global-memory layout of the saved activations ("tile-panel" v2): a tile is one contiguous
// block, written with coalesced 16-byte stores and fetched back with a single bulk async copy.
#pragma once
#include "tcgen05.cuh"

namespace ngp {
namespace tsw {

__device__ __forceinline__ uint32_t sw_shift(uint32_t W) { return W == 64 ? 0u : (W == 32 ? 1u : 2u); }

// byte offset of chunk j (8 halves) of row r in an [R x W] tile
__device__ __forceinline__ uint32_t chunk_off(uint32_t W, uint32_t r, uint32_t j) {
    return r * (2 * W) + ((j ^ ((r >> sw_shift(W)) & (W / 8 - 1))) << 4);
}

// smem matrix descriptor of a swizzled tile (Blackwell version field = 1)
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t W) {
    const uint64_t layout_type = W == 64 ? 2ull : (W == 32 ? 4ull : 6ull);     // SWIZZLE_128B / 64B / 32B
    return tc::smem_desc(saddr, lbo_bytes, sbo_bytes) | (layout_type << 61);
}
// K-major operand (MN = row): k-step ks covers columns [16 ks, 16 ks + 16)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, uint32_t W, uint32_t ks) {
    return desc(tile_saddr + ks * 32, 16, 8 * 2 * W, W);
}
// MN-major operand (MN = column, K = row): k-step ks covers rows [16 ks, 16 ks + 16); mn_block_bytes = distance to the next
// block of W columns (only followed when the MMA's MN extent exceeds W)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, uint32_t W, uint32_t ks, uint32_t mn_block_bytes) {
    return desc(tile_saddr + ks * 16 * 2 * W, mn_block_bytes, 8 * 2 * W, W);
}

// weights [N, K] row-major fp16 in global memory -> swizzled tile with N rows
__device__ __forceinline__ void load_weight_tile(uint8_t* dst, const __half* __restrict__ w, uint32_t N, uint32_t K) {
    const uint32_t chunks = K / 8;
    for (uint32_t i = threadIdx.x; i < N * chunks; i += blockDim.x) {
        const uint32_t n = i / chunks, c = i - n * chunks;
        *reinterpret_cast<uint4*>(dst + chunk_off(K, n, c)) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + c * 8));
    }
}

__device__ __forceinline__ bool width_ok(uint32_t W) { return W == 16 || W == 32 || W == 64; }

// ---- any width that is a multiple of 16 (the light-stage view_mlp: 48-wide input, 80-wide hidden layers) ------------------
// A tile of R rows x W columns whose width is not 16 / 32 / 64 is stored as W / 16 PANELS of [R x 16] SWIZZLE_32B tiles, panel
// p = columns [16 p, 16 p + 16) at byte offset p * R * 32.  One k-step of a K-major operand is exactly one panel, and the
// MN-major view steps from panel to panel with LBO = R * 32 -- the same two descriptor forms as above with W = 16.
__device__ __forceinline__ uint32_t chunk_off_r(uint32_t W, uint32_t R, uint32_t r, uint32_t j) {
    if (width_ok(W)) return chunk_off(W, r, j);
    return (j >> 1) * (R * 32) + r * 32 + (((j & 1u) ^ ((r >> 2) & 1u)) << 4);
}
__device__ __forceinline__ uint64_t desc_kmajor_r(uint32_t tile_saddr, uint32_t W, uint32_t R, uint32_t ks) {
    if (width_ok(W)) return desc_kmajor(tile_saddr, W, ks);
    return desc(tile_saddr + ks * R * 32, 16, 8 * 32, 16);
}
__device__ __forceinline__ void load_weight_tile_r(uint8_t* dst, const __half* __restrict__ w, uint32_t N, uint32_t K) {
    const uint32_t chunks = K / 8;
    for (uint32_t i = threadIdx.x; i < N * chunks; i += blockDim.x) {
        const uint32_t n = i / chunks, c = i - n * chunks;
        *reinterpret_cast<uint4*>(dst + chunk_off_r(K, N, n, c)) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + c * 8));
    }
}

}  // namespace tsw
}  // namespace ngp
