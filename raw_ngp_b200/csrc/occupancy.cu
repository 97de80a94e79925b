// occupancy.cu -- device side of NeRFRenderer.update_extra_state (nerf/renderer.py:811-897).
//
// The reference builds the density grid with a Python loop of torch ops per cascade (meshgrid, morton3D,
// scale, rand_like, index_put, boolean-mask maximum, mean().item()).  Here each stage is one kernel:
// sample positions (Morton index + cascade scaling + jitter), scatter of the queried sigmas, and a fused
// EMA-max update that also produces the clamped mean on the device so that packbits can read the threshold
// without a host sync.
#include "common.cuh"

namespace ngp {
namespace {

__device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t compact3(uint32_t v) {
    v &= 0x49249249u;
    v = (v | (v >> 2)) & 0xc30c30c3u;
    v = (v | (v >> 4)) & 0x0f00f00fu;
    v = (v | (v >> 8)) & 0xff0000ffu;
    v = (v | (v >> 16)) & 0x0000ffffu;
    return v;
}

// renderer.py:833-846 (full) / :857-876 (partial)
// Arithmetic mirrors the torch expression chain of the reference op for op (every product / sum rounded
// separately, hence the explicit __fmul_rn/__fadd_rn: no FMA contraction), so that with the same torch RNG
// stream the sample positions are the reference's.
__global__ void occ_sample_kernel(const int* __restrict__ cell_indices, const float* __restrict__ noise, uint32_t n,
                                  uint32_t H, float span, float hgs, float* __restrict__ xyzs, int* __restrict__ indices_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cx, cy, cz, morton;
    if (cell_indices) {
        morton = (uint32_t)__ldg(cell_indices + i);
        cx = compact3(morton); cy = compact3(morton >> 1); cz = compact3(morton >> 2);
    } else {
        // custom_meshgrid(xs, ys, zs) with indexing='ij', flattened: x slowest, z fastest
        cz = i % H; cy = (i / H) % H; cx = i / (H * H);
        morton = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
    }
    const float c[3] = {(float)cx, (float)cy, (float)cz};
    // torch divides a tensor by a host scalar as a * (1/b) with the reciprocal rounded to fp32
    // (BinaryDivTrueKernel.cu), so 2*c/(H-1) is mirrored as a multiplication
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float w = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, c[a]), inv_hm1), -1.0f);       // 2*c/(H-1) - 1 in [-1, 1]
        const float jitter = __fmul_rn(__fadd_rn(__fmul_rn(__ldg(noise + (size_t)i * 3 + a), 2.0f), -1.0f), hgs);
        xyzs[(size_t)i * 3 + a] = __fadd_rn(__fmul_rn(w, span), jitter);
    }
    if (indices_out) indices_out[i] = (int)morton;
}

__global__ void occ_scatter_kernel(const int* __restrict__ indices, const float* __restrict__ sigmas, uint32_t n,
                                   float* __restrict__ tmp_grid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tmp_grid[__ldg(indices + i)] = __ldg(sigmas + i);
}

// renderer.py:883-887
__global__ void __launch_bounds__(256)
occ_ema_kernel(float* __restrict__ density_grid, const float* __restrict__ tmp_grid, uint32_t n_cells, float decay,
               double* __restrict__ accum, float* __restrict__ mean_out, unsigned int* __restrict__ ticket) {
    float local = 0.f;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
        float d = density_grid[i];
        const float t = __ldg(tmp_grid + i);
        if (d >= 0 && t >= 0) {
            d = fmaxf(d * decay, t);
            density_grid[i] = d;
        }
        local += fmaxf(d, 0.f);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
    __shared__ float s_part[8];
    __shared__ bool s_last;
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < 8; w++) b += (double)s_part[w];
        atomicAdd(accum, b);
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        if (s_last) {
            const double total = atomicAdd(accum, 0.0);
            mean_out[0] = (float)(total / (double)n_cells);
            *ticket = 0;
        }
    }
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_occ_sample_positions(const int32_t* cell_indices, const float* noise, uint32_t n, uint32_t H,
                                        float bound, float* xyzs_out, int32_t* indices_out, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!noise || !xyzs_out) return NGP_ERR_NULL;
    if (H < 2 || H > 1024) return NGP_ERR_BAD_ARG;
    // python-double arithmetic of the reference (renderer.py:841-844), cast to fp32 where torch casts the scalar
    const double hgs = (double)bound / (double)H;
    const double span = (double)bound - hgs;
    occ_sample_kernel<<<div_up(n, 256u), 256, 0, (cudaStream_t)stream>>>(cell_indices, noise, n, H, (float)span, (float)hgs, xyzs_out, indices_out);
    return finish_launch();
}

extern "C" int ngp_occ_scatter_sigmas(const int32_t* indices, const float* sigmas, uint32_t n, float* tmp_grid,
                                      ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!indices || !sigmas || !tmp_grid) return NGP_ERR_NULL;
    occ_scatter_kernel<<<div_up(n, 256u), 256, 0, (cudaStream_t)stream>>>(indices, sigmas, n, tmp_grid);
    return finish_launch();
}

extern "C" int ngp_occ_ema_update(float* density_grid, const float* tmp_grid, uint32_t n_cells, float decay,
                                  double* accum, float* mean_out, ngp_stream_t stream) {
    if (n_cells == 0) return NGP_OK;
    if (!density_grid || !tmp_grid || !accum || !mean_out) return NGP_ERR_NULL;
    if (!aligned(accum, 16)) return NGP_ERR_ALIGN;
    // accum is fp64[2]: [0] running sum, [1] reused as the block ticket (both zero-filled by the caller)
    const uint32_t blocks = min(div_up(n_cells, 256u), (uint32_t)(kNumSMs * 8));
    occ_ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(density_grid, tmp_grid, n_cells, decay, accum, mean_out,
                                                             reinterpret_cast<unsigned int*>(accum + 1));
    return finish_launch();
}
