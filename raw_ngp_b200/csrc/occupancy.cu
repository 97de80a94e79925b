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



// ---- partial update (renderer.py:853-876): H^3/4 uniformly random cells + H^3/4 random OCCUPIED cells per cascade ----------
// The reference draws them with randint / nonzero / randint / morton3D_invert / cat / rand_like (and a host sync for the size
// of the occupied list).  Here: the ids of the occupied cells are compacted straight from the density grid (two launches,
// order preserved, count on the device) and ONE kernel turns 6 uniforms per sample into (cell, jittered position).
constexpr uint32_t kOccTile = 256 * 16;

__global__ void __launch_bounds__(256)
occ_count_positive_kernel(const float* __restrict__ grid, uint32_t n, int* __restrict__ block_counts) {
    const uint32_t first = blockIdx.x * kOccTile + threadIdx.x * 16;
    uint32_t cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) cnt += (first + k < n && __ldg(grid + first + k) > 0.f) ? 1u : 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ uint32_t s_w[8];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        block_counts[blockIdx.x] = (int)t;
    }
}

__global__ void __launch_bounds__(256)
occ_write_positive_kernel(const float* __restrict__ grid, uint32_t n, const int* __restrict__ block_counts, int* __restrict__ out,
                          int* __restrict__ n_out) {
    __shared__ uint32_t s_w[8], s_red[8];
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += 256) part += (uint32_t)__ldg(block_counts + b);
    part = __reduce_add_sync(0xffffffffu, part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    const uint32_t first = blockIdx.x * kOccTile + threadIdx.x * 16;
    uint32_t mask = 0;
#pragma unroll
    for (uint32_t k = 0; k < 16; k++) mask |= (first + k < n && __ldg(grid + first + k) > 0.f) ? (1u << k) : 0u;
    const uint32_t cnt = __popc(mask), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int sh = 1; sh < 32; sh <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, sh);
        if (lane >= (uint32_t)sh) incl += u;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0, base = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) { const uint32_t c = s_w[w]; if (w < wid) before += c; total += c; base += s_red[w]; }
    uint32_t pos = base + before + incl - cnt;
#pragma unroll
    for (uint32_t k = 0; k < 16; k++)
        if (mask & (1u << k)) out[pos++] = (int)(first + k);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) n_out[0] = (int)(base + total);
}

// u [n, 6] uniforms in [0, 1): sample i < n / 2 is a uniformly random cell (u0..2 -> coordinates), sample i >= n / 2 a random
// entry of the occupied list (u0 -> index; with no occupied cell it revisits the cell of sample i - n / 2, the reference would fail
// on randint(0, 0)); u3..5 jitter the position inside the cell exactly like occ_sample_kernel.
__global__ void occ_sample_partial_kernel(const float* __restrict__ u, uint32_t n, const int* __restrict__ occ_list,
                                          const int* __restrict__ occ_count, uint32_t H, float span, float hgs,
                                          float* __restrict__ xyzs, int* __restrict__ indices_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* ui = u + (size_t)i * 6;
    const uint32_t count = (uint32_t)__ldg(occ_count);
    uint32_t cx, cy, cz, morton;
    if (i >= n / 2 && count > 0) {
        const uint32_t k = min((uint32_t)(__ldg(ui) * (float)count), count - 1);
        morton = (uint32_t)__ldg(occ_list + k);
        cx = compact3(morton); cy = compact3(morton >> 1); cz = compact3(morton >> 2);
    } else {
        const float* uc = (i >= n / 2) ? u + (size_t)(i - n / 2) * 6 : ui;      // no occupied cell: the second half revisits the uniform cells
        cx = min((uint32_t)(__ldg(uc) * (float)H), H - 1); cy = min((uint32_t)(__ldg(uc + 1) * (float)H), H - 1);
        cz = min((uint32_t)(__ldg(uc + 2) * (float)H), H - 1);
        morton = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
    }
    const float c[3] = {(float)cx, (float)cy, (float)cz};
    const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float w = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, c[a]), inv_hm1), -1.0f);
        const float jitter = __fmul_rn(__fadd_rn(__fmul_rn(__ldg(ui + 3 + a), 2.0f), -1.0f), hgs);
        xyzs[(size_t)i * 3 + a] = __fadd_rn(__fmul_rn(w, span), jitter);
    }
    indices_out[i] = (int)morton;
}

// renderer.py:716-809: a cell is "trainable" iff it lies inside the training AABB (grown by half a cell) and inside the view
// frustum of at least one training camera.  The reference evaluates this with a Python loop over 64^3 chunks x cascades x camera
// batches of torch ops (meshgrid, morton3D, batched matmul, boolean masks, index_put); here one thread owns one cell of one
// cascade, the camera poses are staged through shared memory in chunks, and the loop over cameras stops at the first that
// sees the cell.  Arithmetic follows the torch expression chain (scalars rounded to fp32 where torch rounds them; the 3-term
// dot products of the batched matmul as one multiplication + two FMAs).
constexpr uint32_t kCamChunk = 128;
struct CamRec { float r[9]; float t[3]; float cxfx, cyfy, near_z; };

__global__ void __launch_bounds__(256)
mark_untrained_kernel(float* __restrict__ density_grid, const float* __restrict__ poses, uint32_t pose_stride, uint32_t B,
                      const float* __restrict__ half_fov, uint32_t n_intr, const float* __restrict__ cam_near, float min_near,
                      const float* __restrict__ aabb, uint32_t H, uint32_t cascade, float grid_bound) {
    __shared__ CamRec s_cam[kCamChunk];
    const uint32_t H3 = H * H * H;
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;       // Morton index inside the cascade
    const uint32_t cas = blockIdx.y;
    const bool live = cell < H3;
    // python-double arithmetic of the reference (renderer.py:757-760), cast to fp32 where torch casts the scalar
    const double bound_d = fmin((double)(1u << cas), (double)grid_bound);
    const double hgs_d = bound_d / (double)H;
    const float span = (float)(bound_d - hgs_d), hgs = (float)hgs_d, hgs2 = (float)(hgs_d * 2.0);
    float w[3] = {0.f, 0.f, 0.f};
    bool in_aabb = false;
    if (live) {
        const uint32_t c[3] = {compact3(cell), compact3(cell >> 1), compact3(cell >> 2)};
        const float inv_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
        in_aabb = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            w[a] = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.0f, (float)c[a]), inv_hm1), -1.0f), span);
            in_aabb = in_aabb && (w[a] >= __fsub_rn(__ldg(aabb + a), hgs)) && (w[a] <= __fadd_rn(__ldg(aabb + 3 + a), hgs));
        }
    }
    bool seen = false;
    for (uint32_t head = 0; head < B; head += kCamChunk) {
        const uint32_t n = min(kCamChunk, B - head);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const float* p = poses + (size_t)(head + i) * pose_stride;          // row-major [3|4, 4] camera-to-world
            CamRec cr;
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int k = 0; k < 3; k++) cr.r[r * 3 + k] = __ldg(p + r * 4 + k);
                cr.t[r] = __ldg(p + r * 4 + 3);
            }
            const float* q = half_fov + (size_t)(n_intr > 1 ? head + i : 0) * 2;     // (cx / fx, cy / fy)
            cr.cxfx = __ldg(q);
            cr.cyfy = __ldg(q + 1);
            cr.near_z = cam_near ? __ldg(cam_near + head + i) : min_near;
            s_cam[i] = cr;
        }
        __syncthreads();
        if (live && in_aabb && !seen) {
            for (uint32_t i = 0; i < n; i++) {
                const CamRec& cr = s_cam[i];
                const float d0 = __fsub_rn(w[0], cr.t[0]), d1 = __fsub_rn(w[1], cr.t[1]), d2 = __fsub_rn(w[2], cr.t[2]);
                // cam = d @ R  (world -> camera: R is camera-to-world, so its transpose applies; renderer.py:771-773)
                const float x = __fmaf_rn(d2, cr.r[6], __fmaf_rn(d1, cr.r[3], __fmul_rn(d0, cr.r[0])));
                const float y = __fmaf_rn(d2, cr.r[7], __fmaf_rn(d1, cr.r[4], __fmul_rn(d0, cr.r[1])));
                const float z = -__fmaf_rn(d2, cr.r[8], __fmaf_rn(d1, cr.r[5], __fmul_rn(d0, cr.r[2])));
                if (z > cr.near_z && fabsf(x) < __fadd_rn(__fmul_rn(cr.cxfx, z), hgs2) && fabsf(y) < __fadd_rn(__fmul_rn(cr.cyfy, z), hgs2)) {
                    seen = true;
                    break;
                }
            }
        }
    }
    if (live && !(in_aabb && seen)) density_grid[(size_t)cas * H3 + cell] = -1.0f;
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

extern "C" int ngp_mark_untrained_grid(float* density_grid, const float* poses, uint32_t pose_stride, uint32_t B,
                                       const float* half_fov, uint32_t n_intr, const float* cam_near, float min_near,
                                       const float* aabb, uint32_t H, uint32_t cascade, float grid_bound, ngp_stream_t stream) {
    if (!density_grid || !poses || !half_fov || !aabb) return NGP_ERR_NULL;
    if (H < 2 || H > 1024 || cascade == 0 || cascade > 16 || pose_stride < 12 || (n_intr != 1 && n_intr != B)) return NGP_ERR_BAD_ARG;
    const uint32_t H3 = H * H * H;
    mark_untrained_kernel<<<dim3(div_up(H3, 256u), cascade), 256, 0, (cudaStream_t)stream>>>(density_grid, poses, pose_stride, B, half_fov,
                                                                                            n_intr, cam_near, min_near, aabb, H, cascade,
                                                                                            grid_bound);
    return finish_launch();
}

extern "C" int ngp_occ_sample_partial(const float* density_grid_cas, uint32_t H, float bound, const float* u, uint32_t n,
                                      int32_t* occ_list, int32_t* occ_count, int32_t* workspace, float* xyzs_out,
                                      int32_t* indices_out, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!density_grid_cas || !u || !occ_list || !occ_count || !workspace || !xyzs_out || !indices_out) return NGP_ERR_NULL;
    if (H < 2 || H > 1024 || (n & 1u)) return NGP_ERR_BAD_ARG;
    const uint32_t H3 = H * H * H, blocks = div_up(H3, kOccTile);
    cudaStream_t st = (cudaStream_t)stream;
    occ_count_positive_kernel<<<blocks, 256, 0, st>>>(density_grid_cas, H3, workspace);
    occ_write_positive_kernel<<<blocks, 256, 0, st>>>(density_grid_cas, H3, workspace, occ_list, occ_count);
    const double hgs = (double)bound / (double)H;
    occ_sample_partial_kernel<<<div_up(n, 256u), 256, 0, st>>>(u, n, occ_list, occ_count, H, (float)((double)bound - hgs), (float)hgs, xyzs_out,
                                                             indices_out);
    return finish_launch();
}
