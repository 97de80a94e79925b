// mlp.cu -- fully fused bias-free ReLU MLP (network.py:12-35) on the 5th-generation tensor cores.
//
// One CTA (128 threads) owns a tile of 128 samples.  Every layer is a tcgen05.mma with M = 128 (samples on the
// TMEM lanes), the fp32 accumulator lives in TMEM, and the epilogue thread of each sample row reads its row with
// tcgen05.ld, applies ReLU, converts to fp16 and writes it back to shared memory as the A operand of the next
// layer -- hidden activations never leave the SM in the forward pass (they are only streamed out once, for the
// backward pass).  The backward kernel keeps the weight-gradient accumulators of all layers resident in TMEM
// across the CTA's whole persistent loop and reduces them to global memory once at the end.
//
// The reference runs these layers as separate cuBLAS GEMMs + elementwise kernels (nn.Linear, F.relu), with the
// [M, 64] activations round-tripping through HBM between them.
#include "common.cuh"
#include "tcgen05.cuh"
#include "mlp_core.cuh"

namespace ngp {
namespace {

using namespace mlpcore;

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kFwdTmemCols = 128;

__global__ void __launch_bounds__(kTile)
mlp_forward_kernel(const __half* __restrict__ x, uint32_t ldx, MlpArgs p, uint32_t M, __half* __restrict__ y, uint32_t ldy,
                   float* __restrict__ rgb_out, int head_act, uint32_t a_tile_off, uint32_t ctrl_off,
                   const int* __restrict__ m_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t t = threadIdx.x, warp = t >> 5;
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));   // sample count produced on the device (no host sync)
    uint8_t* a_tile = smem + a_tile_off;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + ctrl_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ctrl_off + 8);

    uint32_t w_off[kMaxLayers];
    {
        uint32_t o = 0;
        for (uint32_t l = 0; l < p.n_layers; l++) { w_off[l] = o; o += p.dims[l] * p.dims[l + 1] * 2; }
    }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kFwdTmemCols);
    if (t == 0) tc::mbar_init(tc::smem_u32(mbar), 1);
    for (uint32_t l = 0; l < p.n_layers; l++) load_weight_tile(smem + w_off[l], p.w[l], p.dims[l + 1], p.dims[l]);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + ((warp * 32u) << 16);
    const uint32_t a_saddr = tc::smem_u32(a_tile), mbar_saddr = tc::smem_u32(mbar);

    uint32_t phase = 0;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kTile + t;
        load_row_tile(a_tile, x, ldx, p.dims[0], row, M);
        tc::cp_async_wait_all();
        tc::fence_async_smem();
        __syncthreads();
        for (uint32_t l = 0; l < p.n_layers; l++) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            if (t == 0) {
                tc::fence_after_sync();
                const uint32_t idesc = tc::instr_desc(kTile, N, false, false);
                const uint32_t w_saddr = tc::smem_u32(smem + w_off[l]);
                for (uint32_t ks = 0; ks < K / 16; ks++) {
                    const uint64_t ad = tc::smem_desc(a_saddr + ks * 2 * kPanel, kPanel, 128);
                    const uint64_t bd = tc::smem_desc(w_saddr + ks * 2 * (N * 16), N * 16, 128);
                    tc::mma_f16_ss(tmem, ad, bd, idesc, ks > 0);
                }
                tc::mma_commit(mbar_saddr);
            }
            tc::mbar_wait(mbar_saddr, phase);
            phase ^= 1;
            tc::fence_after_sync();
            const bool last = (l + 1 == p.n_layers);
            for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                tc::tmem_ld16(lane_addr + c0, v);
                if (!last) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
                }
                uint4 lo, hi;
                pack16(v, lo, hi);
                if (!last) {
                    *reinterpret_cast<uint4*>(a_tile + (c0 / 8) * kPanel + t * 16) = lo;
                    *reinterpret_cast<uint4*>(a_tile + (c0 / 8 + 1) * kPanel + t * 16) = hi;
                    if (p.acts[l] && row < M) {
                        uint4* g = reinterpret_cast<uint4*>(p.acts[l] + (size_t)row * N + c0);
                        g[0] = lo; g[1] = hi;
                    }
                } else if (row < M) {
                    if (head_act == 0) {
                        uint4* g = reinterpret_cast<uint4*>(y + (size_t)row * ldy + c0);
                        g[0] = lo; g[1] = hi;
                    } else if (c0 == 0) {
                        // colour head (network.py:131-138): fp16 linear output, `color - 5` in fp16, exp in fp32
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float o = __half2float(__float2half_rn(v[c]));
                            float r;
                            if (head_act == 2) r = __half2float(__float2half_rn(1.0f / (1.0f + expf(-o))));
                            else {
                                r = expf(__half2float(__float2half_rn(o - 5.0f)));
                                if (head_act == 3) r = fminf(r, 5.0f);
                            }
                            rgb_out[(size_t)row * 3 + c] = r;
                        }
                    }
                }
            }
            if (!last) tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
        }
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kFwdTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kBwdTmemCols = 256;

// 512 threads per CTA, thread (row, grp) like field_backward_density_kernel: row = (warp % 4) * 32 + lane is the sample (TMEM
// lane), grp = warp / 4 takes every fourth 16-byte chunk of the saved-activation loads and every fourth 16-column block of the
// epilogues.  (With 128 threads the per-tile chain -- 26 cp.async per thread, three epilogues of up to 80 columns -- ran on
// 4-8 warps per SM: 6 % of the warp slots.)
constexpr uint32_t kMlpBwdThreads = 512, kMlpBwdGroups = kMlpBwdThreads / kTile;

__device__ __forceinline__ void load_row_tile_g(uint8_t* tile, const __half* __restrict__ src, uint32_t ld, uint32_t F, uint32_t row, uint32_t M,
                                                uint32_t t, uint32_t grp) {
    if (row < M) {
        const __half* p = src + (size_t)row * ld;
        for (uint32_t c = grp; c < F / 8; c += kMlpBwdGroups) tc::cp_async16(tc::smem_u32(tile + c * kPanel + t * 16), p + c * 8);
    } else {
        for (uint32_t c = grp; c < F / 8; c += kMlpBwdGroups) *reinterpret_cast<uint4*>(tile + c * kPanel + t * 16) = make_uint4(0, 0, 0, 0);
    }
}

__global__ void __launch_bounds__(kMlpBwdThreads, 2)
mlp_backward_kernel(const __half* __restrict__ dy, uint32_t lddy, const __half* __restrict__ x, uint32_t ldx, MlpArgs p,
                    uint32_t M, __half* __restrict__ dx, uint32_t lddx, const float* __restrict__ d_rgb,
                    const float* __restrict__ rgb, int head_act, uint32_t dz_off, uint32_t dz_bytes,
                    uint32_t w_base, uint32_t ctrl_off, const int* __restrict__ m_dev, uint32_t dz_reuse) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t t = (warp & 3u) * 32u + (threadIdx.x & 31u);     // sample row inside the tile == TMEM lane
    const uint32_t grp = warp >> 2;
    const uint32_t L = p.n_layers;
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + ctrl_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ctrl_off + 8);

    // shared memory: [in tiles l = 0..L-1 (X, H1, ...)] [dZ ping] [dZ pong] [weights] [ctrl]
    uint32_t in_off[kMaxLayers], w_off[kMaxLayers], acc_col[kMaxLayers];
    uint32_t work_cols = 0;
    {
        uint32_t o = 0, wo = w_base, col = 0;
        for (uint32_t l = 0; l < L; l++) {
            in_off[l] = o; o += kTile * p.dims[l] * 2;
            w_off[l] = wo; wo += p.dims[l] * p.dims[l + 1] * 2;
            work_cols = max(work_cols, p.dims[l]);
        }
        col = work_cols;
        for (uint32_t l = 0; l < L; l++) { acc_col[l] = col; col += p.dims[l + 1]; }
    }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kBwdTmemCols);
    if (threadIdx.x == 0) tc::mbar_init(tc::smem_u32(mbar), 1);
    for (uint32_t l = 0; l < L; l++) load_weight_tile(smem + w_off[l], p.w[l], p.dims[l + 1], p.dims[l]);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t mbar_saddr = tc::smem_u32(mbar);

    uint32_t phase = 0, iter = 0;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, iter++) {
        const uint32_t row = tile * kTile + t;
        uint32_t cur = 0;  // which dZ buffer holds dZ of the layer being processed
        // Where dZ of layer l lives.  Ping-pong between two buffers of the widest layer, or (dz_reuse: 20 KB less at 80-wide layers,
        // which is what lets a second CTA share the SM) head buffer | one private buffer | the saved-input tile of layer l + 2,
        // which is dead once that layer's MMAs have retired and its ReLU mask has been applied.
        auto dz_of = [&](int l) -> uint8_t* {
            if (!dz_reuse) return smem + dz_off + (((L - 1 - (uint32_t)l) & 1u) * dz_bytes);
            if ((uint32_t)l == L - 1) return smem + dz_off;
            if ((uint32_t)l + 2 == L) return smem + dz_off + dz_bytes;
            return smem + in_off[l + 2];
        };
        if (head_act == 0) {
            load_row_tile_g(dz_of((int)L - 1), dy, lddy, p.dims[L], row, M, t, grp);
        } else if (grp == 0) {
            // d out = d rgb * d act / d out from the activated colour (exp: rgb; clamped exp: rgb below the clamp;
            // sigmoid: rgb (1 - rgb)); columns 3.. of the padded output carry no gradient
            __align__(16) __half dz[16];
#pragma unroll
            for (int i = 0; i < 16; i++) dz[i] = __float2half_rn(0.f);
            if (row < M) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float gc = __ldg(d_rgb + (size_t)row * 3 + c), r = __ldg(rgb + (size_t)row * 3 + c);
                    float d;
                    if (head_act == 2) d = gc * r * (1.0f - r);
                    else if (head_act == 3) d = (r < 5.0f) ? gc * r : 0.f;
                    else d = gc * r;
                    dz[c] = __float2half_rn(d);
                }
            }
            uint8_t* dzt = dz_of((int)L - 1);
            for (uint32_t c = 0; c < p.dims[L] / 8; c++)
                *reinterpret_cast<uint4*>(dzt + c * kPanel + t * 16) = (c < 2) ? reinterpret_cast<const uint4*>(dz)[c] : make_uint4(0, 0, 0, 0);
        }
        load_row_tile_g(smem + in_off[0], x, ldx, p.dims[0], row, M, t, grp);
        for (uint32_t l = 1; l < L; l++) load_row_tile_g(smem + in_off[l], p.acts[l - 1], p.dims[l], p.dims[l], row, M, t, grp);
        tc::cp_async_wait_all();
        tc::fence_async_smem();
        __syncthreads();
        for (int l = (int)L - 1; l >= 0; l--) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            const bool need_dh = (l > 0) || (dx != nullptr);
            const uint32_t dz_saddr = tc::smem_u32(dz_of(l));
            if (threadIdx.x == 0) {
                tc::fence_after_sync();
                // dW_l^T [K x N] += in_l^T [K x 128] * dZ_l [128 x N]   (both operands MN-major views of row tiles)
                const uint32_t in_saddr = tc::smem_u32(smem + in_off[l]);
                const uint32_t idw = tc::instr_desc(kTile, N, true, true);
                for (uint32_t ks = 0; ks < kTile / 16; ks++) {
                    const uint64_t ad = tc::smem_desc(in_saddr + ks * 256, 128, kPanel);
                    const uint64_t bd = tc::smem_desc(dz_saddr + ks * 256, 128, kPanel);
                    tc::mma_f16_ss(tmem + acc_col[l], ad, bd, idw, (iter > 0 || ks > 0) ? 1u : 0u);
                }
                if (need_dh) {
                    // dH [128 x K] = dZ_l [128 x N] * W_l [N x K]   (A K-major, B = MN-major view of the weight tile)
                    const uint32_t w_saddr = tc::smem_u32(smem + w_off[l]);
                    const uint32_t idh = tc::instr_desc(kTile, K, false, true);
                    for (uint32_t ks = 0; ks < N / 16; ks++) {
                        const uint64_t ad = tc::smem_desc(dz_saddr + ks * 2 * kPanel, kPanel, 128);
                        const uint64_t bd = tc::smem_desc(w_saddr + ks * 256, 128, N * 16);
                        tc::mma_f16_ss(tmem, ad, bd, idh, ks > 0);
                    }
                }
                tc::mma_commit(mbar_saddr);
            }
            tc::mbar_wait(mbar_saddr, phase);
            phase ^= 1;
            tc::fence_after_sync();
            if (need_dh) {
                uint8_t* nxt = l > 0 ? dz_of(l - 1) : nullptr;
                const uint8_t* in_tile = smem + in_off[l];
                for (uint32_t c0 = grp * 16; c0 < K; c0 += kMlpBwdGroups * 16) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + c0, v);
                    if (l > 0) {
                        // ReLU mask from the saved post-activation input of layer l
                        const uint4 m0 = *reinterpret_cast<const uint4*>(in_tile + (c0 / 8) * kPanel + t * 16);
                        const uint4 m1 = *reinterpret_cast<const uint4*>(in_tile + (c0 / 8 + 1) * kPanel + t * 16);
                        const __half* h0 = reinterpret_cast<const __half*>(&m0);
                        const __half* h1 = reinterpret_cast<const __half*>(&m1);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            if (!(__half2float(h0[i]) > 0.f)) v[i] = 0.f;
                            if (!(__half2float(h1[i]) > 0.f)) v[8 + i] = 0.f;
                        }
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        *reinterpret_cast<uint4*>(nxt + (c0 / 8) * kPanel + t * 16) = lo;
                        *reinterpret_cast<uint4*>(nxt + (c0 / 8 + 1) * kPanel + t * 16) = hi;
                    } else if (row < M) {
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        uint4* g = reinterpret_cast<uint4*>(dx + (size_t)row * lddx + c0);
                        g[0] = lo; g[1] = hi;
                    }
                }
                if (l > 0) tc::fence_async_smem();
            }
            tc::fence_before_sync();
            __syncthreads();
            cur ^= 1;
        }
    }
    // reduce this CTA's weight-gradient accumulators (TMEM lane i = input feature i) into global memory
    if (iter > 0) {
        tc::fence_after_sync();
        for (uint32_t l = 0; l < L; l++) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            for (uint32_t c0 = grp * 16; c0 < N; c0 += kMlpBwdGroups * 16) {
                float v[16];
                tc::tmem_ld16(lane_addr + acc_col[l] + c0, v);   // warp-collective: every lane participates
                if (t < K) {
#pragma unroll
                    for (int i = 0; i < 16; i++) red_add_f32(p.dw[l] + (size_t)(c0 + i) * K + t, v[i]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kBwdTmemCols);
}

bool dims_ok(const uint32_t* dims, uint32_t n_layers) {
    if (n_layers < 1 || n_layers > kMaxLayers) return false;
    for (uint32_t l = 0; l <= n_layers; l++)
        if (dims[l] == 0 || dims[l] % 16 != 0 || dims[l] > 128) return false;
    return true;
}

}  // namespace
}  // namespace ngp

using namespace ngp;

static int mlp_forward_impl(const void* x, uint32_t ldx, const void* const* weights, const uint32_t* dims,
                            uint32_t n_layers, uint32_t M, int act, void* y, uint32_t ldy, void* const* acts_out,
                            float* rgb_out, int head_act, const int32_t* m_dev, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!x || !weights || !dims) return NGP_ERR_NULL;
    if (act != NGP_ACT_RELU) return NGP_ERR_UNSUPPORTED;
    if (!dims_ok(dims, n_layers)) return NGP_ERR_UNSUPPORTED;
    if (ldx < dims[0] || ldy < dims[n_layers] || ldx % 8 || ldy % 8) return NGP_ERR_BAD_ARG;
    if (!aligned(x, 16) || (y && !aligned(y, 16))) return NGP_ERR_ALIGN;
    MlpArgs p = {};
    p.n_layers = n_layers;
    uint32_t w_bytes = 0, max_k = 0;
    for (uint32_t l = 0; l < n_layers; l++) {
        if (!weights[l] || !aligned(weights[l], 16)) return weights[l] ? NGP_ERR_ALIGN : NGP_ERR_NULL;
        p.w[l] = (const __half*)weights[l];
        p.acts[l] = (acts_out && l + 1 < n_layers) ? (__half*)acts_out[l] : nullptr;
        if (p.acts[l] && !aligned(p.acts[l], 16)) return NGP_ERR_ALIGN;
        w_bytes += dims[l] * dims[l + 1] * 2;
        max_k = std::max(max_k, dims[l]);
    }
    for (uint32_t l = 0; l <= n_layers; l++) p.dims[l] = dims[l];
    const uint32_t a_off = (w_bytes + 127) & ~127u;
    const uint32_t ctrl_off = a_off + kTile * max_k * 2;
    const uint32_t smem_bytes = ctrl_off + 16;
    static thread_local SmemCache cache = {};
    if (const int rc = ensure_dynamic_smem(mlp_forward_kernel, smem_bytes, cache)) return rc;
    const uint32_t n_tiles = div_up(M, kTile);
    const uint32_t grid = std::min<uint32_t>(n_tiles, kNumSMs * 4);   // 4 CTAs/SM: 4 x 128 TMEM columns
    mlp_forward_kernel<<<grid, kTile, smem_bytes, (cudaStream_t)stream>>>((const __half*)x, ldx, p, M, (__half*)y, ldy, rgb_out, head_act, a_off, ctrl_off, m_dev);
    return finish_launch();
}

extern "C" int ngp_mlp_forward(const void* x, uint32_t ldx, const void* const* weights, const uint32_t* dims,
                               uint32_t n_layers, uint32_t M, int act, void* y, uint32_t ldy, void* const* acts_out,
                               ngp_stream_t stream) {
    if (M > 0 && !y) return NGP_ERR_NULL;
    return mlp_forward_impl(x, ldx, weights, dims, n_layers, M, act, y, ldy, acts_out, nullptr, 0, nullptr, stream);
}

extern "C" int ngp_mlp_forward_rgb(const void* x, uint32_t ldx, const void* const* weights, const uint32_t* dims,
                                   uint32_t n_layers, uint32_t M, const int32_t* m_dev, int act, int color_act,
                                   float* rgb_out, void* const* acts_out, ngp_stream_t stream) {
    if (M > 0 && !rgb_out) return NGP_ERR_NULL;
    if (color_act < 1 || color_act > 3) return NGP_ERR_BAD_ARG;
    return mlp_forward_impl(x, ldx, weights, dims, n_layers, M, act, nullptr, 16, acts_out, rgb_out, color_act, m_dev, stream);
}

static int mlp_backward_impl(const void* dy, uint32_t lddy, const void* x, uint32_t ldx, const void* const* weights,
                             const void* const* acts, const uint32_t* dims, uint32_t n_layers, uint32_t M, int act,
                             void* dx, uint32_t lddx, float* const* dweights, const float* d_rgb, const float* rgb,
                             int head_act, const int32_t* m_dev, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!x || !weights || !dims || !dweights) return NGP_ERR_NULL;
    if (head_act == 0 && !dy) return NGP_ERR_NULL;
    if (head_act != 0 && (!d_rgb || !rgb)) return NGP_ERR_NULL;
    if (n_layers > 1 && !acts) return NGP_ERR_NULL;
    if (act != NGP_ACT_RELU) return NGP_ERR_UNSUPPORTED;
    if (!dims_ok(dims, n_layers)) return NGP_ERR_UNSUPPORTED;
    if (ldx < dims[0] || lddy < dims[n_layers] || ldx % 8 || lddy % 8 || (dx && (lddx < dims[0] || lddx % 8))) return NGP_ERR_BAD_ARG;
    if (!aligned(x, 16) || (dy && !aligned(dy, 16)) || (dx && !aligned(dx, 16))) return NGP_ERR_ALIGN;
    MlpArgs p = {};
    p.n_layers = n_layers;
    uint32_t w_bytes = 0, in_bytes = 0, max_n = 0, max_k = 0, acc_cols = 0;
    for (uint32_t l = 0; l < n_layers; l++) {
        if (!weights[l] || !dweights[l]) return NGP_ERR_NULL;
        if (l + 1 < n_layers && !acts[l]) return NGP_ERR_NULL;
        p.w[l] = (const __half*)weights[l];
        p.acts[l] = (l + 1 < n_layers) ? (__half*)acts[l] : nullptr;
        p.dw[l] = dweights[l];
        w_bytes += dims[l] * dims[l + 1] * 2;
        in_bytes += kTile * dims[l] * 2;
        max_n = std::max(max_n, dims[l + 1]);
        max_k = std::max(max_k, dims[l]);
        acc_cols += dims[l + 1];
    }
    for (uint32_t l = 0; l <= n_layers; l++) p.dims[l] = dims[l];
    if (max_k + acc_cols > kBwdTmemCols) return NGP_ERR_UNSUPPORTED;
    // dZ buffers: two of the widest layer (ping-pong), or -- when every dZ_l with l <= L - 3 fits the saved-input tile of layer
    // l + 2 -- a head buffer + one private buffer, the deeper dZ tiles reusing dead input tiles (see the kernel)
    bool reuse = n_layers >= 2;
    for (uint32_t l = 0; l + 3 <= n_layers; l++) reuse = reuse && dims[l + 1] <= dims[l + 2];
    uint32_t dz_bytes = kTile * std::max(max_n, max_k) * 2, dz_total = 2 * dz_bytes;
    if (reuse) { dz_bytes = kTile * dims[n_layers] * 2; dz_total = dz_bytes + kTile * dims[n_layers - 1] * 2; }
    const uint32_t dz_off = in_bytes;
    const uint32_t w_base = dz_off + dz_total;
    const uint32_t ctrl_off = (w_base + w_bytes + 127) & ~127u;
    // the M = 128 MN-major A view of an input tile spans 16 panels (32 KiB) from the tile start: keep that inside the
    // allocation (rows past dims[l] only feed TMEM lanes that are never read)
    const uint32_t last_in_off = in_bytes - kTile * dims[n_layers - 1] * 2;
    const uint32_t smem_bytes = std::max(ctrl_off + 16, last_in_off + 16 * kPanel + 2 * kPanel);
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    static thread_local SmemCache cache = {};
    if (const int rc = ensure_dynamic_smem(mlp_backward_kernel, smem_bytes, cache)) return rc;
    const uint32_t n_tiles = div_up(M, kTile);
    const uint32_t grid = std::min<uint32_t>(n_tiles, kNumSMs * 2);   // 2 CTAs/SM: 2 x 256 TMEM columns
    mlp_backward_kernel<<<grid, kMlpBwdThreads, smem_bytes, (cudaStream_t)stream>>>((const __half*)dy, lddy, (const __half*)x, ldx, p, M,
                                                                          (__half*)dx, lddx, d_rgb, rgb, head_act, dz_off, dz_bytes, w_base, ctrl_off, m_dev, reuse ? 1u : 0u);
    return finish_launch();
}

extern "C" int ngp_mlp_backward(const void* dy, uint32_t lddy, const void* x, uint32_t ldx, const void* const* weights,
                                const void* const* acts, const uint32_t* dims, uint32_t n_layers, uint32_t M, int act,
                                void* dx, uint32_t lddx, float* const* dweights, ngp_stream_t stream) {
    return mlp_backward_impl(dy, lddy, x, ldx, weights, acts, dims, n_layers, M, act, dx, lddx, dweights, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int ngp_mlp_backward_rgb(const float* d_rgb, const float* rgb, int color_act, const void* x, uint32_t ldx,
                                    const void* const* weights, const void* const* acts, const uint32_t* dims,
                                    uint32_t n_layers, uint32_t M, const int32_t* m_dev, int act, void* dx, uint32_t lddx,
                                    float* const* dweights, ngp_stream_t stream) {
    if (color_act < 1 || color_act > 3) return NGP_ERR_BAD_ARG;
    return mlp_backward_impl(nullptr, 16, x, ldx, weights, acts, dims, n_layers, M, act, dx, lddx, dweights, d_rgb, rgb, color_act, m_dev, stream);
}
