// field_bwd_ws.cu -- grid_mlp backward + hash-table gradient scatter as ONE warp-specialised persistent kernel.
//
//     [d sigma, d in2[:, :15]] --MLP warps--> grid_mlp backward (tcgen05: dW in TMEM, dH chain) --> d enc tile
//                                                                  [ring in shared memory] --scatter warps--> table gradient
//
// One CTA per SM, 24 warps:
//   * warps 16-19 and 20-23 (two MLP groups, alternating tiles) own the tensor-core chain of a tile: the saved activations (enc, h1, h2 in the tile-panel
//     layout written by field_ws.cu) arrive by BULK ASYNC COPIES (cp.async.bulk, one per tensor and tile, mbarrier
//     complete_tx) into a double-buffered set of shared-memory tiles, so the next tile streams in while this one is being
//     processed.  Per layer: dW_l^T += in_l^T dZ_l (accumulators of all layers stay in TMEM for the whole kernel) and
//     dH = dZ_l W_l, ReLU-masked into the next dZ.  The last dH is d enc: it is rounded to fp16 and handed to the scatter
//     warps through a 2-deep ring.
//   * warps 0-15 only scatter: thread (row, g) takes levels g, g+4, ... of its sample; runs of consecutive samples in the
//     same cell are merged by a segmented shuffle reduction in packed fp16x2 and only run heads issue
//     red.global.add.noftz.v4.f16x2 (x-neighbour corners share one 16-byte reduction).  The scatter is the throughput
//     bound of the backward pass; the ring keeps these warps busy while the tensor-core chain of the next tile runs.
#include "field_core.cuh"

namespace ngp {
namespace {

using namespace mlpcore;
using namespace gridcore;
using namespace fieldcore;

constexpr uint32_t kScatterThreads = 512;
constexpr uint32_t kScatterGroups = kScatterThreads / kTile;
constexpr uint32_t kBwsGroups = 2;                             // MLP groups; group g owns input stage g, d enc stage g, TMEM half g
constexpr uint32_t kBwsThreads = kScatterThreads + kBwsGroups * kTile;     // 768
constexpr uint32_t kInStages = kBwsGroups;
constexpr uint32_t kEncStages = kBwsGroups;
constexpr uint32_t kGroupCols = 256;
constexpr uint32_t kBwsTmemCols = kBwsGroups * kGroupCols;
constexpr uint32_t kBwsLayers = 3;

// control block (byte offsets from ctrl_off)
constexpr uint32_t kInFull = 0, kInEmpty = kInFull + 8 * kInStages, kEncFull = kInEmpty + 8 * kInStages,
                   kEncEmpty = kEncFull + 8 * kEncStages, kDone = kEncEmpty + 8 * kEncStages, kSlot = kDone + 8 * kBwsGroups;
constexpr uint32_t kBLevels = (kSlot + 4 + 15) & ~15u;
constexpr uint32_t kBPlans = kBLevels + kMaxLevels * sizeof(LevelConst);
constexpr uint32_t kBCtrlBytes = kBPlans + kInStages * kBwsLayers * 2 * sizeof(MmaPlan);

struct BwsArgs {
    const float* xyzs; const float* d_sigma; const float* sigma;
    const __half* d_in2; uint32_t ld2;          // tiled [tiles][ld2 / 8][128][8]
    const __half* in[kBwsLayers];               // tiled enc, h1, h2
    GridArgs g;
    const __half* w[kBwsLayers];
    float* dw[kBwsLayers];
    uint32_t dims[kBwsLayers + 1];
    uint32_t M; const int* m_dev;
    __half* grad_table;
    int density_act; float beta;
    uint32_t w_off[kBwsLayers], in_off[kBwsLayers], in_stage_bytes, dz_off, dz_bytes, enc_off, enc_stage_bytes, ctrl_off;
    uint32_t acc_col[kBwsLayers];
};

__global__ void __launch_bounds__(kBwsThreads, 1)
field_backward_ws_kernel(const BwsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t M = a.M;
    if (a.m_dev) M = min(M, (uint32_t)__ldg(a.m_dev));
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t* ctrl = smem + a.ctrl_off;
    const uint32_t in_full = tc::smem_u32(ctrl + kInFull), in_empty = tc::smem_u32(ctrl + kInEmpty);
    const uint32_t enc_full = tc::smem_u32(ctrl + kEncFull), enc_empty = tc::smem_u32(ctrl + kEncEmpty);
    const uint32_t done = tc::smem_u32(ctrl + kDone);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kSlot);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(ctrl + kBLevels);
    MmaPlan* plans = reinterpret_cast<MmaPlan*>(ctrl + kBPlans);      // [stage][layer][0 = dW, 1 = dH]
    const uint32_t L = kBwsLayers;

    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kBwsTmemCols);
    if (threadIdx.x == 32) {
        for (uint32_t s = 0; s < kInStages; s++) { tc::mbar_init(in_full + 8 * s, 1); tc::mbar_init(in_empty + 8 * s, kTile); }
        for (uint32_t s = 0; s < kEncStages; s++) { tc::mbar_init(enc_full + 8 * s, kTile); tc::mbar_init(enc_empty + 8 * s, kScatterThreads); }
        for (uint32_t gI = 0; gI < kBwsGroups; gI++) tc::mbar_init(done + 8 * gI, 1);
    }
    for (uint32_t l = 0; l < L; l++) load_weight_tile(smem + a.w_off[l], a.w[l], a.dims[l + 1], a.dims[l]);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kInStages * L) {
        const uint32_t i = threadIdx.x - 64, st = i / L, l = i % L, K = a.dims[l], N = a.dims[l + 1];
        // dZ of layer l sits in the ping-pong buffer (L - 1 - l) & 1
        const uint32_t dz_saddr = tc::smem_u32(smem + a.dz_off + (2 * st + ((L - 1 - l) & 1u)) * a.dz_bytes);
        const uint32_t in_saddr = tc::smem_u32(smem + a.in_off[l] + st * a.in_stage_bytes), w_saddr = tc::smem_u32(smem + a.w_off[l]);
        MmaPlan& dw = plans[(st * L + l) * 2];       // dW_l^T [K x N] += in_l^T [K x 128] * dZ_l [128 x N]  (MN-major views of row tiles)
        dw.idesc = tc::instr_desc(kTile, N, true, true);
        dw.n_steps = kTile / 16; dw.d_col = st * kGroupCols + a.acc_col[l]; dw.pad = 0;
        for (uint32_t ks = 0; ks < kTile / 16; ks++) {
            dw.step[ks].a = tc::smem_desc(in_saddr + ks * 256, 128, kPanel);
            dw.step[ks].b = tc::smem_desc(dz_saddr + ks * 256, 128, kPanel);
        }
        MmaPlan& dh = plans[(st * L + l) * 2 + 1];   // dH [128 x K] = dZ_l [128 x N] * W_l [N x K]
        dh.idesc = tc::instr_desc(kTile, K, false, true);
        dh.n_steps = N / 16; dh.d_col = st * kGroupCols; dh.pad = 0;
        for (uint32_t ks = 0; ks < N / 16; ks++) {
            dh.step[ks].a = tc::smem_desc(dz_saddr + ks * 2 * kPanel, kPanel, 128);
            dh.step[ks].b = tc::smem_desc(w_saddr + ks * 256, 128, N * 16);
        }
    }
    load_level_consts(s_lv, a.g);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    const GridArgs& g = a.g;
    const uint32_t F = a.dims[0];

    if (warp < kScatterThreads / 32) {
        // ================================ scatter warps ================================
        const uint32_t r = threadIdx.x & (kTile - 1), grp = threadIdx.x / kTile;
        for (uint32_t it = 0;; it++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t e = it % kEncStages;
            const uint32_t row = tile * kTile + r;
            const bool live = row < M;
            float x[3] = {2.f, 2.f, 2.f};
            if (live) unit_cube(a.xyzs + (size_t)row * 3, g.bound, x);
            tc::mbar_wait(enc_full + 8 * e, (it / kEncStages) & 1u);
            // this thread's d enc values (levels grp, grp + 4, ...) leave the ring at once, which frees the stage early
            const uint8_t* de = smem + a.enc_off + e * a.enc_stage_bytes;
            __half2 gh[kMaxLevels / kScatterGroups];
#pragma unroll
            for (uint32_t j = 0; j < kMaxLevels / kScatterGroups; j++) {
                const uint32_t level = grp + j * kScatterGroups;
                if (level < g.L) gh[j] = *reinterpret_cast<const __half2*>(de + (level / 4) * kPanel + r * 16 + (level % 4) * 4);
            }
            tc::mbar_arrive(enc_empty + 8 * e);
#pragma unroll
            for (uint32_t j = 0; j < kMaxLevels / kScatterGroups; j++) {
                const uint32_t level = grp + j * kScatterGroups;
                if (level < g.L) scatter_level(g, s_lv[level], level, x, live, gh[j], a.grad_table, lane);
            }
        }
    } else {
        // ================================ MLP groups ================================
        const uint32_t gI = (warp - kScatterThreads / 32) / 4;           // group: tiles gI, gI + 2, ...; stage gI everywhere
        const uint32_t tg = threadIdx.x - kScatterThreads - gI * kTile;  // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + (((warp & 3u) * 32u) << 16) + gI * kGroupCols;
        const uint32_t st = gI, e = gI;
        const uint32_t done_g = done + 8 * gI;
        uint8_t* dz_base = smem + a.dz_off + 2 * gI * a.dz_bytes;
        uint32_t in_bytes = 0;
        for (uint32_t l = 0; l < L; l++) in_bytes += kTile * a.dims[l] * 2;
        auto load_tile = [&](uint32_t tile, uint32_t st) {               // one thread: bulk async copies of the saved tiles
            tc::mbar_arrive_expect_tx(in_full + 8 * st, in_bytes);
            for (uint32_t l = 0; l < L; l++)
                tc::bulk_g2s(tc::smem_u32(smem + a.in_off[l] + st * a.in_stage_bytes), a.in[l] + (size_t)tile * (a.dims[l] * kTile),
                             kTile * a.dims[l] * 2, in_full + 8 * st);
        };
        if (tg == 0) {
            const uint32_t tile = blockIdx.x + gI * gridDim.x;
            if (tile < n_tiles) load_tile(tile, st);
        }
        uint32_t ph = 0, iter = 0;
        for (uint32_t it = gI;; it += kBwsGroups, iter++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t row = tile * kTile + tg;
            const bool live = row < M;
            // d out1 = [d sigma * d act / d out0, d feat(15)]
            {
                uint4 z0 = make_uint4(0, 0, 0, 0), z1 = z0;
                if (live) {
                    const float sg = __ldg(a.sigma + row);
                    float dact;
                    if (a.density_act == 0) dact = sg;                        // trunc_exp backward: g * exp(x) (activation.py:18-21)
                    else dact = 1.0f - expf(-a.beta * sg);                    // softplus' = sigmoid(beta x) = 1 - exp(-beta y)
                    const __half d0 = __float2half_rn(__ldg(a.d_sigma + row) * dact);
                    const __half* d_tile = a.d_in2 + (size_t)tile * (a.ld2 * kTile);
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(d_tile + tg * 8));
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(d_tile + (kTile + tg) * 8));
                    // shift the 15 feature gradients up by one half and put d out0 in front
                    const uint32_t s0 = (uint32_t)__half_as_ushort(d0);
                    z0.x = s0 | (u.x << 16); z0.y = (u.x >> 16) | (u.y << 16); z0.z = (u.y >> 16) | (u.z << 16); z0.w = (u.z >> 16) | (u.w << 16);
                    z1.x = (u.w >> 16) | (v.x << 16); z1.y = (v.x >> 16) | (v.y << 16); z1.z = (v.y >> 16) | (v.z << 16); z1.w = (v.z >> 16) | (v.w << 16);
                }
                uint8_t* dzt = dz_base;
                *reinterpret_cast<uint4*>(dzt + tg * 16) = z0;
                *reinterpret_cast<uint4*>(dzt + kPanel + tg * 16) = z1;
            }
            tc::mbar_wait(in_full + 8 * st, iter & 1u);
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::named_bar_sync(1 + gI, kTile);
            uint32_t cur = 0;
            for (int l = (int)L - 1; l >= 0; l--) {
                const uint32_t K = a.dims[l];
                if (tg == 0) {
                    tc::fence_after_sync();
                    issue_plan(tmem, plans[(st * L + l) * 2], iter > 0);
                    issue_plan(tmem, plans[(st * L + l) * 2 + 1], false);
                    tc::mma_commit(done_g);
                }
                tc::mbar_wait(done_g, ph);
                ph ^= 1;
                tc::fence_after_sync();
                if (l > 0) {
                    uint8_t* nxt = dz_base + (cur ^ 1) * a.dz_bytes;
                    const uint8_t* in_tile = smem + a.in_off[l] + st * a.in_stage_bytes;
                    for (uint32_t c0 = 0; c0 < K; c0 += 16) {
                        float v[16];
                        tc::tmem_ld16(lane_addr + c0, v);
                        const uint4 m0 = *reinterpret_cast<const uint4*>(in_tile + (c0 / 8) * kPanel + tg * 16);
                        const uint4 m1 = *reinterpret_cast<const uint4*>(in_tile + (c0 / 8 + 1) * kPanel + tg * 16);
                        const __half* h0 = reinterpret_cast<const __half*>(&m0);
                        const __half* h1 = reinterpret_cast<const __half*>(&m1);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            if (!(__half2float(h0[i]) > 0.f)) v[i] = 0.f;
                            if (!(__half2float(h1[i]) > 0.f)) v[8 + i] = 0.f;
                        }
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        *reinterpret_cast<uint4*>(nxt + (c0 / 8) * kPanel + tg * 16) = lo;
                        *reinterpret_cast<uint4*>(nxt + (c0 / 8 + 1) * kPanel + tg * 16) = hi;
                    }
                    tc::fence_async_smem();
                    tc::fence_before_sync();
                    tc::named_bar_sync(1 + gI, kTile);
                } else {
                    // every MMA that reads this stage's tiles has completed: hand the stage back and refill it
                    tc::mbar_arrive(in_empty + 8 * st);
                    // d enc -> fp16 tile for the scatter warps
                    tc::mbar_wait(enc_empty + 8 * e, (iter & 1u) ^ 1u);
                    uint8_t* de = smem + a.enc_off + e * a.enc_stage_bytes;
                    for (uint32_t c0 = 0; c0 < F; c0 += 16) {
                        float v[16];
                        tc::tmem_ld16(lane_addr + c0, v);
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        *reinterpret_cast<uint4*>(de + (c0 / 8) * kPanel + tg * 16) = lo;
                        *reinterpret_cast<uint4*>(de + (c0 / 8 + 1) * kPanel + tg * 16) = hi;
                    }
                    tc::mbar_arrive(enc_full + 8 * e);
                    if (tg == 0) {
                        const uint32_t nt = blockIdx.x + (it + kBwsGroups) * gridDim.x;
                        if (nt < n_tiles) {
                            tc::mbar_wait(in_empty + 8 * st, iter & 1u);
                            load_tile(nt, st);
                        }
                    }
                    tc::fence_before_sync();
                    tc::named_bar_sync(1 + gI, kTile);      // TMEM work columns and dZ buffer 0 are rewritten by the next tile
                }
                cur ^= 1;
            }
        }
        // reduce this CTA's weight-gradient accumulators (TMEM lane i = input feature i) into global memory
        if (iter > 0) {
            tc::fence_after_sync();
            for (uint32_t l = 0; l < L; l++) {
                const uint32_t K = a.dims[l], N = a.dims[l + 1];
                for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + a.acc_col[l] + c0, v);   // warp-collective: every lane participates
                    if (tg < K) {
#pragma unroll
                        for (int i = 0; i < 16; i++) red_add_f32(a.dw[l] + (size_t)(c0 + i) * K + tg, v[i]);
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kBwsTmemCols);
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_field_backward_ws(const float* xyzs, const float* d_sigma, const float* sigma, const void* d_in2,
                                     uint32_t ld2, const void* enc, const int32_t* offsets, const float* feat_weights,
                                     float bound, float S, uint32_t H, uint32_t L, uint32_t gridtype, int align_corners,
                                     uint32_t interp, const void* const* weights, const void* const* acts,
                                     const uint32_t* dims, uint32_t M, const int32_t* m_dev, int density_act, float beta,
                                     void* grad_table, float* const* dweights, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!xyzs || !d_sigma || !sigma || !d_in2 || !enc || !offsets || !weights || !acts || !dims || !grad_table || !dweights) return NGP_ERR_NULL;
    if (L == 0 || L > kMaxLevels || L % 4 != 0 || gridtype > 1 || interp > 1 || density_act < 0 || density_act > 1) return NGP_ERR_BAD_ARG;
    if (dims[0] != 2 * L || dims[3] != 16 || ld2 < 16 || ld2 % 8) return NGP_ERR_UNSUPPORTED;
    BwsArgs a = {};
    a.xyzs = xyzs; a.d_sigma = d_sigma; a.sigma = sigma; a.d_in2 = (const __half*)d_in2; a.ld2 = ld2;
    a.g = {nullptr, offsets, feat_weights, S, bound, H, L, gridtype, interp, align_corners != 0};
    uint32_t off = 0, max_n = 0, max_k = 0, acc = 0;
    for (uint32_t l = 0; l <= kBwsLayers; l++) {
        if (dims[l] == 0 || dims[l] % 16 || dims[l] > 128) return NGP_ERR_UNSUPPORTED;
        a.dims[l] = dims[l];
    }
    for (uint32_t l = 0; l < kBwsLayers; l++) {
        if (!weights[l] || !dweights[l] || (l > 0 && !acts[l - 1])) return NGP_ERR_NULL;
        a.w[l] = (const __half*)weights[l];
        a.dw[l] = dweights[l];
        a.in[l] = (const __half*)(l == 0 ? enc : acts[l - 1]);
        if (!aligned(a.w[l], 16) || !aligned(a.in[l], 16)) return NGP_ERR_ALIGN;
        a.w_off[l] = off;
        off += dims[l] * dims[l + 1] * 2;
        max_n = std::max(max_n, dims[l + 1]);
        max_k = std::max(max_k, dims[l]);
    }
    acc = max_k;
    for (uint32_t l = 0; l < kBwsLayers; l++) { a.acc_col[l] = acc; acc += dims[l + 1]; }
    if (acc > kGroupCols) return NGP_ERR_UNSUPPORTED;
    if (!aligned(grad_table, 16) || !aligned(d_in2, 16)) return NGP_ERR_ALIGN;
    off = (off + 127) & ~127u;
    uint32_t in_stage = 0;
    for (uint32_t l = 0; l < kBwsLayers; l++) { a.in_off[l] = off + in_stage; in_stage += kTile * dims[l] * 2; }
    // the M = 128 MN-major A view of an input tile spans 16 panels (32 KiB) from the tile start: keep that inside the
    // allocation (rows past dims[l] only feed TMEM lanes that are never read); the next stage / dZ buffers follow
    a.in_stage_bytes = in_stage;
    off += kInStages * in_stage;
    a.dz_off = off; a.dz_bytes = kTile * std::max(max_n, max_k) * 2;
    off += 2 * kBwsGroups * a.dz_bytes;
    a.enc_off = off; a.enc_stage_bytes = kTile * dims[0] * 2;
    off += kEncStages * a.enc_stage_bytes;
    a.ctrl_off = off;
    const uint32_t last_in_end = a.in_off[kBwsLayers - 1] + (kInStages - 1) * in_stage + 18 * kPanel;
    const uint32_t smem_bytes = std::max(off + kBCtrlBytes, last_in_end);
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    a.M = M; a.m_dev = m_dev; a.grad_table = (__half*)grad_table; a.density_act = density_act; a.beta = beta;
    static thread_local uint32_t configured = 0;
    if (smem_bytes > configured) {
        if (cudaFuncSetAttribute(field_backward_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            set_last_cuda_error(cudaGetLastError());
            return NGP_ERR_CUDA;
        }
        configured = smem_bytes;
    }
    const uint32_t grid = std::min<uint32_t>(div_up(M, kTile), kNumSMs);
    field_backward_ws_kernel<<<grid, kBwsThreads, smem_bytes, (cudaStream_t)stream>>>(a);
    return finish_launch();
}
