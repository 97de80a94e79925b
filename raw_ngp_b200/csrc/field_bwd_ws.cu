// field_bwd_ws.cu -- the whole NeRF field backward as ONE warp-specialised persistent kernel:
//
//   d rgb --view group--> view_mlp backward --d in2[:, :15] ring--> grid group --> grid_mlp backward --d enc ring-->
//                                                                   d sigma ---^                  scatter warps --> table gradient
//
// One CTA per SM, 24 warps:
//   * warps 16-19, the VIEW group: colour-activation backward, then the three view_mlp layers on the tensor cores
//     (dW_l^T += in_l^T dZ_l with the accumulators of all layers resident in TMEM for the whole kernel; dH = dZ_l W_l,
//     ReLU-masked into the next dZ).  The first 16 columns of its last dH (d feat; the SH inputs need no gradient) go to
//     the grid group through a 2-deep shared-memory ring -- d in2 never touches HBM.
//   * warps 20-23, the GRID group: the same chain for grid_mlp, one tile behind the view group; its last dH is d enc,
//     handed to the scatter warps through a second ring.
//   * both groups fetch their saved activations (tile-panel layout of field_ws.cu) with BULK ASYNC COPIES, one per tensor
//     and tile, completing on an mbarrier; a tensor's slot is refilled for the next tile as soon as its layer's MMAs have
//     retired, so the loads run two layers ahead without a second set of buffers.  Per layer dH is issued before dW: the
//     epilogue only waits for dH, the dW accumulation overlaps it.
//   * warps 0-15 only scatter: thread (row, g) takes levels g, g+4, ... of its sample; runs of consecutive samples in the
//     same cell are merged by a segmented shuffle reduction in packed fp16x2 and only run heads issue reductions
//     (red.global.add.noftz.v2.f16x2 for an aligned x-neighbour pair).  The scatter is the throughput bound of the backward
//     pass; the rings keep these warps busy while the tensor-core chains of the next tiles run.
#include "field_core.cuh"
#include "tile_sw.cuh"

namespace ngp {
namespace {

using namespace mlpcore;
using namespace gridcore;
using namespace fieldcore;

constexpr uint32_t kScatterThreads = 512;
constexpr uint32_t kScatterGroups = kScatterThreads / kTile;
constexpr uint32_t kChains = 2;                                 // 0 = view_mlp, 1 = grid_mlp
constexpr uint32_t kBwsThreads = kScatterThreads + kChains * kTile;     // 768
constexpr uint32_t kRing = 2;
constexpr uint32_t kChainCols = 256;
constexpr uint32_t kBwsTmemCols = kChains * kChainCols;
constexpr uint32_t kL = 3;

// control block (byte offsets from ctrl_off)
constexpr uint32_t kTFull = 0;                                   // [chain][layer]
constexpr uint32_t kEncFull = kTFull + 8 * kChains * kL, kEncEmpty = kEncFull + 8 * kRing;
constexpr uint32_t kDinFull = kEncEmpty + 8 * kRing, kDinEmpty = kDinFull + 8 * kRing;
constexpr uint32_t kDone = kDinEmpty + 8 * kRing, kTail = kDone + 8 * kChains, kSlot = kTail + 8 * kChains;
constexpr uint32_t kBLevels = (kSlot + 4 + 15) & ~15u;           // LevelConst[L], then the MMA plans

struct Chain {
    const __half* in[kL];          // tiled saved tensors: layer inputs (in2 | enc, h1, h2)
    const __half* w[kL];
    float* dw[kL];
    uint32_t dims[kL + 1];
    uint32_t w_off[kL], in_off[kL], dz_off, dz_bytes, acc_col[kL];
};

struct BwsArgs {
    Chain c[kChains];
    const float* xyzs; const float* d_sigma; const float* sigma; const float* d_rgb; const float* rgb;
    GridArgs g;
    uint32_t M; const int* m_dev;
    __half* grad_table;
    int density_act, color_act; float beta;
    uint32_t enc_off, enc_stage_bytes, din_off, ctrl_off, plans_off;
};

__global__ void __launch_bounds__(kBwsThreads, 1)
field_backward_ws_kernel(const BwsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t M = a.M;
    if (a.m_dev) M = min(M, (uint32_t)__ldg(a.m_dev));
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t* ctrl = smem + a.ctrl_off;
    const uint32_t t_full = tc::smem_u32(ctrl + kTFull);
    const uint32_t enc_full = tc::smem_u32(ctrl + kEncFull), enc_empty = tc::smem_u32(ctrl + kEncEmpty);
    const uint32_t din_full = tc::smem_u32(ctrl + kDinFull), din_empty = tc::smem_u32(ctrl + kDinEmpty);
    const uint32_t done = tc::smem_u32(ctrl + kDone), tail = tc::smem_u32(ctrl + kTail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kSlot);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(ctrl + kBLevels);
    MmaPlan* plans = reinterpret_cast<MmaPlan*>(ctrl + a.plans_off);  // [chain][layer][0 = dW, 1 = dH]

    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kBwsTmemCols);
    if (threadIdx.x == 32) {
        for (uint32_t i = 0; i < kChains * kL; i++) tc::mbar_init(t_full + 8 * i, 1);
        for (uint32_t s = 0; s < kRing; s++) {
            tc::mbar_init(enc_full + 8 * s, kTile); tc::mbar_init(enc_empty + 8 * s, kScatterThreads);
            tc::mbar_init(din_full + 8 * s, kTile); tc::mbar_init(din_empty + 8 * s, kTile);
        }
        for (uint32_t ci = 0; ci < kChains; ci++) { tc::mbar_init(done + 8 * ci, 1); tc::mbar_init(tail + 8 * ci, 1); }
    }
    for (uint32_t ci = 0; ci < kChains; ci++)
        for (uint32_t l = 0; l < kL; l++) tsw::load_weight_tile(smem + a.c[ci].w_off[l], a.c[ci].w[l], a.c[ci].dims[l + 1], a.c[ci].dims[l]);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kChains * kL) {
        const uint32_t i = threadIdx.x - 64, ci = i / kL, l = i % kL;
        const Chain& c = a.c[ci];
        const uint32_t K = c.dims[l], N = c.dims[l + 1];
        // dZ of layer l sits in the chain's ping-pong buffer (kL - 1 - l) & 1
        const uint32_t dz_saddr = tc::smem_u32(smem + c.dz_off + ((kL - 1 - l) & 1u) * c.dz_bytes);
        const uint32_t in_saddr = tc::smem_u32(smem + c.in_off[l]), w_saddr = tc::smem_u32(smem + c.w_off[l]);
        // all operands are swizzled row-major tiles (tile_sw.cuh): saved input [128 x K], dZ [128 x N], weights [N x K]
        MmaPlan& dw = plans[i * 2];       // dW_l^T [K x N] += in_l^T [K x 128] * dZ_l [128 x N]: both MN-major views, k = sample rows
        dw.idesc = tc::instr_desc(kTile, N, true, true);
        dw.n_steps = kTile / 16; dw.d_col = ci * kChainCols + c.acc_col[l]; dw.pad = 0;
        for (uint32_t ks = 0; ks < kTile / 16; ks++) {
            // M = 128 > K: the MN blocks past the tile (stride = tile size) only feed TMEM lanes that are never read
            dw.step[ks].a = tsw::desc_mnmajor(in_saddr, K, ks, kTile * K * 2);
            dw.step[ks].b = tsw::desc_mnmajor(dz_saddr, N, ks, kTile * N * 2);
        }
        MmaPlan& dh = plans[i * 2 + 1];   // dH [128 x K] = dZ_l [128 x N] (K-major) * W_l [N x K] (MN-major view: MN = K, k = N rows)
        dh.idesc = tc::instr_desc(kTile, K, false, true);
        dh.n_steps = N / 16; dh.d_col = ci * kChainCols; dh.pad = 0;
        for (uint32_t ks = 0; ks < N / 16; ks++) {
            dh.step[ks].a = tsw::desc_kmajor(dz_saddr, N, ks);
            dh.step[ks].b = tsw::desc_mnmajor(w_saddr, K, ks, N * K * 2);
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

    if (warp < kScatterThreads / 32) {
        // ================================ scatter warps ================================
        const uint32_t r = threadIdx.x & (kTile - 1), grp = threadIdx.x / kTile;
        for (uint32_t it = 0;; it++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t e = it % kRing;
            const uint32_t row = tile * kTile + r;
            const bool live = row < M;
            float x[3] = {2.f, 2.f, 2.f};
            if (live) unit_cube(a.xyzs + (size_t)row * 3, g.bound, x);
            tc::mbar_wait(enc_full + 8 * e, (it / kRing) & 1u);
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
        // ================================ MLP chains ================================
        const uint32_t ci = (warp - kScatterThreads / 32) / 4;            // 0 = view group, 1 = grid group
        const bool view = ci == 0;
        const Chain& c = a.c[ci];
        const uint32_t tg = threadIdx.x - kScatterThreads - ci * kTile;   // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + (((warp & 3u) * 32u) << 16) + ci * kChainCols;
        const uint32_t done_c = done + 8 * ci, tail_c = tail + 8 * ci, tf = t_full + 8 * ci * kL;
        uint8_t* dz_base = smem + c.dz_off;
        const MmaPlan* pl = plans + ci * kL * 2;
        auto load_tensor = [&](uint32_t tile, uint32_t l) {               // one thread: bulk async copy of one saved tile
            const uint32_t bytes = kTile * c.dims[l] * 2;
            tc::mbar_arrive_expect_tx(tf + 8 * l, bytes);
            tc::bulk_g2s(tc::smem_u32(smem + c.in_off[l]), c.in[l] + (size_t)tile * (c.dims[l] * kTile), bytes, tf + 8 * l);
        };
        if (tg == 0 && blockIdx.x < n_tiles)
            for (uint32_t l = 0; l < kL; l++) load_tensor(blockIdx.x, l);
        // per-sample head inputs, fetched one tile ahead: (d rgb, rgb) for the view group, (d sigma, sigma) for the grid group
        float h0[3] = {0.f, 0.f, 0.f}, h1[3] = {0.f, 0.f, 0.f};
        auto fetch_head = [&](uint32_t tile) {
            const uint32_t row = tile * kTile + tg;
            h0[0] = h0[1] = h0[2] = h1[0] = h1[1] = h1[2] = 0.f;
            if (tile < n_tiles && row < M) {
                if (view) {
#pragma unroll
                    for (int k = 0; k < 3; k++) { h0[k] = __ldg(a.d_rgb + (size_t)row * 3 + k); h1[k] = __ldg(a.rgb + (size_t)row * 3 + k); }
                } else {
                    h0[0] = __ldg(a.d_sigma + row); h1[0] = __ldg(a.sigma + row);
                }
            }
        };
        fetch_head(blockIdx.x);
        uint32_t ph = 0;
        uint32_t it = 0;
        for (;; it++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t e = it % kRing, rp = (it / kRing) & 1u;
            const bool live = tile * kTile + tg < M;
            // ---- dZ of the last layer (16 columns) ----
            uint4 z0 = make_uint4(0, 0, 0, 0), z1 = z0;
            if (view) {
                if (live) {
                    // d out = d rgb * d act / d out from the activated colour (exp: rgb; clamped exp: rgb below the clamp; sigmoid:
                    // rgb (1 - rgb)); columns 3.. of the padded output carry no gradient
                    float d[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (a.color_act == 2) d[k] = h0[k] * h1[k] * (1.0f - h1[k]);
                        else if (a.color_act == 3) d[k] = (h1[k] < 5.0f) ? h0[k] * h1[k] : 0.f;
                        else d[k] = h0[k] * h1[k];
                    }
                    z0.x = pack_h2(d[0], d[1]); z0.y = pack_h2(d[2], 0.f);
                }
            } else {
                // [d sigma * d act / d out0, d feat(15)]: the feature gradients come from the view group through the ring
                tc::mbar_wait(din_full + 8 * e, rp);
                const uint8_t* di = smem + a.din_off + e * (2 * kPanel);
                const uint4 u = *reinterpret_cast<const uint4*>(di + tg * 16);
                const uint4 v = *reinterpret_cast<const uint4*>(di + kPanel + tg * 16);
                tc::mbar_arrive(din_empty + 8 * e);
                if (live) {
                    const float sg = h1[0];
                    float dact;
                    if (a.density_act == 0) dact = sg;                        // trunc_exp backward: g * exp(x) (activation.py:18-21)
                    else dact = 1.0f - expf(-a.beta * sg);                    // softplus' = sigmoid(beta x) = 1 - exp(-beta y)
                    const uint32_t s0 = (uint32_t)__half_as_ushort(__float2half_rn(h0[0] * dact));
                    // shift the 15 feature gradients up by one half and put d out0 in front
                    z0.x = s0 | (u.x << 16); z0.y = (u.x >> 16) | (u.y << 16); z0.z = (u.y >> 16) | (u.z << 16); z0.w = (u.z >> 16) | (u.w << 16);
                    z1.x = (u.w >> 16) | (v.x << 16); z1.y = (v.x >> 16) | (v.y << 16); z1.z = (v.y >> 16) | (v.z << 16); z1.w = (v.z >> 16) | (v.w << 16);
                }
            }
            *reinterpret_cast<uint4*>(dz_base + tsw::chunk_off(16, tg, 0)) = z0;
            *reinterpret_cast<uint4*>(dz_base + tsw::chunk_off(16, tg, 1)) = z1;
            fetch_head(blockIdx.x + (it + 1) * gridDim.x);      // next tile's head inputs: in flight during this tile's chain
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::named_bar_sync(1 + ci, kTile);
            const uint32_t next_tile = blockIdx.x + (it + 1) * gridDim.x;
            uint32_t cur = 0;
            for (int l = (int)kL - 1; l >= 0; l--) {
                const uint32_t K = c.dims[l];
                tc::mbar_wait(tf + 8 * l, it & 1u);              // this layer's saved tile has landed
                if (tg == 0) {
                    // dH first: it is all the epilogue waits for.  The (longer) dW accumulation is issued behind it and runs on
                    // the tensor core while the chain's warps do the epilogue; the NEXT commit covers it (MMAs retire in order).
                    tc::fence_after_sync();
                    issue_plan(tmem, pl[2 * l + 1], false);
                    tc::mma_commit(done_c);
                    issue_plan(tmem, pl[2 * l], it > 0);
                    if (l == 0) tc::mma_commit(tail_c);          // the tile's last MMA group completes on its own barrier
                }
                tc::mbar_wait(done_c, ph);
                ph ^= 1;
                tc::fence_after_sync();
                // dW of the previous layer (l + 1) has retired with this commit: its saved tile can be refilled for the next tile
                if (tg == 0 && l + 1 < (int)kL && next_tile < n_tiles) load_tensor(next_tile, (uint32_t)(l + 1));
                if (l > 0) {
                    uint8_t* nxt = dz_base + (cur ^ 1) * c.dz_bytes;
                    const uint8_t* in_tile = smem + c.in_off[l];
                    for (uint32_t c0 = 0; c0 < K; c0 += 16) {
                        float v[16];
                        tc::tmem_ld16(lane_addr + c0, v);
                        const uint32_t o0 = tsw::chunk_off(K, tg, c0 / 8), o1 = tsw::chunk_off(K, tg, c0 / 8 + 1);
                        const uint4 m0 = *reinterpret_cast<const uint4*>(in_tile + o0);
                        const uint4 m1 = *reinterpret_cast<const uint4*>(in_tile + o1);
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        // ReLU mask on packed halves: the saved activation is a ReLU output (>= +0), so "was active" == "bits != 0";
                        // __hgt2_mask gives 0xFFFF per active half and one AND zeroes the gradient of the inactive ones
                        const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
                        lo.x &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m0.x), zero2);
                        lo.y &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m0.y), zero2);
                        lo.z &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m0.z), zero2);
                        lo.w &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m0.w), zero2);
                        hi.x &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m1.x), zero2);
                        hi.y &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m1.y), zero2);
                        hi.z &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m1.z), zero2);
                        hi.w &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m1.w), zero2);
                        *reinterpret_cast<uint4*>(nxt + o0) = lo;        // dZ of layer l - 1 has the same width K
                        *reinterpret_cast<uint4*>(nxt + o1) = hi;
                    }
                    tc::fence_async_smem();
                    tc::fence_before_sync();
                    tc::named_bar_sync(1 + ci, kTile);
                } else {
                    if (view) {
                        // d in2[:, :16] -> the grid group (column 15 is an SH input: ignored there)
                        tc::mbar_wait(din_empty + 8 * e, rp ^ 1u);
                        float v[16];
                        tc::tmem_ld16(lane_addr, v);
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        uint8_t* di = smem + a.din_off + e * (2 * kPanel);
                        *reinterpret_cast<uint4*>(di + tg * 16) = lo;
                        *reinterpret_cast<uint4*>(di + kPanel + tg * 16) = hi;
                        tc::mbar_arrive(din_full + 8 * e);
                    } else {
                        // d enc -> fp16 tile for the scatter warps
                        tc::mbar_wait(enc_empty + 8 * e, rp ^ 1u);
                        uint8_t* de = smem + a.enc_off + e * a.enc_stage_bytes;
                        for (uint32_t c0 = 0; c0 < K; c0 += 16) {
                            float v[16];
                            tc::tmem_ld16(lane_addr + c0, v);
                            uint4 lo, hi;
                            pack16(v, lo, hi);
                            *reinterpret_cast<uint4*>(de + (c0 / 8) * kPanel + tg * 16) = lo;
                            *reinterpret_cast<uint4*>(de + (c0 / 8 + 1) * kPanel + tg * 16) = hi;
                        }
                        tc::mbar_arrive(enc_full + 8 * e);
                    }
                    // dW of layer 0 still reads dZ buffer 0 and the layer-0 saved tile: wait for its commit before the next tile
                    // rewrites them
                    tc::mbar_wait(tail_c, it & 1u);
                    if (tg == 0 && next_tile < n_tiles) load_tensor(next_tile, 0u);
                    tc::fence_before_sync();
                    tc::named_bar_sync(1 + ci, kTile);      // TMEM work columns and dZ buffer 0 are rewritten by the next tile
                }
                cur ^= 1;
            }
        }
        // reduce this chain's weight-gradient accumulators (TMEM lane i = input feature i) into global memory
        if (it > 0) {
            tc::fence_after_sync();
            for (uint32_t l = 0; l < kL; l++) {
                const uint32_t K = c.dims[l], N = c.dims[l + 1];
                for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + c.acc_col[l] + c0, v);   // warp-collective: every lane participates
                    if (tg < K) {
#pragma unroll
                        for (int i = 0; i < 16; i++) red_add_f32(c.dw[l] + (size_t)(c0 + i) * K + tg, v[i]);
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

extern "C" int ngp_field_backward_full(const float* xyzs, const float* d_sigma, const float* sigma, const float* d_rgb,
                                       const float* rgb, const void* enc, const void* const* grid_acts, const void* in2,
                                       const void* const* view_acts, const int32_t* offsets, const float* feat_weights,
                                       float bound, float S, uint32_t H, uint32_t L, uint32_t gridtype, int align_corners,
                                       uint32_t interp, const void* const* grid_weights, const uint32_t* grid_dims,
                                       const void* const* view_weights, const uint32_t* view_dims, uint32_t M,
                                       const int32_t* m_dev, int density_act, float beta, int color_act, void* grad_table,
                                       float* const* grid_dweights, float* const* view_dweights, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!xyzs || !d_sigma || !sigma || !d_rgb || !rgb || !enc || !grid_acts || !in2 || !view_acts || !offsets || !grid_weights ||
        !grid_dims || !view_weights || !view_dims || !grad_table || !grid_dweights || !view_dweights)
        return NGP_ERR_NULL;
    if (L == 0 || L > kMaxLevels || L % 4 != 0 || gridtype > 1 || interp > 1 || density_act < 0 || density_act > 1 ||
        color_act < 1 || color_act > 3)
        return NGP_ERR_BAD_ARG;
    if (grid_dims[0] != 2 * L || grid_dims[3] != 16 || view_dims[3] != 16 || view_dims[0] < 16) return NGP_ERR_UNSUPPORTED;
    BwsArgs a = {};
    a.xyzs = xyzs; a.d_sigma = d_sigma; a.sigma = sigma; a.d_rgb = d_rgb; a.rgb = rgb;
    a.g = {nullptr, offsets, feat_weights, S, bound, H, L, gridtype, interp, align_corners != 0};
    uint32_t off = 0;
    for (uint32_t ci = 0; ci < kChains; ci++) {
        Chain& c = a.c[ci];
        const uint32_t* dims = ci == 0 ? view_dims : grid_dims;
        const void* const* w = ci == 0 ? view_weights : grid_weights;
        float* const* dw = ci == 0 ? view_dweights : grid_dweights;
        const void* const* acts = ci == 0 ? view_acts : grid_acts;
        uint32_t max_k = 0, max_n = 0;
        for (uint32_t l = 0; l <= kL; l++) {
            // swizzled tiles of width 16 / 32 / 64 (tile_sw.cuh); wider layers use the two-kernel path of field.cu / mlp.cu
            if (dims[l] != 16 && dims[l] != 32 && dims[l] != 64) return NGP_ERR_UNSUPPORTED;
            c.dims[l] = dims[l];
        }
        for (uint32_t l = 0; l < kL; l++) {
            if (!w[l] || !dw[l] || (l > 0 && !acts[l - 1])) return NGP_ERR_NULL;
            c.w[l] = (const __half*)w[l];
            c.dw[l] = dw[l];
            c.in[l] = (const __half*)(l == 0 ? (ci == 0 ? in2 : enc) : acts[l - 1]);
            if (!aligned(c.w[l], 16) || !aligned(c.in[l], 16)) return NGP_ERR_ALIGN;
            max_k = std::max(max_k, dims[l]);
            max_n = std::max(max_n, dims[l + 1]);
        }
        uint32_t acc = max_k;
        for (uint32_t l = 0; l < kL; l++) { c.acc_col[l] = acc; acc += dims[l + 1]; }
        if (acc > kChainCols) return NGP_ERR_UNSUPPORTED;
        c.dz_bytes = kTile * std::max(max_n, max_k) * 2;
    }
    if (!aligned(grad_table, 16)) return NGP_ERR_ALIGN;
    // shared memory: weights | saved tiles of both chains | dZ ping-pong of both chains | d enc ring | d in2 ring | control.
    // The M = 128 MN-major A view of a saved tile spans 16 panels (32 KiB) from its start (rows past dims[l] only feed TMEM
    // lanes that are never read): the buffers that follow the tiles are larger than that.
    for (uint32_t ci = 0; ci < kChains; ci++)
        for (uint32_t l = 0; l < kL; l++) { a.c[ci].w_off[l] = off; off += a.c[ci].dims[l] * a.c[ci].dims[l + 1] * 2; }
    off = (off + 1023) & ~1023u;      // swizzle atoms are 1024-byte aligned
    for (uint32_t ci = 0; ci < kChains; ci++)
        for (uint32_t l = 0; l < kL; l++) { a.c[ci].in_off[l] = off; off += kTile * a.c[ci].dims[l] * 2; }
    const uint32_t tiles_end = off;
    for (uint32_t ci = 0; ci < kChains; ci++) { a.c[ci].dz_off = off; off += 2 * a.c[ci].dz_bytes; }
    a.enc_off = off; a.enc_stage_bytes = kTile * grid_dims[0] * 2;
    off += kRing * a.enc_stage_bytes;
    a.din_off = off;
    off += kRing * 2 * kPanel;
    a.ctrl_off = off;
    a.plans_off = (kBLevels + L * (uint32_t)sizeof(LevelConst) + 15) & ~15u;
    const uint32_t smem_bytes = std::max(off + a.plans_off + kChains * kL * 2 * (uint32_t)sizeof(MmaPlan), tiles_end + 16 * kPanel);
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    a.M = M; a.m_dev = m_dev; a.grad_table = (__half*)grad_table;
    a.density_act = density_act; a.color_act = color_act; a.beta = beta;
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
