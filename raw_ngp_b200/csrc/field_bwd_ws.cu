// field_bwd_ws.cu -- the whole NeRF field backward as ONE warp-specialised persistent kernel:
//
//   d rgb -> [V1: view_mlp layers 2,1] -dZ1'-> [V0: view layer 0] -d in2[:, :15]-> [G1: grid_mlp layers 2,1] -dZ1->
//   d sigma ---------------------------------------------------------------------^   [G0: grid layer 0] -d enc-> scatter warps
//
// One CTA per SM, 32 warps:
//   * warps 16-31 are four MLP groups (4 warps each) forming a pipeline over the 128-sample tiles; every group owns one or
//     two layers: dW_l^T += in_l^T dZ_l (accumulators resident in TMEM for the whole kernel) and dH = dZ_l W_l, ReLU-masked
//     into the next dZ.  Per layer dH is issued first (all the epilogue waits for), dW behind it.  Between groups dZ / d in2 /
//     d enc travel through shared-memory rings -- nothing of the backward chain touches HBM except the saved activations,
//     which arrive as BULK ASYNC COPIES (one per tensor and tile, mbarrier complete_tx) and are refilled for the next tile
//     as soon as their layer's MMAs have retired.
//     Splitting each MLP over two groups halves the serial tensor-core chain per tile, which is what bounds the kernel once
//     the scatter is fed continuously.
//   * warps 0-15 only scatter: thread (row, g) takes levels g, g+4, ... of its sample; runs of consecutive samples in the
//     same cell are merged by a segmented shuffle reduction in packed fp16x2 and only run heads issue reductions
//     (red.global.add.noftz.v2.f16x2 for an aligned x-neighbour pair).
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
constexpr uint32_t kRoles = 4;                                  // V1, V0, G1, G0
constexpr uint32_t kBwsThreads = kScatterThreads + kRoles * kTile;      // 1024
constexpr uint32_t kRing = 2;
constexpr uint32_t kBwsTmemCols = 512;
constexpr uint32_t kL = 3;

// control block (byte offsets from ctrl_off)
constexpr uint32_t kTFull = 0;                                   // [chain][layer]
constexpr uint32_t kEncFull = kTFull + 8 * kChains * kL, kEncEmpty = kEncFull + 8 * kRing;
constexpr uint32_t kDinFull = kEncEmpty + 8 * kRing, kDinEmpty = kDinFull + 8 * kRing;
constexpr uint32_t kDzFull = kDinEmpty + 8 * kRing, kDzEmpty = kDzFull + 8 * kChains;      // dZ1 hand-over inside a chain
constexpr uint32_t kDone = kDzEmpty + 8 * kChains, kTail = kDone + 8 * kRoles, kDxDone = kTail + 8 * kRoles, kSlot = kDxDone + 8;
constexpr uint32_t kDxRing = 3;                                  // per-tile d xyz accumulators (input gradients)
constexpr uint32_t kBLevels = (kSlot + 4 + 15) & ~15u;           // LevelConst[L], then the MMA plans

struct Chain {
    const __half* in[kL];          // tiled saved tensors: layer inputs (in2 | enc, h1, h2)
    const __half* w[kL];
    float* dw[kL];
    uint32_t dims[kL + 1];
    uint32_t w_off[kL], in_off[kL];
    uint32_t dz_off[kL];           // dZ tile read by layer l: [2] head buffer (16 wide), [1] private buffer, [0] hand-over buffer
    uint32_t acc_col[kL], work_col[kL];     // TMEM columns: dW accumulator / dH output of layer l
};

struct BwsArgs {
    Chain c[kChains];
    const float* xyzs; const float* d_sigma; const float* sigma; const float* d_rgb; const float* rgb;
    const float* dirs; float* d_xyzs; float* d_dirs;       // input gradients (IG): march dirs in, d xyzs / d dirs [M,3] out
    const __half* dydx;                                    // d enc / d x saved by the forward (layout: field_ws.cu)
    GridArgs g;
    uint32_t M; const int* m_dev;
    __half* grad_table;
    int density_act, color_act; float beta;
    uint32_t enc_off, enc_stage_bytes, din_off, dx_off, ctrl_off, plans_off;
};

// IG: also produce the gradients with respect to the sample positions and view directions (BARF pose refinement):
//   * the scatter warps contract the dy_dx their thread's levels got from the forward (12 coalesced 4-byte loads per
//     thread at L = 16) with d enc (kernel_input_backward, gridencoder.cu:352-378);
//     the four level groups of a sample meet in a ring of shared-memory accumulators (red.shared), which the G0 group
//     writes out three tiles later -- by then every scatter thread has moved past that tile (it has acknowledged the
//     d enc stage of the tile after it), so no extra barrier sits on the pipeline;
//   * the V0 group turns d in2[:, 15:31] (the SH features) into d dirs through the SH Jacobian and the two
//     normalisations of the forward (renderer.py:544, sphere_harmonics.py:81).
template <bool IG>
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
    const uint32_t dz_full = tc::smem_u32(ctrl + kDzFull), dz_empty = tc::smem_u32(ctrl + kDzEmpty);
    const uint32_t done = tc::smem_u32(ctrl + kDone), tail = tc::smem_u32(ctrl + kTail), dx_done = tc::smem_u32(ctrl + kDxDone);
    float* dx_ring = reinterpret_cast<float*>(smem + a.dx_off);          // [kDxRing][kTile][3]
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
        for (uint32_t ci = 0; ci < kChains; ci++) { tc::mbar_init(dz_full + 8 * ci, kTile); tc::mbar_init(dz_empty + 8 * ci, 1); }
        for (uint32_t r = 0; r < kRoles; r++) { tc::mbar_init(done + 8 * r, 1); tc::mbar_init(tail + 8 * r, 1); }
        tc::mbar_init(dx_done, kScatterThreads);
    }
    if (IG) for (uint32_t i = threadIdx.x; i < kDxRing * kTile * 3; i += blockDim.x) dx_ring[i] = 0.f;
    for (uint32_t ci = 0; ci < kChains; ci++)
        for (uint32_t l = 0; l < kL; l++) tsw::load_weight_tile(smem + a.c[ci].w_off[l], a.c[ci].w[l], a.c[ci].dims[l + 1], a.c[ci].dims[l]);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kChains * kL) {
        const uint32_t i = threadIdx.x - 64, ci = i / kL, l = i % kL;
        const Chain& c = a.c[ci];
        const uint32_t K = c.dims[l], N = c.dims[l + 1];
        const uint32_t dz_saddr = tc::smem_u32(smem + c.dz_off[l]);
        const uint32_t in_saddr = tc::smem_u32(smem + c.in_off[l]), w_saddr = tc::smem_u32(smem + c.w_off[l]);
        // all operands are swizzled row-major tiles (tile_sw.cuh): saved input [128 x K], dZ [128 x N], weights [N x K]
        MmaPlan& dw = plans[i * 2];       // dW_l^T [K x N] += in_l^T [K x 128] * dZ_l [128 x N]: both MN-major views, k = sample rows
        dw.idesc = tc::instr_desc(kTile, N, true, true);
        dw.n_steps = kTile / 16; dw.d_col = c.acc_col[l]; dw.pad = 0;
        for (uint32_t ks = 0; ks < kTile / 16; ks++) {
            // M = 128 > K: the MN blocks past the tile (stride = tile size) only feed TMEM lanes that are never read
            dw.step[ks].a = tsw::desc_mnmajor(in_saddr, K, ks, kTile * K * 2);
            dw.step[ks].b = tsw::desc_mnmajor(dz_saddr, N, ks, kTile * N * 2);
        }
        MmaPlan& dh = plans[i * 2 + 1];   // dH [128 x K] = dZ_l [128 x N] (K-major) * W_l [N x K] (MN-major view: MN = K, k = N rows)
        dh.idesc = tc::instr_desc(kTile, K, false, true);
        dh.n_steps = N / 16; dh.d_col = c.work_col[l]; dh.pad = 0;
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
            float dxa[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (uint32_t j = 0; j < kMaxLevels / kScatterGroups; j++) {
                const uint32_t level = grp + j * kScatterGroups;
                if (level < g.L) {
                    if (IG && live) {
                        // word (j % 2) * 3 + d of level pair j / 2, row-contiguous (layout: field_ws.cu)
                        const uint32_t* dy = reinterpret_cast<const uint32_t*>(a.dydx) +
                                             (((size_t)(tile * kScatterGroups + grp) * (g.L / 8) + j / 2) * 6 + (j % 2) * 3) * kTile + r;
                        float2 gf = __half22float2(gh[j]);
                        if (g.feat_weights) {      // the window multiplies the encoder output (network.py:99-109): chain rule
                            const __half2 gw = __floats2half2_rn(gf.x * __ldg(g.feat_weights + 2 * level), gf.y * __ldg(g.feat_weights + 2 * level + 1));
                            gf = __half22float2(gw);
                        }
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            const uint32_t u = __ldg(dy + d * kTile);
                            const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&u));
                            dxa[d] += gf.x * y.x + gf.y * y.y;
                        }
                    }
                    scatter_level(g, s_lv[level], level, x, live, gh[j], a.grad_table, lane);
                }
            }
            if (IG) {
                // the arrival on enc_empty of this thread's NEXT tile (a release) publishes these to the G0 group
                float* acc = dx_ring + ((it % kDxRing) * kTile + r) * 3;
                const float inv2b = __fdiv_rn(1.0f, 2.0f * g.bound);
#pragma unroll
                for (int d = 0; d < 3; d++) atomicAdd(acc + d, dxa[d] * inv2b);
            }
        }
        if (IG) tc::mbar_arrive(dx_done);
    } else {
        // ================================ MLP groups ================================
        const uint32_t role = (warp - kScatterThreads / 32) / 4;          // 0 = V1, 1 = V0, 2 = G1, 3 = G0
        const uint32_t ci = role / 2;                                     // chain: 0 = view_mlp, 1 = grid_mlp
        const bool upper = (role & 1u) == 0;                              // layers {2, 1} (else layer {0})
        const bool view = ci == 0;
        const Chain& c = a.c[ci];
        const uint32_t tg = threadIdx.x - kScatterThreads - role * kTile; // row inside the tile == TMEM lane
        const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);
        const uint32_t done_r = done + 8 * role, tail_r = tail + 8 * role, tf = t_full + 8 * ci * kL;
        const uint32_t dzf = dz_full + 8 * ci, dze = dz_empty + 8 * ci;
        const MmaPlan* pl = plans + ci * kL * 2;
        auto load_tensor = [&](uint32_t tile, uint32_t l) {               // one thread: bulk async copy of one saved tile
            const uint32_t bytes = kTile * c.dims[l] * 2;
#ifdef NGP_BWS_SKIP_HIDDEN_LOADS      // timing experiment: the hidden tiles are not fetched (what recomputing them would save)
            if (l > 0) { tc::mbar_arrive(tf + 8 * l); return; }
#endif
            tc::mbar_arrive_expect_tx(tf + 8 * l, bytes);
            tc::bulk_g2s(tc::smem_u32(smem + c.in_off[l]), c.in[l] + (size_t)tile * (c.dims[l] * kTile), bytes, tf + 8 * l);
        };
        // dH first (all the epilogue waits for), dW behind it: it runs on the tensor core during the epilogue
        auto issue_layer = [&](uint32_t l, bool accumulate, bool last_of_tile) {
            tc::fence_after_sync();
            issue_plan(tmem, pl[2 * l + 1], false);
            tc::mma_commit(done_r);
            issue_plan(tmem, pl[2 * l], accumulate);
            if (last_of_tile) tc::mma_commit(tail_r);
        };
        // ReLU-masked dH of layer l -> dZ tile `dst` of layer l - 1 (width K = dims[l])
        auto masked_epilogue = [&](uint32_t l, uint8_t* dst) {
            const uint32_t K = c.dims[l];
            const uint8_t* in_tile = smem + c.in_off[l];
            const uint32_t lane_addr = lane_base + c.work_col[l];
            for (uint32_t c0 = 0; c0 < K; c0 += 16) {
                float v[16];
                tc::tmem_ld16(lane_addr + c0, v);
                const uint32_t o0 = tsw::chunk_off(K, tg, c0 / 8), o1 = tsw::chunk_off(K, tg, c0 / 8 + 1);
                const uint4 m0 = *reinterpret_cast<const uint4*>(in_tile + o0);
                const uint4 m1 = *reinterpret_cast<const uint4*>(in_tile + o1);
                uint4 lo, hi;
                pack16(v, lo, hi);
                // ReLU mask on packed halves: the saved activation is a ReLU output (>= +0), so "was active" == "> 0";
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
                *reinterpret_cast<uint4*>(dst + o0) = lo;
                *reinterpret_cast<uint4*>(dst + o1) = hi;
            }
        };
        uint32_t ph = 0, it = 0;
        if (upper) {
            // ---------------- layers 2 and 1 of the chain ----------------
            if (tg == 0 && blockIdx.x < n_tiles) { load_tensor(blockIdx.x, 2); load_tensor(blockIdx.x, 1); }
            // per-sample head inputs, fetched one tile ahead: (d rgb, rgb) for the view chain, (d sigma, sigma) for the grid chain
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
            uint8_t* dz_head = smem + c.dz_off[2];
            for (;; it++) {
                const uint32_t tile = blockIdx.x + it * gridDim.x;
                if (tile >= n_tiles) break;
                const uint32_t e = it % kRing, rp = (it / kRing) & 1u;
                const bool live = tile * kTile + tg < M;
                const uint32_t next_tile = blockIdx.x + (it + 1) * gridDim.x;
                // ---- dZ of the last layer (16 columns) ----
                uint4 z0 = make_uint4(0, 0, 0, 0), z1 = z0;
                if (view) {
                    if (live) {
                        // d out = d rgb * d act / d out from the activated colour (exp: rgb; clamped exp: rgb below the clamp;
                        // sigmoid: rgb (1 - rgb)); columns 3.. of the padded output carry no gradient
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
                    // [d sigma * d act / d out0, d feat(15)]: the feature gradients come from the view chain through the ring
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
                *reinterpret_cast<uint4*>(dz_head + tsw::chunk_off(16, tg, 0)) = z0;
                *reinterpret_cast<uint4*>(dz_head + tsw::chunk_off(16, tg, 1)) = z1;
                fetch_head(next_tile);      // next tile's head inputs: in flight during this tile's chain
                tc::fence_async_smem();
                tc::fence_before_sync();
                tc::named_bar_sync(1 + role, kTile);
                // ---- layer 2 ----
                tc::mbar_wait(tf + 8 * 2, it & 1u);
                if (tg == 0) issue_layer(2, it > 0, false);
                tc::mbar_wait(done_r, ph); ph ^= 1;
                tc::fence_after_sync();
                masked_epilogue(2, smem + c.dz_off[1]);
                tc::fence_async_smem();
                tc::fence_before_sync();
                tc::named_bar_sync(1 + role, kTile);
                // ---- layer 1: its dZ output is handed to the group that owns layer 0 ----
                tc::mbar_wait(tf + 8 * 1, it & 1u);
                if (tg == 0) issue_layer(1, it > 0, true);
                tc::mbar_wait(done_r, ph); ph ^= 1;
                tc::fence_after_sync();
                if (tg == 0 && next_tile < n_tiles) load_tensor(next_tile, 2);      // dW of layer 2 has retired with this commit
                tc::mbar_wait(dze, (it & 1u) ^ 1u);                                  // the hand-over buffer is free again
                masked_epilogue(1, smem + c.dz_off[0]);
                tc::fence_async_smem();
                tc::mbar_arrive(dzf);
                // dW of layer 1 still reads its saved tile and the private dZ buffer: wait for it before the next tile's layer 2
                // epilogue rewrites that buffer
                tc::mbar_wait(tail_r, it & 1u);
                tc::fence_before_sync();
                tc::named_bar_sync(1 + role, kTile);
                // every thread has finished reading the layer-1 saved tile (ReLU mask) and its dW has retired: refill it
                if (tg == 0 && next_tile < n_tiles) load_tensor(next_tile, 1);
            }
        } else {
            // ---------------- layer 0 of the chain ----------------
            if (tg == 0 && blockIdx.x < n_tiles) load_tensor(blockIdx.x, 0);
            const uint32_t lane_addr = lane_base + c.work_col[0];
            for (;; it++) {
                const uint32_t tile = blockIdx.x + it * gridDim.x;
                if (tile >= n_tiles) break;
                const uint32_t e = it % kRing, rp = (it / kRing) & 1u;
                const uint32_t next_tile = blockIdx.x + (it + 1) * gridDim.x;
                tc::mbar_wait(dzf, it & 1u);                    // dZ1 of this tile has been written by the upper group
                tc::mbar_wait(tf + 8 * 0, it & 1u);
                if (tg == 0) issue_layer(0, it > 0, true);
                tc::mbar_wait(done_r, ph); ph ^= 1;
                tc::fence_after_sync();
                if (view) {
                    // d in2[:, :16] -> the grid chain (column 15 is an SH input: ignored there)
                    tc::mbar_wait(din_empty + 8 * e, rp ^ 1u);
                    float v[16];
                    tc::tmem_ld16(lane_addr, v);
                    uint4 lo, hi;
                    pack16(v, lo, hi);
                    uint8_t* di = smem + a.din_off + e * (2 * kPanel);
                    *reinterpret_cast<uint4*>(di + tg * 16) = lo;
                    *reinterpret_cast<uint4*>(di + kPanel + tg * 16) = hi;
                    tc::mbar_arrive(din_full + 8 * e);
                    if (IG) {
                        // d SH(dir) = d in2[:, 15:31] (fp16 like the autocast tensor) -> d dirs
                        float g2[16], gs[16];
                        tc::tmem_ld16(lane_addr + 16, g2);
                        gs[0] = half_round(v[15]);
#pragma unroll
                        for (int k = 1; k < 16; k++) gs[k] = half_round(g2[k - 1]);
                        const uint32_t row = tile * kTile + tg;
                        if (row < M) {
                            const float d0 = __ldg(a.dirs + (size_t)row * 3), d1 = __ldg(a.dirs + (size_t)row * 3 + 1), d2 = __ldg(a.dirs + (size_t)row * 3 + 2);
                            const float inv0 = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);            // renderer.py:544
                            const float u0 = d0 * inv0, u1 = d1 * inv0, u2 = d2 * inv0;
                            const float inv1 = 1.0f / sqrtf(u0 * u0 + u1 * u1 + u2 * u2);            // sphere_harmonics.py:81
                            const float x = u0 * inv1, y = u1 * inv1, z = u2 * inv1, zz = z * z;
                            float gx = 0.f, gy = 0.f, gz = 0.f;
                            constexpr int DEG = 4;
#define SH_TERM(i, val, ddx, ddy, ddz) gx += gs[i] * (ddx); gy += gs[i] * (ddy); gz += gs[i] * (ddz);
#include "sh_basis.inc"
#undef SH_TERM
                            // v = n / |n|  =>  d n = (d v - v (v . d v)) / |n|, twice
                            float t = x * gx + y * gy + z * gz;
                            gx = (gx - x * t) * inv1; gy = (gy - y * t) * inv1; gz = (gz - z * t) * inv1;
                            t = u0 * gx + u1 * gy + u2 * gz;
                            a.d_dirs[(size_t)row * 3] = (gx - u0 * t) * inv0;
                            a.d_dirs[(size_t)row * 3 + 1] = (gy - u1 * t) * inv0;
                            a.d_dirs[(size_t)row * 3 + 2] = (gz - u2 * t) * inv0;
                        }
                    }
                } else {
                    // d enc -> fp16 tile for the scatter warps
                    tc::mbar_wait(enc_empty + 8 * e, rp ^ 1u);
                    if (IG && it >= kDxRing) {
                        // every scatter thread has acknowledged tile it - 2, hence finished tile it - 3: write its d xyz out and
                        // clear the accumulators for tile `it` (same slot; published by the arrival on enc_full below)
                        float* acc = dx_ring + ((it % kDxRing) * kTile + tg) * 3;
                        const uint32_t row = (blockIdx.x + (it - kDxRing) * gridDim.x) * kTile + tg;
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            if (row < M) a.d_xyzs[(size_t)row * 3 + d] = acc[d];
                            acc[d] = 0.f;
                        }
                    }
                    uint8_t* de = smem + a.enc_off + e * a.enc_stage_bytes;
                    for (uint32_t c0 = 0; c0 < c.dims[0]; c0 += 16) {
                        float v[16];
                        tc::tmem_ld16(lane_addr + c0, v);
                        uint4 lo, hi;
                        pack16(v, lo, hi);
                        *reinterpret_cast<uint4*>(de + (c0 / 8) * kPanel + tg * 16) = lo;
                        *reinterpret_cast<uint4*>(de + (c0 / 8 + 1) * kPanel + tg * 16) = hi;
                    }
                    tc::mbar_arrive(enc_full + 8 * e);
                }
                // dW of layer 0 still reads the hand-over buffer and the layer-0 saved tile
                tc::mbar_wait(tail_r, it & 1u);
                if (tg == 0) {
                    tc::mbar_arrive(dze);
                    if (next_tile < n_tiles) load_tensor(next_tile, 0);
                }
                tc::fence_before_sync();
                tc::named_bar_sync(1 + role, kTile);        // the TMEM work columns are rewritten by the next tile
            }
        }
        if (IG && !upper && !view) {
            // the last tiles' d xyz: wait until every scatter thread has left its loop
            tc::mbar_wait(dx_done, 0);
            for (uint32_t j = (it > kDxRing ? it - kDxRing : 0u); j < it; j++) {
                const float* acc = dx_ring + ((j % kDxRing) * kTile + tg) * 3;
                const uint32_t row = (blockIdx.x + j * gridDim.x) * kTile + tg;
                if (row < M) {
#pragma unroll
                    for (int d = 0; d < 3; d++) a.d_xyzs[(size_t)row * 3 + d] = acc[d];
                }
            }
        }
        // reduce this group's weight-gradient accumulators (TMEM lane i = input feature i) into global memory
        if (it > 0) {
            tc::fence_after_sync();
            for (uint32_t l = upper ? 1u : 0u; l < (upper ? 3u : 1u); l++) {
                const uint32_t K = c.dims[l], N = c.dims[l + 1];
                for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                    float v[16];
                    tc::tmem_ld16(lane_base + c.acc_col[l] + c0, v);   // warp-collective: every lane participates
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
                                       float* const* grid_dweights, float* const* view_dweights, const void* dydx,
                                       const float* dirs, float* d_xyzs, float* d_dirs, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    const bool ig = d_xyzs != nullptr || d_dirs != nullptr;
    if (ig && (!d_xyzs || !d_dirs || !dydx || !dirs)) return NGP_ERR_NULL;
    if (ig && !aligned(dydx, 16)) return NGP_ERR_ALIGN;
    if (ig && view_dims[0] != 32) return NGP_ERR_UNSUPPORTED;      // d dirs reads the 16 SH columns of a 32-wide view input
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
    a.dirs = dirs; a.d_xyzs = d_xyzs; a.d_dirs = d_dirs; a.dydx = (const __half*)dydx;
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
        (void)max_k; (void)max_n;
    }
    // TMEM columns: per group [dH work | dW accumulators]; V1 (layers 2,1), V0 (layer 0), G1, G0
    {
        uint32_t col = 0;
        for (uint32_t ci = 0; ci < kChains; ci++) {
            Chain& c = a.c[ci];
            const uint32_t work_hi = std::max(c.dims[2], c.dims[1]);
            c.work_col[2] = c.work_col[1] = col; col += work_hi;
            c.acc_col[2] = col; col += c.dims[3];
            c.acc_col[1] = col; col += c.dims[2];
            c.work_col[0] = col; col += c.dims[0];
            c.acc_col[0] = col; col += c.dims[1];
        }
        if (col > kBwsTmemCols) return NGP_ERR_UNSUPPORTED;
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
    // dZ tiles per chain: head buffer (16 wide) | private buffer (layer 1, width dims[2]) | hand-over buffer (layer 0, width dims[1])
    for (uint32_t ci = 0; ci < kChains; ci++) {
        Chain& c = a.c[ci];
        c.dz_off[1] = off; off += kTile * c.dims[2] * 2;
        c.dz_off[0] = off; off += kTile * c.dims[1] * 2;
        c.dz_off[2] = off; off += kTile * c.dims[3] * 2;
    }
    off = (off + 1023) & ~1023u;
    a.enc_off = off; a.enc_stage_bytes = kTile * grid_dims[0] * 2;
    off += kRing * a.enc_stage_bytes;
    a.din_off = off;
    off += kRing * 2 * kPanel;
    a.dx_off = off;
    if (ig) off += kDxRing * kTile * 3 * (uint32_t)sizeof(float);
    a.ctrl_off = off;
    a.plans_off = (kBLevels + L * (uint32_t)sizeof(LevelConst) + 15) & ~15u;
    const uint32_t smem_bytes = std::max(off + a.plans_off + kChains * kL * 2 * (uint32_t)sizeof(MmaPlan), tiles_end + 16 * kPanel);
    if (a.c[0].dz_off[1] < tiles_end) return NGP_ERR_UNSUPPORTED;
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    a.M = M; a.m_dev = m_dev; a.grad_table = (__half*)grad_table;
    a.density_act = density_act; a.color_act = color_act; a.beta = beta;
    static thread_local SmemCache cache[2] = {};
    if (const int rc = ig ? ensure_dynamic_smem(field_backward_ws_kernel<true>, smem_bytes, cache[1])
                          : ensure_dynamic_smem(field_backward_ws_kernel<false>, smem_bytes, cache[0])) return rc;
    const uint32_t grid = std::min<uint32_t>(div_up(M, kTile), kNumSMs);
    if (ig) field_backward_ws_kernel<true><<<grid, kBwsThreads, smem_bytes, (cudaStream_t)stream>>>(a);
    else field_backward_ws_kernel<false><<<grid, kBwsThreads, smem_bytes, (cudaStream_t)stream>>>(a);
    return finish_launch();
}
