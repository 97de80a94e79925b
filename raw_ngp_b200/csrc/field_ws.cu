// field_ws.cu -- the whole NeRF field forward (nerf/network.py:74-143) as ONE warp-specialised persistent kernel:
//
//     xyz --gather warps--> hash-grid features --> [A0 ring in shared memory] --MLP warps--> grid_mlp -> sigma, feat
//                                                                                   SH(dir) -> view_mlp -> colour -> rgb
//
// One CTA per SM, 26 warps:
//   * warps 0-15 (512 threads) only gather: thread (row, g) encodes levels g, g+4, g+8, g+12 of its sample with several levels
//     (8 table rows each) in flight, writes the fp16 features into stage s of a ring of 128 x 2L tiles and arrives on
//     full[s].  The gathers are bound by the SM's L1 look-up rate on the L2-resident table, so these warps never wait for
//     anything else: the ring decouples them from the tensor-core chain.
//   * warps 16-23 are two MLP groups (4 warps each), warps 24-25 their MMA issuers (one thread each).  A group keeps TWO
//     tiles in flight ("slots", each with its own 128 TMEM columns, activation tile and barriers) and alternates between
//     them layer by layer: its 128 threads run the epilogue of layer l of one slot (TMEM -> registers -> ReLU / fp16 ->
//     shared memory) and signal ready[slot]; the issuer answers with the tcgen05.mma of layer l+1 (M = 128), the bulk copy
//     of the tile for the backward pass and a commit on done[slot] -- while the epilogue threads are already busy with the
//     other slot.  Neither the serial issue sequence nor the MMA latency sits on the epilogue threads' path any more (with
//     one tile per group and the issue inside the epilogue warps that was 40 % of the MLP warps' time, and the MLP chains,
//     not the gathers, set the pace of the kernel).
// All tiles are swizzled row-major (tile_sw.cuh: the UMMA SWIZZLE_32B/64B/128B canonical layouts), so the tensor core reads
// its operands without bank conflicts.  Saved activations (for the backward kernel) are written to global memory as the
// shared-memory image of the tile ("tile-panel" layout), so that the backward kernel fetches a whole tile with one bulk
// async copy.
//
// The reference runs this as ~60 PyTorch kernels per step (encoder, 6 nn.Linear, activations, SHEncoder, cat, casts).
#include "field_core.cuh"
#include "tile_sw.cuh"

namespace ngp {
namespace {

using namespace mlpcore;
using namespace gridcore;
using namespace fieldcore;

#ifndef NGP_WS_ISOLATE
#define NGP_WS_ISOLATE 0                // timing experiments only: 1 = the MLP side alone (no gathers), 2 = the gather warps alone,
                                        // bit 2 (4) = no saves, bit 3 (8) = no st.shared in the hidden epilogues
#endif
#ifndef NGP_WS_GATHER_LEVELS
#define NGP_WS_GATHER_LEVELS 2          // levels a gather thread keeps in flight (8 table rows each)
#endif
constexpr uint32_t kGatherThreads = 512;
constexpr uint32_t kGatherGroups = kGatherThreads / kTile;
constexpr uint32_t kMlpGroups = 2;
constexpr uint32_t kSlots = 2;                                           // tiles in flight per MLP group
constexpr uint32_t kWsThreads = kGatherThreads + kMlpGroups * kTile + kMlpGroups * 32;     // 832: gather | epilogue | issuers
// Ring stage <-> (group, slot): tile number `it` of the CTA goes to group it % kMlpGroups, which puts its tiles alternately
// into its two slots, so stage it % kStages is only ever consumed by one (group, slot) and in order (a waiter two phases
// early would read the parity of an older phase as "complete").
constexpr uint32_t kStages = kMlpGroups * kSlots;
constexpr uint32_t kSlotTmemCols = 128;
constexpr uint32_t kWsTmemCols = 512;
constexpr uint32_t kWsLayers = 6;
static_assert(kStages * kSlotTmemCols <= kWsTmemCols, "one accumulator per tile in flight");

// control block (byte offsets from ctrl_off)
constexpr uint32_t kWsFull = 0, kWsEmpty = 8 * kStages, kWsDone = 16 * kStages, kWsReady = 24 * kStages, kWsTmemSlot = 32 * kStages;
constexpr uint32_t kWsLevels = (kWsTmemSlot + 4 + 15) & ~15u;
constexpr uint32_t kWsPlans = kWsLevels + kMaxLevels * sizeof(LevelConst);
// One layer's tcgen05.mma sequence: k-step ks uses descriptors a0 + ks * a_step, b0 + ks * b_step (the start-address field
// advances by a constant per k-step in every layout of tile_sw.cuh), so the issuer reads 32 bytes per layer.
struct IssuePlan { uint64_t a0, b0; uint32_t a_step, b_step, idesc, n_steps; };
constexpr uint32_t kWsCtrlBytes = kWsPlans + kStages * kWsLayers * sizeof(IssuePlan);

struct WsArgs {
    const float* xyzs; const float* dirs; const float* ldirs;
    GridArgs g;
    const __half* w[kWsLayers];        // grid_mlp 0..2, view_mlp 0..2; [N, K] row-major fp16
    uint32_t K[kWsLayers], N[kWsLayers];
    __half* enc_out;                   // tile-panel (swizzled tile images, tile_sw.cuh) or nullptr
    __half* acts[kWsLayers];           // tiled hidden activations of layers 0,1 (grid) and 3,4 (view) or nullptr
    __half* in2_out;                   // tiled view_mlp input or nullptr
    __half* dydx_out;                  // d enc / d x, [tile][gather group][level pair][2 levels x 3 dims][row][2 channels] fp16, or nullptr
    float* sigma_out; float* rgb_out;
    uint32_t M; const int* m_dev;
    int density_act, color_act; float beta;
    uint32_t n_run;                    // layers to run: 6, or 3 for a density-only query (NeRFNetwork.density)
    uint32_t rowmajor;                 // saved tensors as plain [M, width] rows (for the kernel-pair backward) instead of tile images
    uint32_t w_off[kWsLayers], a0_off, a0_stage_bytes, h_off, h_bytes, ctrl_off;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) { return pack_h2(a, b); }

// DYDX: the gather warps also produce d enc / d x (calc_grad_inputs of the reference, gridencoder.cu:216-245) for the input
// gradients of the backward kernel, stored word-major / row-minor per gather group (coalesced 4-byte stores).
template <bool LDIR, bool DYDX>
__global__ void __launch_bounds__(kWsThreads, 1)
field_forward_ws_kernel(const WsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t M = a.M;
    if (a.m_dev) M = min(M, (uint32_t)__ldg(a.m_dev));   // sample count produced on the device (no host sync)
    const uint32_t warp = threadIdx.x >> 5;
    uint8_t* ctrl = smem + a.ctrl_off;
    const uint32_t full_s = tc::smem_u32(ctrl + kWsFull), empty_s = tc::smem_u32(ctrl + kWsEmpty), done_s = tc::smem_u32(ctrl + kWsDone);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kWsTmemSlot);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(ctrl + kWsLevels);
    const uint32_t ready_s = tc::smem_u32(ctrl + kWsReady);
    IssuePlan* plans = reinterpret_cast<IssuePlan*>(ctrl + kWsPlans);     // [stage = slot * kMlpGroups + group][layer]

    // ---- prologue: TMEM, barriers, weights, per-level constants, MMA descriptors -----------------------------------
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kWsTmemCols);
    if (threadIdx.x == 32) {
        for (uint32_t s = 0; s < kStages; s++) {
            tc::mbar_init(full_s + 8 * s, kGatherThreads);
            tc::mbar_init(empty_s + 8 * s, 1);
            tc::mbar_init(done_s + 8 * s, 1);
            tc::mbar_init(ready_s + 8 * s, kTile);
        }
    }
    for (uint32_t l = 0; l < a.n_run; l++) tsw::load_weight_tile_r(smem + a.w_off[l], a.w[l], a.N[l], a.K[l]);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kStages * kWsLayers && (threadIdx.x - 64) % kWsLayers < a.n_run) {
        const uint32_t i = threadIdx.x - 64, st = i / kWsLayers, l = i % kWsLayers;
        const uint32_t K = a.K[l], N = a.N[l];
        // Y = A [128 x K] (K-major) * W_l^T; A = the tile's ring stage for layer 0, else the slot's activation tile
        const uint32_t a_saddr = tc::smem_u32(smem + (l == 0 ? a.a0_off + st * a.a0_stage_bytes : a.h_off + st * a.h_bytes));
        const uint32_t w_saddr = tc::smem_u32(smem + a.w_off[l]);
        IssuePlan& pl = plans[i];
        pl.idesc = tc::instr_desc(kTile, N, false, false);
        pl.n_steps = K / 16;
        pl.a0 = tsw::desc_kmajor_r(a_saddr, K, kTile, 0);     // both operands K-major swizzled tiles of width K
        pl.b0 = tsw::desc_kmajor_r(w_saddr, K, N, 0);
        pl.a_step = (uint32_t)(tsw::desc_kmajor_r(a_saddr, K, kTile, 1) - pl.a0);
        pl.b_step = (uint32_t)(tsw::desc_kmajor_r(w_saddr, K, N, 1) - pl.b0);
    }
    load_level_consts(s_lv, a.g);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    const GridArgs& g = a.g;
    const uint32_t F = 2 * g.L;

    if (warp < kGatherThreads / 32) {
        // ================================ gather warps ================================
        const uint32_t r = threadIdx.x & (kTile - 1), grp = threadIdx.x / kTile;
        for (uint32_t it = 0;; it++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t s = it % kStages;
            const uint32_t row = tile * kTile + r;
            const bool live = row < M;
            float x[3] = {2.f, 2.f, 2.f};
            if (live) unit_cube(a.xyzs + (size_t)row * 3, g.bound, x);
            const bool inside = x[0] >= 0 && x[0] <= 1 && x[1] >= 0 && x[1] <= 1 && x[2] >= 0 && x[2] <= 1;
            const float xc[3] = {fminf(fmaxf(x[0], 0.f), 1.f), fminf(fmaxf(x[1], 0.f), 1.f), fminf(fmaxf(x[2], 0.f), 1.f)};
            uint8_t* a0 = smem + a.a0_off + s * a.a0_stage_bytes;
#if (NGP_WS_ISOLATE & 3) == 1
            tc::mbar_wait(empty_s + 8 * s, ((it / kStages) & 1u) ^ 1u);
            tc::fence_async_smem();
            tc::mbar_arrive(full_s + 8 * s);
            continue;
#endif
            bool waited = false;
#if NGP_WS_GATHER_LEVELS == 4
            // all four levels of the thread (32 table rows) in flight at once
            if (!DYDX && g.L == 4 * kGatherGroups && s_lv[grp].mode != 2 && s_lv[grp + kGatherGroups].mode != 2 &&
                s_lv[grp + 2 * kGatherGroups].mode != 2 && s_lv[grp + 3 * kGatherGroups].mode != 2) {
                LevelGather q[4];
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) gather_issue(q[j], g, s_lv[grp + j * kGatherGroups], xc);
                __half2 f[4];
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) f[j] = gather_finish(q[j], g, grp + j * kGatherGroups, inside);
                tc::mbar_wait(empty_s + 8 * s, ((it / kStages) & 1u) ^ 1u);
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) {
                    const uint32_t lv = grp + j * kGatherGroups;
                    *reinterpret_cast<__half2*>(a0 + tsw::chunk_off(F, r, lv / 4) + (lv % 4) * 4) = f[j];
                }
            } else
#endif
            for (uint32_t level = grp; level < g.L; level += 2 * kGatherGroups) {
                const uint32_t la = level, lb = level + kGatherGroups;       // L % 8 == 0
                __half2 f0, f1;
                uint32_t dy[6];
                if (s_lv[la].mode == 2 || s_lv[lb].mode == 2) {              // warp-uniform, rare
                    if (DYDX) {
                        dydx_level_generic(g.table, g.gridtype, g.align_corners, g.interp, s_lv[la].res, s_lv[la].hashmap_size, s_lv[la].offset,
                                           xc[0], xc[1], xc[2], inside, dy);
                        dydx_level_generic(g.table, g.gridtype, g.align_corners, g.interp, s_lv[lb].res, s_lv[lb].hashmap_size, s_lv[lb].offset,
                                           xc[0], xc[1], xc[2], inside, dy + 3);
                    }
                    const uint32_t ra = gather_level_generic(g.table, g.feat_weights, g.gridtype, g.align_corners, g.interp, s_lv[la].res,
                                                             s_lv[la].hashmap_size, s_lv[la].offset, xc[0], xc[1], xc[2], la, inside);
                    const uint32_t rb = gather_level_generic(g.table, g.feat_weights, g.gridtype, g.align_corners, g.interp, s_lv[lb].res,
                                                             s_lv[lb].hashmap_size, s_lv[lb].offset, xc[0], xc[1], xc[2], lb, inside);
                    f0 = *reinterpret_cast<const __half2*>(&ra); f1 = *reinterpret_cast<const __half2*>(&rb);
                } else {
                    LevelGather q0, q1;
                    gather_issue(q0, g, s_lv[la], xc);
                    gather_issue(q1, g, s_lv[lb], xc);
                    f0 = gather_finish(q0, g, la, inside); f1 = gather_finish(q1, g, lb, inside);
                    if (DYDX) {
                        float df[3];
                        uint32_t o3[3];
                        locate3_dfrac(xc, s_lv[la].res, g.align_corners, g.interp, df);
                        gather_finish_dydx(q0, df, (float)(g.align_corners ? s_lv[la].res - 1 : s_lv[la].res), inside, o3);
                        dy[0] = o3[0]; dy[1] = o3[1]; dy[2] = o3[2];
                        locate3_dfrac(xc, s_lv[lb].res, g.align_corners, g.interp, df);
                        gather_finish_dydx(q1, df, (float)(g.align_corners ? s_lv[lb].res - 1 : s_lv[lb].res), inside, o3);
                        dy[3] = o3[0]; dy[4] = o3[1]; dy[5] = o3[2];
                    }
                }
                if (DYDX) {
                    // levels la, lb are this thread's pair number p: six 4-byte words, each stored row-contiguous
                    // ([tile][group][pair][word][row]) so that a warp writes 128 contiguous bytes per word
                    uint32_t* dst = reinterpret_cast<uint32_t*>(a.dydx_out) +
                                    ((size_t)(tile * kGatherGroups + grp) * (g.L / 8) + (level - grp) / 8) * (6 * kTile) + r;
#pragma unroll
                    for (uint32_t k = 0; k < 6; k++) dst[k * kTile] = dy[k];
                }
                if (!waited) {      // the ring stage is needed only now: the first gathers of the tile overlap the wait
                    tc::mbar_wait(empty_s + 8 * s, ((it / kStages) & 1u) ^ 1u);
                    waited = true;
                }
                // features 2l, 2l+1 of row r: chunk l / 4 of the row, byte (l % 4) * 4 inside it
                const uint32_t oa = tsw::chunk_off(F, r, la / 4) + (la % 4) * 4, ob = tsw::chunk_off(F, r, lb / 4) + (lb % 4) * 4;
                *reinterpret_cast<__half2*>(a0 + oa) = f0;
                *reinterpret_cast<__half2*>(a0 + ob) = f1;
            }
            tc::fence_async_smem();
            tc::mbar_arrive(full_s + 8 * s);
        }
    } else if ((NGP_WS_ISOLATE & 3) == 2) {
        if (warp >= (kGatherThreads + kMlpGroups * kTile) / 32 && (threadIdx.x & 31u) == 0) {
            const uint32_t gI = warp - (kGatherThreads + kMlpGroups * kTile) / 32;
            for (uint32_t it = gI; blockIdx.x + it * gridDim.x < n_tiles; it += kMlpGroups) {
                tc::mbar_wait(full_s + 8 * (it % kStages), (it / kStages) & 1u);
                tc::mbar_arrive(empty_s + 8 * (it % kStages));
            }
        }
    } else if (warp < (kGatherThreads + kMlpGroups * kTile) / 32) {
        // ================================ MLP groups: epilogues ================================
        const uint32_t gI = (warp - kGatherThreads / 32) / 4;             // group
        const uint32_t tg = threadIdx.x - kGatherThreads - gI * kTile;    // row inside the tile == TMEM lane
        const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);
        uint32_t tile_of[kSlots] = {0u, 0u}, ph[kSlots] = {0u, 0u};
        bool active[kSlots];
        uint32_t n_next = 0;                                              // index of the group's next tile in its own sequence
        // the group's tiles go alternately into its two slots: tile number n of the group sits in slot n % kSlots
        auto next_tile = [&](uint32_t slot) {
            const uint32_t tile = blockIdx.x + (gI + n_next * kMlpGroups) * gridDim.x;
            active[slot] = tile < n_tiles;
            if (active[slot]) { tile_of[slot] = tile; n_next++; }
        };
        next_tile(0);
        next_tile(1);
        while (active[0] || active[1]) {
            for (uint32_t l = 0; l < a.n_run; l++) {
#pragma unroll
                for (uint32_t slot = 0; slot < kSlots; slot++) {
                    if (!active[slot]) continue;
                    const uint32_t tile = tile_of[slot], st = slot * kMlpGroups + gI;
                    const uint32_t row = tile * kTile + tg;
                    const bool live = row < M;
                    const uint32_t lane_addr = lane_base + st * kSlotTmemCols;
                    uint8_t* h = smem + a.h_off + st * a.h_bytes;
                    tc::mbar_wait(done_s + 8 * st, ph[slot]);             // layer l of this slot has retired (and its operand tile is free)
                    ph[slot] ^= 1u;
                    tc::fence_after_sync();
                    if (a.rowmajor && l == 0 && a.enc_out && live) {
                        // plain-row saves (what the kernel-pair backward of field.cu / mlp.cu reads): the encoder output row by row
                        // out of the ring stage; the hidden rows below go out straight from the epilogue's registers
                        const uint8_t* src = smem + a.a0_off + st * a.a0_stage_bytes;
                        uint4* dst = reinterpret_cast<uint4*>(a.enc_out + (size_t)row * a.K[0]);
                        for (uint32_t j = 0; j < a.K[0] / 8; j++) dst[j] = *reinterpret_cast<const uint4*>(src + tsw::chunk_off_r(a.K[0], kTile, tg, j));
                    }
                    const uint32_t N = a.N[l];
                    if (l != 2 && l != 5) {
                        // hidden layer: ReLU, fp16, next layer's A operand (+ saved for the backward pass)
                        uint4* grow = (a.rowmajor && a.acts[l] && live) ? reinterpret_cast<uint4*>(a.acts[l] + (size_t)row * N) : nullptr;
                        for (uint32_t c0 = 0; c0 < N; c0 += 16) {
                            float v[16];
                            tc::tmem_ld16(lane_addr + c0, v);
                            uint4 lo, hi;
                            pack16(v, lo, hi);
                            {   // ReLU on the packed halves (8 HMNMX2 instead of 16 FMNMX; rounding is monotone, so the order is immaterial)
                                const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
                                __half2* ql = reinterpret_cast<__half2*>(&lo);
                                __half2* qh = reinterpret_cast<__half2*>(&hi);
#pragma unroll
                                for (int i = 0; i < 4; i++) { ql[i] = __hmax2(ql[i], zero2); qh[i] = __hmax2(qh[i], zero2); }
                            }
                            const uint32_t o0 = tsw::chunk_off_r(N, kTile, tg, c0 / 8), o1 = tsw::chunk_off_r(N, kTile, tg, c0 / 8 + 1);
                            if (!(NGP_WS_ISOLATE & 8) || lo.x == 0x12345678u) {
                                *reinterpret_cast<uint4*>(h + o0) = lo;
                                *reinterpret_cast<uint4*>(h + o1) = hi;
                            }
                            if (grow) { grow[c0 / 8] = lo; grow[c0 / 8 + 1] = hi; }
                        }
                    } else if (l == 2) {
                        // grid_mlp output: sigma (network.py:112-115: fp16 linear output, activation in fp32) and the view_mlp
                        // input [feat(15), SH(dir)(16), (SH(light dir)(16)), 0]
                        float out[16];
                        tc::tmem_ld16(lane_addr, out);
                        if (live) {
                            const float o0 = half_round(out[0]);
                            float sg;
                            if (a.density_act == 0) sg = expf(o0);
                            else { const float bx = a.beta * o0; sg = (bx > 20.f) ? o0 : log1pf(expf(bx)) / a.beta; }
                            a.sigma_out[row] = sg;
                        }
                        if (a.n_run > 3) {          // (a density-only query has no view branch)
                            float sh[LDIR ? 32 : 16];
                            {
                                float dx = 0.f, dy = 0.f, dz = 1.f;
                                if (live) { dx = __ldg(a.dirs + (size_t)row * 3); dy = __ldg(a.dirs + (size_t)row * 3 + 1); dz = __ldg(a.dirs + (size_t)row * 3 + 2); }
                                float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);          // renderer.py:544
                                dx *= inv; dy *= inv; dz *= inv;
                                inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);                // SHEncoder.forward, sphere_harmonics.py:81
                                const float x = dx * inv, y = dy * inv, z = dz * inv, zz = z * z;
                                constexpr int DEG = 4;
#define SH_TERM(i, v, ddx, ddy, ddz) sh[i] = v;
#include "sh_basis.inc"
#undef SH_TERM
                            }
                            if (LDIR) {
                                float dx = 0.f, dy = 0.f, dz = 1.f;
                                if (live) { dx = __ldg(a.ldirs + (size_t)row * 3); dy = __ldg(a.ldirs + (size_t)row * 3 + 1); dz = __ldg(a.ldirs + (size_t)row * 3 + 2); }
                                const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);    // SHEncoder.forward only (ldirs are not pre-normalised)
                                const float x = dx * inv, y = dy * inv, z = dz * inv, zz = z * z;
                                constexpr int DEG = 4;
#define SH_TERM(i, v, ddx, ddy, ddz) sh[16 + i] = v;
#include "sh_basis.inc"
#undef SH_TERM
                            }
                            // row = out[1..15], sh[0..15], (sh[16..31]), 0  -> 4 (6) chunks of 8 halves
                            uint4 ch[LDIR ? 6 : 4];
                            ch[0] = make_uint4(pack2(out[1], out[2]), pack2(out[3], out[4]), pack2(out[5], out[6]), pack2(out[7], out[8]));
                            ch[1] = make_uint4(pack2(out[9], out[10]), pack2(out[11], out[12]), pack2(out[13], out[14]), pack2(out[15], sh[0]));
                            ch[2] = make_uint4(pack2(sh[1], sh[2]), pack2(sh[3], sh[4]), pack2(sh[5], sh[6]), pack2(sh[7], sh[8]));
                            if (LDIR) {
                                ch[3] = make_uint4(pack2(sh[9], sh[10]), pack2(sh[11], sh[12]), pack2(sh[13], sh[14]), pack2(sh[15], sh[16]));
                                ch[4] = make_uint4(pack2(sh[17], sh[18]), pack2(sh[19], sh[20]), pack2(sh[21], sh[22]), pack2(sh[23], sh[24]));
                                ch[5] = make_uint4(pack2(sh[25], sh[26]), pack2(sh[27], sh[28]), pack2(sh[29], sh[30]), pack2(sh[31], 0.f));
                            } else {
                                ch[3] = make_uint4(pack2(sh[9], sh[10]), pack2(sh[11], sh[12]), pack2(sh[13], sh[14]), pack2(sh[15], 0.f));
                            }
                            uint4* grow = (a.rowmajor && a.in2_out && live) ? reinterpret_cast<uint4*>(a.in2_out + (size_t)row * a.K[3]) : nullptr;
#pragma unroll
                            for (uint32_t c = 0; c < (LDIR ? 6u : 4u); c++) {
                                *reinterpret_cast<uint4*>(h + tsw::chunk_off_r(a.K[3], kTile, tg, c)) = ch[c];
                                if (grow) grow[c] = ch[c];
                            }
                        }
                    } else {
                        // colour head (network.py:131-138): fp16 linear output, `color - 5` in fp16, exp in fp32
                        float out[16];
                        tc::tmem_ld16(lane_addr, out);
                        if (live) {
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const float o = half_round(out[c]);
                                float rc;
                                if (a.color_act == 2) rc = half_round(1.0f / (1.0f + expf(-o)));
                                else {
                                    rc = expf(half_round(o - 5.0f));
                                    if (a.color_act == 3) rc = fminf(rc, 5.0f);
                                }
                                a.rgb_out[(size_t)row * 3 + c] = rc;
                            }
                        }
                    }
                    // hand the slot to the issuer: the activation tile is written (visible to the tensor core's proxy) and the
                    // accumulator columns have been read
                    tc::fence_async_smem();
                    tc::fence_before_sync();
                    tc::mbar_arrive(ready_s + 8 * st);
                    if (l + 1 == a.n_run) next_tile(slot);
                }
            }
        }
    } else if ((threadIdx.x & 31u) == 0) {
        // ================================ MMA issuers (one thread per group) ================================
        const uint32_t gI = warp - (kGatherThreads + kMlpGroups * kTile) / 32;
        uint32_t tile_of[kSlots] = {0u, 0u}, rph[kSlots] = {0u, 0u};
        bool active[kSlots];
        uint32_t n_next = 0;
        auto issue = [&](const IssuePlan& pl, uint32_t d_col) {
            uint64_t ad = pl.a0, bd = pl.b0;
            const uint32_t n = pl.n_steps, as = pl.a_step, bs = pl.b_step, idesc = pl.idesc;
            for (uint32_t ks = 0; ks < n; ks++, ad += as, bd += bs) tc::mma_f16_ss(tmem + d_col, ad, bd, idesc, ks > 0);
        };
        // Take the group's next tile into `slot`: wait for the gather warps, issue layer 0 on the ring stage and the bulk copy
        // of the stage (the encoder output, saved for the backward pass).  The commits are issued after the copy has finished
        // reading shared memory, so "layer retired" also means "operand tile may be overwritten".
        auto start_tile = [&](uint32_t slot) {
            const uint32_t it = gI + n_next * kMlpGroups;
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            active[slot] = tile < n_tiles;
            if (!active[slot]) return;
            n_next++;
            tile_of[slot] = tile;
            const uint32_t st = it % kStages;                             // == slot * kMlpGroups + gI
            tc::mbar_wait(full_s + 8 * st, (it / kStages) & 1u);
            tc::fence_after_sync();
            issue(plans[st * kWsLayers], st * kSlotTmemCols);
            if (a.enc_out && !a.rowmajor && !(NGP_WS_ISOLATE & 4)) {
                const uint32_t bytes = kTile * a.K[0] * 2;
                tc::bulk_s2g(reinterpret_cast<uint8_t*>(a.enc_out) + (size_t)tile * bytes, tc::smem_u32(smem + a.a0_off + st * a.a0_stage_bytes), bytes);
                tc::bulk_wait_read();
            }
            tc::mma_commit(done_s + 8 * st);
            // the ring stage goes back to the gather warps once layer 0 (and the copy) has read it; with plain-row saves the
            // epilogue threads still copy their rows out of it (released below, after their layer-0 epilogue)
            if (!a.rowmajor) tc::mma_commit(empty_s + 8 * st);
        };
        start_tile(0);
        start_tile(1);
        while (active[0] || active[1]) {
            for (uint32_t l = 0; l < a.n_run; l++) {
#pragma unroll
                for (uint32_t slot = 0; slot < kSlots; slot++) {
                    if (!active[slot]) continue;
                    const uint32_t st = slot * kMlpGroups + gI;
                    tc::mbar_wait(ready_s + 8 * st, rph[slot]);           // the epilogue of layer l of this slot is finished
                    rph[slot] ^= 1u;
                    if (a.rowmajor && l == 0) tc::mbar_arrive(empty_s + 8 * st);
                    if (l + 1 < a.n_run) {
                        tc::fence_after_sync();
                        issue(plans[st * kWsLayers + l + 1], st * kSlotTmemCols);
                        // the A operand of layer l + 1 (the tile just written) is what the backward pass needs (h1, h2 | in2, h1',
                        // h2'): its image goes to global memory as ONE bulk async copy while the MMA runs
                        __half* save = (l + 1 == 3) ? a.in2_out : a.acts[l];
                        if (save && !a.rowmajor && !(NGP_WS_ISOLATE & 4)) {
                            const uint32_t bytes = kTile * a.K[l + 1] * 2;
                            tc::bulk_s2g(reinterpret_cast<uint8_t*>(save) + (size_t)tile_of[slot] * bytes, tc::smem_u32(smem + a.h_off + st * a.h_bytes), bytes);
                            tc::bulk_wait_read();
                        }
                        tc::mma_commit(done_s + 8 * st);
                    } else {
                        start_tile(slot);
                    }
                }
            }
        }
        tc::bulk_wait_all();
    }
    __syncwarp();                 // the issuer lanes rejoin their warps before the block-wide barrier
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kWsTmemCols);
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_field_forward_full(const float* xyzs, const float* dirs, const float* ldirs, const void* table,
                                      const int32_t* offsets, const float* feat_weights, float bound, float S, uint32_t H,
                                      uint32_t L, uint32_t gridtype, int align_corners, uint32_t interp,
                                      const void* const* grid_weights, const uint32_t* grid_dims,
                                      const void* const* view_weights, const uint32_t* view_dims, uint32_t M,
                                      const int32_t* m_dev, int density_act, float beta, int color_act, void* enc_out,
                                      void* const* grid_acts_out, void* in2_out, void* const* view_acts_out, float* sigma_out,
                                      float* rgb_out, void* dydx_out, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    const bool density_only = view_weights == nullptr;      /* NeRFNetwork.density: grid_mlp only, dirs / rgb_out unused */
    if (!xyzs || !table || !offsets || !grid_weights || !grid_dims || !sigma_out) return NGP_ERR_NULL;
    if (!density_only && (!dirs || !view_dims || !rgb_out)) return NGP_ERR_NULL;
    if (L == 0 || L > kMaxLevels || L % 8 != 0 || gridtype > 1 || interp > 1 || density_act < 0 || density_act > 1 ||
        color_act < 1 || color_act > 3)
        return NGP_ERR_BAD_ARG;
    const uint32_t in2_w = ldirs ? 48u : 32u;
    if (grid_dims[0] != 2 * L || grid_dims[3] != 16) return NGP_ERR_UNSUPPORTED;
    if (!density_only && (view_dims[0] != in2_w || view_dims[3] != 16)) return NGP_ERR_UNSUPPORTED;
    WsArgs a = {};
    a.n_run = density_only ? 3u : kWsLayers;
    a.xyzs = xyzs; a.dirs = dirs; a.ldirs = ldirs;
    a.g = {(const __half*)table, offsets, feat_weights, S, bound, H, L, gridtype, interp, align_corners != 0};
    uint32_t off = 0, hmax = in2_w;
    for (uint32_t l = 0; l < a.n_run; l++) {
        const uint32_t* d = l < 3 ? grid_dims : view_dims;
        const void* const* w = l < 3 ? grid_weights : view_weights;
        const uint32_t j = l % 3;
        // layer inputs are swizzled tiles of width 16 / 32 / 64, or panels of 16 columns for any other multiple of 16 (the
        // light-stage view_mlp: 48 -> 80 -> 80; tile_sw.cuh).  The tile images of the first kind are what the warp-specialised
        // backward reads; with other widths the saved tensors are written as plain rows for the kernel-pair backward.
        if (d[j] == 0 || d[j] % 16 || d[j] > 128 || d[j + 1] == 0 || d[j + 1] % 16 || d[j + 1] > 128) return NGP_ERR_UNSUPPORTED;
        if (d[j] != 16 && d[j] != 32 && d[j] != 64) a.rowmajor = 1;
        if (!w[j]) return NGP_ERR_NULL;
        if (!aligned(w[j], 16)) return NGP_ERR_ALIGN;
        a.w[l] = (const __half*)w[j];
        a.K[l] = d[j]; a.N[l] = d[j + 1];
        a.w_off[l] = off;
        off += d[j] * d[j + 1] * 2;
        if (j < 2) hmax = std::max(hmax, d[j + 1]);
        void* const* acts = l < 3 ? grid_acts_out : view_acts_out;
        a.acts[l] = (acts && j < 2) ? (__half*)acts[j] : nullptr;
        if (a.acts[l] && !aligned(a.acts[l], 16)) return NGP_ERR_ALIGN;
    }
    if (!aligned(table, 4) || (enc_out && !aligned(enc_out, 16)) || (in2_out && !aligned(in2_out, 16))) return NGP_ERR_ALIGN;
    if (dydx_out && !aligned(dydx_out, 16)) return NGP_ERR_ALIGN;
    a.enc_out = (__half*)enc_out; a.in2_out = (__half*)in2_out; a.dydx_out = (__half*)dydx_out;
    a.sigma_out = sigma_out; a.rgb_out = rgb_out;
    a.M = M; a.m_dev = m_dev;
    a.density_act = density_act; a.color_act = color_act; a.beta = beta;
    off = (off + 1023) & ~1023u;      // swizzle atoms are 1024-byte aligned
    a.a0_off = off; a.a0_stage_bytes = kTile * 2 * L * 2;
    off += kStages * a.a0_stage_bytes;
    a.h_off = off; a.h_bytes = kTile * hmax * 2;      // one activation tile per tile in flight
    off += kStages * a.h_bytes;
    a.ctrl_off = off;
    const uint32_t smem_bytes = off + kWsCtrlBytes;
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t grid = std::min<uint32_t>(div_up(M, kTile), kNumSMs);
#define NGP_LAUNCH_WS(LD, DY)                                                                                               \
    {                                                                                                                       \
        static thread_local SmemCache cache = {};                                                                           \
        if (const int rc = ensure_dynamic_smem(field_forward_ws_kernel<LD, DY>, smem_bytes, cache)) return rc;              \
        field_forward_ws_kernel<LD, DY><<<grid, kWsThreads, smem_bytes, st>>>(a);                                           \
    }
    if (dydx_out) { if (ldirs) NGP_LAUNCH_WS(true, true) else NGP_LAUNCH_WS(false, true) }
    else { if (ldirs) NGP_LAUNCH_WS(true, false) else NGP_LAUNCH_WS(false, false) }
#undef NGP_LAUNCH_WS
    return finish_launch();
}
