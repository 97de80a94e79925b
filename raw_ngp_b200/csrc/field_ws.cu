// field_ws.cu -- the whole NeRF field forward (nerf/network.py:74-143) as ONE warp-specialised persistent kernel:
//
//     xyz --gather warps--> hash-grid features --> [A0 ring in shared memory] --MLP warps--> grid_mlp -> sigma, feat
//                                                                                   SH(dir) -> view_mlp -> colour -> rgb
//
// One CTA per SM, 28 warps:
//   * warps 0-15 (512 threads) only gather: thread (row, g) encodes levels g, g+4, g+8, g+12 of its sample with two levels
//     (16 table rows) in flight, writes the fp16 features into stage s of a 3-deep ring of 128 x 2L tiles and arrives on
//     full[s].  The gathers are bound by the SM's L1-miss path into the L2-resident table, so these warps never wait for
//     anything else: the ring decouples them from the tensor-core chain.
//   * warps 16-27 are three MLP groups (4 warps each) that take tiles in turn.  A group waits for full[s], then runs the six
//     layers of grid_mlp and view_mlp as tcgen05.mma (M = 128, accumulators in the group's 128 TMEM columns), with the
//     fp16 activations going TMEM -> registers -> shared memory between layers; tcgen05.commit on empty[s] hands the ring
//     stage back to the gather warps as soon as the first layer has consumed it.
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

constexpr uint32_t kGatherThreads = 512;
constexpr uint32_t kGatherGroups = kGatherThreads / kTile;
constexpr uint32_t kMlpGroups = 3;
constexpr uint32_t kWsThreads = kGatherThreads + kMlpGroups * kTile;     // 768
constexpr uint32_t kStages = kMlpGroups;      // one ring stage per MLP group: a stage barrier is only ever waited on by its own group (phase parity!)
constexpr uint32_t kGroupTmemCols = 128;
constexpr uint32_t kWsTmemCols = 512;
constexpr uint32_t kWsLayers = 6;

// control block (byte offsets from ctrl_off)
constexpr uint32_t kWsFull = 0, kWsEmpty = 8 * kStages, kWsDone = 16 * kStages, kWsTmemSlot = kWsDone + 8 * kMlpGroups;
constexpr uint32_t kWsLevels = (kWsTmemSlot + 4 + 15) & ~15u;
constexpr uint32_t kWsPlans = kWsLevels + kMaxLevels * sizeof(LevelConst);
constexpr uint32_t kWsCtrlBytes = kWsPlans + kMlpGroups * kWsLayers * sizeof(MmaPlan);

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
    uint32_t w_off[kWsLayers], a0_off, a0_stage_bytes, h_off[kMlpGroups], ctrl_off;
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
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t* ctrl = smem + a.ctrl_off;
    const uint32_t full_s = tc::smem_u32(ctrl + kWsFull), empty_s = tc::smem_u32(ctrl + kWsEmpty), done_s = tc::smem_u32(ctrl + kWsDone);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + kWsTmemSlot);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(ctrl + kWsLevels);
    MmaPlan* plans = reinterpret_cast<MmaPlan*>(ctrl + kWsPlans);     // [group][layer]

    // ---- prologue: TMEM, barriers, weights, per-level constants, MMA descriptors -----------------------------------
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kWsTmemCols);
    if (threadIdx.x == 32) {
        for (uint32_t s = 0; s < kStages; s++) {
            tc::mbar_init(full_s + 8 * s, kGatherThreads);
            tc::mbar_init(empty_s + 8 * s, 1);
        }
        for (uint32_t gI = 0; gI < kMlpGroups; gI++) tc::mbar_init(done_s + 8 * gI, 1);
    }
    for (uint32_t l = 0; l < a.n_run; l++) tsw::load_weight_tile_r(smem + a.w_off[l], a.w[l], a.N[l], a.K[l]);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + kMlpGroups * kWsLayers && (threadIdx.x - 64) % kWsLayers < a.n_run) {
        const uint32_t i = threadIdx.x - 64, gI = i / kWsLayers, l = i % kWsLayers;
        const uint32_t K = a.K[l], N = a.N[l];
        // Y = A [128 x K] (K-major) * W_l^T; A = ring stage 0 for layer 0 (the issuer adds the stage offset), else the group's tile
        const uint32_t a_saddr = tc::smem_u32(smem + (l == 0 ? a.a0_off : a.h_off[gI])), w_saddr = tc::smem_u32(smem + a.w_off[l]);
        MmaPlan& pl = plans[i];
        pl.idesc = tc::instr_desc(kTile, N, false, false);
        pl.n_steps = K / 16; pl.d_col = gI * kGroupTmemCols; pl.pad = 0;
        for (uint32_t ks = 0; ks < K / 16; ks++) {     // both operands K-major swizzled tiles of width K
            pl.step[ks].a = tsw::desc_kmajor_r(a_saddr, K, kTile, ks);
            pl.step[ks].b = tsw::desc_kmajor_r(w_saddr, K, N, ks);
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
            bool waited = false;
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
    } else {
        // ================================ MLP groups ================================
        const uint32_t gI = (warp - kGatherThreads / 32) / 4;             // group
        const uint32_t tg = threadIdx.x - kGatherThreads - gI * kTile;    // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + (((warp & 3u) * 32u) << 16) + gI * kGroupTmemCols;
        uint8_t* h = smem + a.h_off[gI];
        const MmaPlan* pl = plans + gI * kWsLayers;
        const uint32_t done = done_s + 8 * gI;
        uint32_t ph = 0;
        for (uint32_t it = gI;; it += kMlpGroups) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            if (tile >= n_tiles) break;
            const uint32_t s = it % kStages;
            const uint32_t row = tile * kTile + tg;
            const bool live = row < M;
            tc::mbar_wait(full_s + 8 * s, (it / kStages) & 1u);
            float out[16];
            for (uint32_t l = 0; l < a.n_run; l++) {
                if (tg == 0) {
                    tc::fence_after_sync();
                    const uint64_t stage_add = (l == 0) ? (uint64_t)((s * a.a0_stage_bytes) >> 4) : 0ull;
                    for (uint32_t ks = 0; ks < pl[l].n_steps; ks++)
                        tc::mma_f16_ss(tmem + pl[l].d_col, pl[l].step[ks].a + stage_add, pl[l].step[ks].b, pl[l].idesc, ks > 0);
                    // The A operand of this layer is also what the backward pass needs (enc, h1, h2 | in2, h1', h2'): its tile
                    // image goes to global memory as ONE bulk async copy while the MMA runs.  The commit is issued after the
                    // copy has finished reading shared memory, so "MMA done" also means "tile may be overwritten".
                    __half* save = (l == 0) ? a.enc_out : (l == 3) ? a.in2_out : a.acts[l - 1];
                    if (save && !a.rowmajor) {
                        const uint32_t bytes = kTile * a.K[l] * 2;
                        const uint32_t src = tc::smem_u32((l == 0) ? smem + a.a0_off + s * a.a0_stage_bytes : h);
                        tc::bulk_s2g(reinterpret_cast<uint8_t*>(save) + (size_t)tile * bytes, src, bytes);
                        tc::bulk_wait_read();
                    }
                    tc::mma_commit(done);
                    if (l == 0 && !a.rowmajor) tc::mma_commit(empty_s + 8 * s);      // ring stage free once layer 0 (and the copy) has read it
                }
                if (a.rowmajor) {
                    // the A operand as plain rows [M, K] (what the kernel-pair backward of field.cu / mlp.cu reads): every thread
                    // copies its own row out of the tile while the MMA runs
                    __half* save = (l == 0) ? a.enc_out : (l == 3) ? a.in2_out : a.acts[l - 1];
                    if (save && live) {
                        const uint32_t K = a.K[l];
                        const uint8_t* src = (l == 0) ? smem + a.a0_off + s * a.a0_stage_bytes : h;
                        uint4* dst = reinterpret_cast<uint4*>(save + (size_t)row * K);
                        for (uint32_t j = 0; j < K / 8; j++) dst[j] = *reinterpret_cast<const uint4*>(src + tsw::chunk_off_r(K, kTile, tg, j));
                    }
                    if (l == 0) {       // the ring stage goes back to the gather warps only after every row has been read
                        tc::named_bar_sync(1 + gI, kTile);
                        if (tg == 0) tc::mma_commit(empty_s + 8 * s);
                    }
                }
                tc::mbar_wait(done, ph);
                ph ^= 1;
                tc::fence_after_sync();
                const uint32_t N = a.N[l];
                if (l != 2 && l != 5) {
                    // hidden layer: ReLU, fp16, next layer's A operand (+ saved for the backward pass)
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
                        *reinterpret_cast<uint4*>(h + o0) = lo;
                        *reinterpret_cast<uint4*>(h + o1) = hi;
                    }
                } else if (l == 2) {
                    // grid_mlp output: sigma (network.py:112-115: fp16 linear output, activation in fp32) and the view_mlp
                    // input [feat(15), SH(dir)(16), (SH(light dir)(16)), 0]
                    tc::tmem_ld16(lane_addr, out);
                    if (live) {
                        const float o0 = half_round(out[0]);
                        float sg;
                        if (a.density_act == 0) sg = expf(o0);
                        else { const float bx = a.beta * o0; sg = (bx > 20.f) ? o0 : log1pf(expf(bx)) / a.beta; }
                        a.sigma_out[row] = sg;
                    }
                    if (a.n_run == 3) {          // density only: no view branch
                        tc::fence_before_sync();
                        tc::named_bar_sync(1 + gI, kTile);
                        continue;
                    }
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
#pragma unroll
                    for (uint32_t c = 0; c < (LDIR ? 6u : 4u); c++) *reinterpret_cast<uint4*>(h + tsw::chunk_off_r(a.K[3], kTile, tg, c)) = ch[c];
                } else {
                    // colour head (network.py:131-138): fp16 linear output, `color - 5` in fp16, exp in fp32
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
                if (l != 5) tc::fence_async_smem();
                tc::fence_before_sync();
                tc::named_bar_sync(1 + gI, kTile);
            }
        }
        if (tg == 0) tc::bulk_wait_all();
    }
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
    for (uint32_t gI = 0; gI < kMlpGroups; gI++) { a.h_off[gI] = off; off += kTile * hmax * 2; }
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
