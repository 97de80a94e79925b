// field.cu -- the NeRF field of nerf/network.py:74-143 fused around the tensor-core MLP kernels.
//
//   forward 1 (density):  xyz -> hash-grid encode (gathers into the A tile, no HBM round trip) -> grid_mlp
//                         -> sigma = act(out[0]) ; in2 = [out[1:16], SH(dir), (SH(light dir)), 0]
//   forward 2 (colour):   in2 -> view_mlp -> colour activation -> rgb           (mlp.cu, head epilogue)
//   backward 2:           d rgb -> d out2 -> view_mlp backward -> d in2          (mlp.cu, head prologue)
//   backward 1:           [d sigma * sigma', d in2[:, :15]] -> grid_mlp backward -> d enc (TMEM)
//                         -> warp-aggregated packed reductions straight into the table gradient
//
// Every elementwise op the reference runs as its own PyTorch kernel between these stages (input scaling, direction
// normalisation, slicing, cat, casts, exp/clamp and all their autograd mirrors) happens in registers here.
// Rounding points follow the autocast pipeline of the reference (fp16 linear outputs, fp32 exp, fp16 features).
#include "field_core.cuh"

namespace ngp {
namespace {

using namespace mlpcore;
using namespace gridcore;
using namespace fieldcore;

// ---------------------------------------------------------------------------------------------------
// forward 1: encode -> grid_mlp -> sigma, in2
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kFwdTmemCols = 128;
constexpr uint32_t kCtrlLevels = 16;                                          // byte offset of the LevelConst table in the control block
constexpr uint32_t kCtrlPlans = kCtrlLevels + kMaxLevels * sizeof(LevelConst);  // byte offset of the MMA plans

// 512 threads per CTA: thread (row, grp) with row = (warp % 4) * 32 + lane (the TMEM lane quadrant a warp may read) and
// grp = warp / 4.  In the encode phase group g takes levels g, g+4, g+8, ... of its sample (coarse and fine levels
// interleaved, so the four groups finish together); hidden-layer epilogues give each group 16 of the 64 columns; only
// group 0 runs the last epilogue while the other groups already gather the next tile.
constexpr uint32_t kFieldThreads = 512;
constexpr uint32_t kGroups = kFieldThreads / kTile;


template <bool LDIR>
__global__ void __launch_bounds__(kFieldThreads, 2)
field_forward_density_kernel(const float* __restrict__ xyzs, const float* __restrict__ dirs, const float* __restrict__ ldirs,
                             GridArgs g, MlpArgs p, uint32_t M, __half* __restrict__ enc_out, float* __restrict__ sigma_out,
                             __half* __restrict__ in2, uint32_t ld2, int density_act, float beta, uint32_t a_tile_off,
                             uint32_t ctrl_off, const int* __restrict__ m_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));   // sample count produced on the device (no host sync)
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t t = (warp & 3u) * 32u + (threadIdx.x & 31u);   // sample row inside the tile == TMEM lane
    const uint32_t grp = warp >> 2;
    uint8_t* a_tile = smem + a_tile_off;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + ctrl_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ctrl_off + 8);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(smem + ctrl_off + kCtrlLevels);
    MmaPlan* plans = reinterpret_cast<MmaPlan*>(smem + ctrl_off + kCtrlPlans);

    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), kFwdTmemCols);
    if (threadIdx.x == 0) tc::mbar_init(tc::smem_u32(mbar), 1);
    {
        uint32_t o = 0;
        for (uint32_t l = 0; l < p.n_layers; l++) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            load_weight_tile(smem + o, p.w[l], N, K);
            if (threadIdx.x == l) {     // Y = A [128 x K] (K-major) * W_l^T: one tcgen05.mma per 16 columns of K
                MmaPlan& pl = plans[l];
                const uint32_t a_saddr = tc::smem_u32(a_tile), w_saddr = tc::smem_u32(smem + o);
                pl.idesc = tc::instr_desc(kTile, N, false, false);
                pl.n_steps = K / 16; pl.d_col = 0; pl.pad = 0;
                for (uint32_t ks = 0; ks < K / 16; ks++) {
                    pl.step[ks].a = tc::smem_desc(a_saddr + ks * 2 * kPanel, kPanel, 128);
                    pl.step[ks].b = tc::smem_desc(w_saddr + ks * 2 * (N * 16), N * 16, 128);
                }
            }
            o += K * N * 2;
        }
    }
    load_level_consts(s_lv, g);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t mbar_saddr = tc::smem_u32(mbar);
    const uint32_t F = p.dims[0];  // = 2 * L

    uint32_t phase = 0;
    const uint32_t n_tiles = (M + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kTile + t;
        const bool live = row < M;
        // ---- hash-grid encode: this thread's levels of its sample, two levels (16 gathers) in flight at a time ----
        float x[3] = {2.f, 2.f, 2.f};
        if (live) unit_cube(xyzs + (size_t)row * 3, g.bound, x);
        const bool inside = x[0] >= 0 && x[0] <= 1 && x[1] >= 0 && x[1] <= 1 && x[2] >= 0 && x[2] <= 1;
        const float xc[3] = {fminf(fmaxf(x[0], 0.f), 1.f), fminf(fmaxf(x[1], 0.f), 1.f), fminf(fmaxf(x[2], 0.f), 1.f)};
        for (uint32_t level = grp; level < g.L; level += 2 * kGroups) {
            __half2 f0, f1;
            if (s_lv[level].mode == 2 || s_lv[level + kGroups].mode == 2) {      // warp-uniform, rare
                const uint32_t la = level, lb = level + kGroups;
                const uint32_t ra = gather_level_generic(g.table, g.feat_weights, g.gridtype, g.align_corners, g.interp, s_lv[la].res,
                                                         s_lv[la].hashmap_size, s_lv[la].offset, xc[0], xc[1], xc[2], la, inside);
                const uint32_t rb = gather_level_generic(g.table, g.feat_weights, g.gridtype, g.align_corners, g.interp, s_lv[lb].res,
                                                         s_lv[lb].hashmap_size, s_lv[lb].offset, xc[0], xc[1], xc[2], lb, inside);
                f0 = *reinterpret_cast<const __half2*>(&ra); f1 = *reinterpret_cast<const __half2*>(&rb);
            } else {
                LevelGather q0, q1;
                gather_issue(q0, g, s_lv[level], xc);
                gather_issue(q1, g, s_lv[level + kGroups], xc);          // L % 8 == 0
                f0 = gather_finish(q0, g, level, inside); f1 = gather_finish(q1, g, level + kGroups, inside);
            }
            // features 2l, 2l+1 of row t: panel l / 4, byte (l % 4) * 4 of the row's 16-byte chunk
            *reinterpret_cast<__half2*>(a_tile + (level / 4) * kPanel + t * 16 + (level % 4) * 4) = f0;
            *reinterpret_cast<__half2*>(a_tile + ((level + kGroups) / 4) * kPanel + t * 16 + ((level + kGroups) % 4) * 4) = f1;
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::fence_after_sync();
            issue_plan(tmem, plans[0], false);
            tc::mma_commit(mbar_saddr);
        }
        if (enc_out && live) {      // saved for the backward pass: chunk grp, grp + 4, ... of the row, read back from the tile
            for (uint32_t c = grp; c < F / 8; c += kGroups)
                *reinterpret_cast<uint4*>(enc_out + (size_t)row * F + c * 8) = *reinterpret_cast<const uint4*>(a_tile + c * kPanel + t * 16);
        }

        float out[16];
        for (uint32_t l = 0; l < p.n_layers; l++) {
            const uint32_t N = p.dims[l + 1];
            tc::mbar_wait(mbar_saddr, phase);
            phase ^= 1;
            tc::fence_after_sync();
            if (l + 1 < p.n_layers) {
                // hidden width <= 64: one 16-column slice per group (groups past N only take part in the barrier)
                const uint32_t c0 = grp * 16;
                const bool mine = c0 < N;
                uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                if (mine) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
                    pack16(v, lo, hi);
                    *reinterpret_cast<uint4*>(a_tile + (c0 / 8) * kPanel + t * 16) = lo;
                    *reinterpret_cast<uint4*>(a_tile + (c0 / 8 + 1) * kPanel + t * 16) = hi;
                    tc::fence_async_smem();
                }
                tc::fence_before_sync();
                __syncthreads();
                if (threadIdx.x == 0) {
                    tc::fence_after_sync();
                    issue_plan(tmem, plans[l + 1], false);
                    tc::mma_commit(mbar_saddr);
                }
                if (mine && p.acts[l] && live) {    // the hidden activations stream out while the next layer's MMA runs
                    uint4* gp = reinterpret_cast<uint4*>(p.acts[l] + (size_t)row * N + c0);
                    gp[0] = lo; gp[1] = hi;
                }
            } else if (grp == 0) {
                tc::tmem_ld16(lane_addr, out);      // grid_mlp output: 16 columns (warp-uniform branch)
                tc::fence_before_sync();
            }
        }

        if (grp != 0) continue;                     // the last epilogue is small: group 0 only
        if (live) {
            // sigma (network.py:112-115): the linear output is fp16 under autocast, the activation runs in fp32
            const float o0 = half_round(out[0]);
            float sg;
            if (density_act == 0) sg = expf(o0);
            else { const float bx = beta * o0; sg = (bx > 20.f) ? o0 : log1pf(expf(bx)) / beta; }
            sigma_out[row] = sg;
        }
        if (live && in2) {
            // in2 = [feat(15), SH(dir)(16), (SH(light dir)(16)), 0]
            __align__(16) __half rowbuf[LDIR ? 48 : 32];
#pragma unroll
            for (int i = 0; i < 15; i++) rowbuf[i] = __float2half_rn(out[i + 1]);
            {
                float dx = __ldg(dirs + (size_t)row * 3), dy = __ldg(dirs + (size_t)row * 3 + 1), dz = __ldg(dirs + (size_t)row * 3 + 2);
                float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);          // renderer.py:544
                dx *= inv; dy *= inv; dz *= inv;
                inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);                // SHEncoder.forward, sphere_harmonics.py:81
                const float x = dx * inv, y = dy * inv, z = dz * inv, zz = z * z;
                constexpr int DEG = 4;
#define SH_TERM(i, v, ddx, ddy, ddz) rowbuf[15 + i] = __float2half_rn(v);
#include "sh_basis.inc"
#undef SH_TERM
            }
            if (LDIR) {
                float dx = __ldg(ldirs + (size_t)row * 3), dy = __ldg(ldirs + (size_t)row * 3 + 1), dz = __ldg(ldirs + (size_t)row * 3 + 2);
                const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);    // SHEncoder.forward only (ldirs are not pre-normalised)
                const float x = dx * inv, y = dy * inv, z = dz * inv, zz = z * z;
                constexpr int DEG = 4;
#define SH_TERM(i, v, ddx, ddy, ddz) rowbuf[31 + i] = __float2half_rn(v);
#include "sh_basis.inc"
#undef SH_TERM
                rowbuf[47] = __float2half_rn(0.f);
            } else {
                rowbuf[31] = __float2half_rn(0.f);
            }
            uint4* dst = reinterpret_cast<uint4*>(in2 + (size_t)row * ld2);
            const uint4* src = reinterpret_cast<const uint4*>(rowbuf);
#pragma unroll
            for (int i = 0; i < (LDIR ? 6 : 4); i++) dst[i] = src[i];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kFwdTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// backward 1: [d sigma, d in2[:, :15]] -> grid_mlp backward -> table gradient
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kBwdTmemCols = 256;
constexpr uint32_t kBwdThreads = 512;
constexpr uint32_t kBwdGroups = kBwdThreads / kTile;

// 512 threads per CTA, thread (row, grp) as in the forward kernel: row = (warp % 4) * 32 + lane is the sample (TMEM lane),
// grp = warp / 4 splits the column work: saved-activation loads, the 64 columns of the hidden-layer epilogues and -- the
// expensive part -- the 16 levels of the table-gradient scatter (4 levels per group).
__global__ void __launch_bounds__(kBwdThreads, 2)
field_backward_density_kernel(const float* __restrict__ xyzs, const float* __restrict__ d_sigma, const float* __restrict__ sigma,
                              const __half* __restrict__ d_in2, uint32_t ld2, const __half* __restrict__ enc, GridArgs g,
                              MlpArgs p, uint32_t M, __half* __restrict__ grad_table, int density_act, float beta,
                              uint32_t dz_off, uint32_t dz_bytes, uint32_t w_base, uint32_t ctrl_off,
                              const int* __restrict__ m_dev, const __half* __restrict__ dydx, float* __restrict__ d_xyzs) {
    extern __shared__ __align__(128) uint8_t smem[];
    // input gradients (dydx != nullptr): the four level groups of a sample add their parts of d loss / d xyz here
    __shared__ float s_dx[kTile * 3];
    if (threadIdx.x < kTile * 3) s_dx[threadIdx.x] = 0.f;
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t t = (warp & 3u) * 32u + lane;     // sample row inside the tile == TMEM lane
    const uint32_t grp = warp >> 2;
    const uint32_t L = p.n_layers;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + ctrl_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ctrl_off + 8);
    LevelConst* s_lv = reinterpret_cast<LevelConst*>(smem + ctrl_off + kCtrlLevels);
    MmaPlan* plans = reinterpret_cast<MmaPlan*>(smem + ctrl_off + kCtrlPlans);    // [2 l] = dW of layer l, [2 l + 1] = dH

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
    if (threadIdx.x < L) {
        const uint32_t l = threadIdx.x, K = p.dims[l], N = p.dims[l + 1];
        // dZ of layer l sits in the ping-pong buffer (L - 1 - l) & 1
        const uint32_t dz_saddr = tc::smem_u32(smem + dz_off + ((L - 1 - l) & 1u) * dz_bytes);
        const uint32_t in_saddr = tc::smem_u32(smem + in_off[l]), w_saddr = tc::smem_u32(smem + w_off[l]);
        MmaPlan& dw = plans[2 * l];      // dW_l^T [K x N] += in_l^T [K x 128] * dZ_l [128 x N]  (both MN-major views of row tiles)
        dw.idesc = tc::instr_desc(kTile, N, true, true);
        dw.n_steps = kTile / 16; dw.d_col = acc_col[l]; dw.pad = 0;
        for (uint32_t ks = 0; ks < kTile / 16; ks++) {
            dw.step[ks].a = tc::smem_desc(in_saddr + ks * 256, 128, kPanel);
            dw.step[ks].b = tc::smem_desc(dz_saddr + ks * 256, 128, kPanel);
        }
        MmaPlan& dh = plans[2 * l + 1];  // dH [128 x K] = dZ_l [128 x N] * W_l [N x K]  (A K-major, B = MN-major view of the weights)
        dh.idesc = tc::instr_desc(kTile, K, false, true);
        dh.n_steps = N / 16; dh.d_col = 0; dh.pad = 0;
        for (uint32_t ks = 0; ks < N / 16; ks++) {
            dh.step[ks].a = tc::smem_desc(dz_saddr + ks * 2 * kPanel, kPanel, 128);
            dh.step[ks].b = tc::smem_desc(w_saddr + ks * 256, 128, N * 16);
        }
    }
    load_level_consts(s_lv, g);
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
        const bool live = row < M;
        uint32_t cur = 0;
        // saved activations of the forward pass: 16-byte chunk c of the row goes to group c % 4
        for (uint32_t l = 0; l < L; l++) {
            const __half* src = (l == 0) ? enc : p.acts[l - 1];
            uint8_t* tile_s = smem + in_off[l];
            for (uint32_t c = grp; c < p.dims[l] / 8; c += kBwdGroups) {
                if (live) tc::cp_async16(tc::smem_u32(tile_s + c * kPanel + t * 16), src + (size_t)row * p.dims[l] + c * 8);
                else *reinterpret_cast<uint4*>(tile_s + c * kPanel + t * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        // d out1 = [d sigma * d act / d out0, d feat(15)]
        if (grp == 0) {
            __align__(16) __half dz[16];
            if (live) {
                const float sg = __ldg(sigma + row);
                float dact;
                if (density_act == 0) dact = sg;                              // trunc_exp backward: g * exp(x) (activation.py:18-21)
                else dact = 1.0f - expf(-beta * sg);                          // softplus' = sigmoid(beta x) = 1 - exp(-beta y)
                dz[0] = __float2half_rn(__ldg(d_sigma + row) * dact);
                const uint4* src = reinterpret_cast<const uint4*>(d_in2 + (size_t)row * ld2);
                const uint4 a = __ldg(src), b = __ldg(src + 1);
                const __half* ha = reinterpret_cast<const __half*>(&a);
                const __half* hb = reinterpret_cast<const __half*>(&b);
#pragma unroll
                for (int i = 0; i < 8; i++) dz[1 + i] = ha[i];
#pragma unroll
                for (int i = 0; i < 7; i++) dz[9 + i] = hb[i];
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) dz[i] = __float2half_rn(0.f);
            }
            uint8_t* dzt = smem + dz_off;
            *reinterpret_cast<uint4*>(dzt + t * 16) = reinterpret_cast<const uint4*>(dz)[0];
            *reinterpret_cast<uint4*>(dzt + kPanel + t * 16) = reinterpret_cast<const uint4*>(dz)[1];
        }
        tc::cp_async_wait_all();
        tc::fence_async_smem();
        __syncthreads();

        for (int l = (int)L - 1; l >= 0; l--) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            if (threadIdx.x == 0) {
                tc::fence_after_sync();
                issue_plan(tmem, plans[2 * l], iter > 0);
                issue_plan(tmem, plans[2 * l + 1], false);
                tc::mma_commit(mbar_saddr);
            }
            tc::mbar_wait(mbar_saddr, phase);
            phase ^= 1;
            tc::fence_after_sync();
            if (l > 0) {
                uint8_t* nxt = smem + dz_off + (cur ^ 1) * dz_bytes;
                const uint8_t* in_tile = smem + in_off[l];
                for (uint32_t c0 = grp * 16; c0 < K; c0 += kBwdGroups * 16) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + c0, v);
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
                }
                tc::fence_async_smem();
            } else {
                // ---- d enc of this thread's sample (TMEM lane) -> hash-table gradient; this group's 4 levels per pass ----
                float x[3] = {2.f, 2.f, 2.f};
                if (live) unit_cube(xyzs + (size_t)row * 3, g.bound, x);
                float dxa[3] = {0.f, 0.f, 0.f};
                for (uint32_t level = grp; level < g.L; level += kBwdGroups) {     // levels interleaved across the groups
                    float v[2];
                    tc::tmem_ld2(lane_addr + 2 * level, v);
                    const __half2 gh = __floats2half2_rn(v[0], v[1]);
                    if (dydx && live) {
                        // d loss / d xyz += d enc(level) . d enc / d x, with the dy_dx the warp-specialised forward saved for this
                        // thread's (row, level group): kernel_input_backward of the reference (gridencoder.cu:352-378); layout and
                        // arithmetic as in field_bwd_ws.cu
                        const uint32_t j = (level - grp) / kBwdGroups;
                        const uint32_t* dy = reinterpret_cast<const uint32_t*>(dydx) +
                                             (((size_t)(tile * kBwdGroups + grp) * (g.L / 8) + j / 2) * 6 + (j % 2) * 3) * kTile + t;
                        float2 gf = __half22float2(gh);
                        if (g.feat_weights) {
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
                    scatter_level(g, s_lv[level], level, x, live, gh, grad_table, lane);
                }
                if (dydx) {
                    const float inv2b = __fdiv_rn(1.0f, 2.0f * g.bound);
#pragma unroll
                    for (int d = 0; d < 3; d++) atomicAdd(&s_dx[t * 3 + d], dxa[d] * inv2b);
                }
            }
            tc::fence_before_sync();
            __syncthreads();
            cur ^= 1;
        }
        if (dydx && grp == 0) {      // all four groups have added their parts (barrier above); the next tile adds after another barrier
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (live) d_xyzs[(size_t)row * 3 + d] = s_dx[t * 3 + d];
                s_dx[t * 3 + d] = 0.f;
            }
        }
    }
    if (iter > 0) {
        tc::fence_after_sync();
        for (uint32_t l = 0; l < L; l++) {
            const uint32_t K = p.dims[l], N = p.dims[l + 1];
            for (uint32_t c0 = grp * 16; c0 < N; c0 += kBwdGroups * 16) {
                float v[16];
                tc::tmem_ld16(lane_addr + acc_col[l] + c0, v);
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

bool fill_args(MlpArgs& p, const void* const* weights, void* const* acts, float* const* dweights, const uint32_t* dims,
               uint32_t n_layers) {
    if (n_layers < 1 || n_layers > kMaxLayers) return false;
    for (uint32_t l = 0; l <= n_layers; l++)
        if (dims[l] == 0 || dims[l] % 16 != 0 || dims[l] > 128) return false;
    p.n_layers = n_layers;
    for (uint32_t l = 0; l < n_layers; l++) {
        if (!weights[l] || !aligned(weights[l], 16)) return false;
        p.w[l] = (const __half*)weights[l];
        p.acts[l] = (acts && l + 1 < n_layers) ? (__half*)acts[l] : nullptr;
        p.dw[l] = dweights ? dweights[l] : nullptr;
    }
    for (uint32_t l = 0; l <= n_layers; l++) p.dims[l] = dims[l];
    return true;
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_field_forward_density(const float* xyzs, const float* dirs, const float* ldirs, const void* table,
                                         const int32_t* offsets, const float* feat_weights, float bound, float S, uint32_t H,
                                         uint32_t L, uint32_t gridtype, int align_corners, uint32_t interp,
                                         const void* const* weights, const uint32_t* dims, uint32_t n_layers, uint32_t M,
                                         const int32_t* m_dev, int density_act, float beta, void* enc_out, void* const* acts_out, float* sigma_out,
                                         void* in2, uint32_t ld2, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!xyzs || !table || !offsets || !weights || !dims || !sigma_out) return NGP_ERR_NULL;
    if (in2 && !dirs) return NGP_ERR_NULL;   /* in2 == NULL: density only (NeRFNetwork.density) */
    if (L == 0 || L > kMaxLevels || L % 8 != 0 || gridtype > 1 || interp > 1 || density_act < 0 || density_act > 1) return NGP_ERR_BAD_ARG;
    MlpArgs p = {};
    if (!fill_args(p, weights, acts_out, nullptr, dims, n_layers)) return NGP_ERR_UNSUPPORTED;
    if (dims[0] != 2 * L || dims[n_layers] != 16) return NGP_ERR_UNSUPPORTED;
    for (uint32_t l = 1; l < n_layers; l++)
        if (dims[l] > kGroups * 16) return NGP_ERR_UNSUPPORTED;     /* hidden width <= 64 (network.py:49) */
    const uint32_t need2 = ldirs ? 48u : 32u;
    if (in2 && (ld2 < need2 || ld2 % 8)) return NGP_ERR_BAD_ARG;
    if (!aligned(table, 16) || (in2 && !aligned(in2, 16)) || (enc_out && !aligned(enc_out, 16))) return NGP_ERR_ALIGN;
    GridArgs g = {(const __half*)table, offsets, feat_weights, S, bound, H, L, gridtype, interp, align_corners != 0};
    uint32_t w_bytes = 0, max_k = 0;
    for (uint32_t l = 0; l < n_layers; l++) { w_bytes += dims[l] * dims[l + 1] * 2; max_k = std::max(max_k, dims[l]); }
    const uint32_t a_off = (w_bytes + 127) & ~127u;
    const uint32_t ctrl_off = a_off + kTile * max_k * 2;
    const uint32_t smem_bytes = ctrl_off + kCtrlPlans + kMaxLayers * sizeof(MmaPlan);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t grid = std::min<uint32_t>(div_up(M, kTile), kNumSMs * 2);
#define NGP_LAUNCH_FWD(LD)                                                                                                \
    {                                                                                                                     \
        static thread_local SmemCache cache = {};                                                                         \
        if (const int rc = ensure_dynamic_smem(field_forward_density_kernel<LD>, smem_bytes, cache)) return rc;           \
        field_forward_density_kernel<LD><<<grid, kFieldThreads, smem_bytes, st>>>(xyzs, dirs, ldirs, g, p, M, (__half*)enc_out,   \
                                                                          sigma_out, (__half*)in2, ld2, density_act, beta, \
                                                                          a_off, ctrl_off, m_dev);                        \
    }
    if (ldirs) NGP_LAUNCH_FWD(true) else NGP_LAUNCH_FWD(false)
#undef NGP_LAUNCH_FWD
    return finish_launch();
}

extern "C" int ngp_field_backward_density(const float* xyzs, const float* d_sigma, const float* sigma, const void* d_in2,
                                          uint32_t ld2, const void* enc, const void* dydx, const int32_t* offsets,
                                          const float* feat_weights, float bound, float S, uint32_t H, uint32_t L,
                                          uint32_t gridtype, int align_corners, uint32_t interp, const void* const* weights,
                                          const void* const* acts, const uint32_t* dims, uint32_t n_layers, uint32_t M,
                                          const int32_t* m_dev, int density_act, float beta, void* grad_table, float* const* dweights,
                                          float* d_xyzs, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if ((dydx != nullptr) != (d_xyzs != nullptr)) return NGP_ERR_NULL;
    if (dydx && (!aligned(dydx, 16) || L % 8 != 0)) return NGP_ERR_BAD_ARG;
    if (!xyzs || !d_sigma || !sigma || !d_in2 || !enc || !offsets || !weights || !dims || !grad_table || !dweights) return NGP_ERR_NULL;
    if (n_layers > 1 && !acts) return NGP_ERR_NULL;
    if (L == 0 || L > kMaxLevels || L % 4 != 0 || gridtype > 1 || interp > 1 || density_act < 0 || density_act > 1) return NGP_ERR_BAD_ARG;
    MlpArgs p = {};
    if (!fill_args(p, weights, (void* const*)acts, dweights, dims, n_layers)) return NGP_ERR_UNSUPPORTED;
    if (dims[0] != 2 * L || dims[n_layers] != 16 || ld2 < 16 || ld2 % 8) return NGP_ERR_UNSUPPORTED;
    if (!aligned(grad_table, 16) || !aligned(d_in2, 16) || !aligned(enc, 16)) return NGP_ERR_ALIGN;
    for (uint32_t l = 0; l < n_layers; l++) {
        if (!dweights[l]) return NGP_ERR_NULL;
        if (l + 1 < n_layers && !acts[l]) return NGP_ERR_NULL;
    }
    GridArgs g = {nullptr, offsets, feat_weights, S, bound, H, L, gridtype, interp, align_corners != 0};
    uint32_t w_bytes = 0, in_bytes = 0, max_n = 0, max_k = 0, acc_cols = 0;
    for (uint32_t l = 0; l < n_layers; l++) {
        w_bytes += dims[l] * dims[l + 1] * 2;
        in_bytes += kTile * dims[l] * 2;
        max_n = std::max(max_n, dims[l + 1]);
        max_k = std::max(max_k, dims[l]);
        acc_cols += dims[l + 1];
    }
    if (max_k + acc_cols > kBwdTmemCols) return NGP_ERR_UNSUPPORTED;
    const uint32_t dz_bytes = kTile * std::max(max_n, max_k) * 2;
    const uint32_t dz_off = in_bytes;
    const uint32_t w_base = dz_off + 2 * dz_bytes;
    const uint32_t ctrl_off = (w_base + w_bytes + 127) & ~127u;
    const uint32_t last_in_off = in_bytes - kTile * dims[n_layers - 1] * 2;
    const uint32_t smem_bytes = std::max<uint32_t>(ctrl_off + kCtrlPlans + 2 * kMaxLayers * sizeof(MmaPlan), last_in_off + 18 * kPanel);
    if (smem_bytes > 227 * 1024) return NGP_ERR_UNSUPPORTED;
    static thread_local SmemCache cache = {};
    if (const int rc = ensure_dynamic_smem(field_backward_density_kernel, smem_bytes, cache)) return rc;
    const uint32_t grid = std::min<uint32_t>(div_up(M, kTile), kNumSMs * 2);
    field_backward_density_kernel<<<grid, kBwdThreads, smem_bytes, (cudaStream_t)stream>>>(
        xyzs, d_sigma, sigma, (const __half*)d_in2, ld2, (const __half*)enc, g, p, M, (__half*)grad_table, density_act, beta,
        dz_off, dz_bytes, w_base, ctrl_off, m_dev, (const __half*)dydx, d_xyzs);
    return finish_launch();
}
