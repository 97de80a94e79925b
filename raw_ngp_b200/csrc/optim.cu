// optim.cu -- fused Adam over the flat parameter buffer (hash table + MLP weights): GradScaler unscale,
// inf/nan skip, Adam update on fp32 master weights, low-precision parameter copy and gradient clearing in
// ONE pass over memory (7 streams instead of torch's foreach chain + separate zero_grad + .half() cast).
// Reference behaviour: torch.optim.Adam(eps=1e-15) (main.py:245) under torch.cuda.amp.GradScaler
// (nerf/train_utils.py:897-904).
#include <cstdlib>
#include "common.cuh"

namespace ngp {
namespace {

template <typename G> __device__ __forceinline__ float load_g(const G* p, uint64_t i) { return to_f32(p[i]); }

// 4 elements per thread and iteration: 16-byte accesses for the fp32 streams (master, m, v), 8-byte for fp16/bf16 ones.
template <typename T> struct Vec4 { T v[4]; };
template <typename T> __device__ __forceinline__ Vec4<T> load4(const T* p) {
    Vec4<T> r;
    if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
    else *reinterpret_cast<uint2*>(r.v) = *reinterpret_cast<const uint2*>(p);
    return r;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const Vec4<T>& r) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
    else *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(r.v);
}

// The Adam step of one element: p -= step_size * m / (sqrt(v) / bias2_sqrt + eps).  The big streaming kernels run beside the
// ray marcher, and both are ISSUE bound (56 % / 62 % of the issue slots each when alone): an IEEE square root and two IEEE
// divisions are ~30 of this kernel's ~45 instructions per element.  NGP_ADAM_FAST_MATH (default) takes sqrt.approx, a
// multiplication by the reciprocal of bias2_sqrt and div.approx instead: relative error <= 2^-21 of the update, far below the
// fp16 resolution of the table it feeds (the fp32 master accumulates it like any rounding of the update).
#ifndef NGP_ADAM_FAST_MATH
#define NGP_ADAM_FAST_MATH 1
#endif
__device__ __forceinline__ float adam_delta(float m, float v, float step_size, float bias2_sqrt, float inv_bias2_sqrt, float eps) {
#if NGP_ADAM_FAST_MATH
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(v));
    return step_size * __fdividef(m, s * inv_bias2_sqrt + eps);
#else
    (void)inv_bias2_sqrt;
    return step_size * (m / (sqrtf(v) / bias2_sqrt + eps));
#endif
}

template <typename G, typename P, bool HasLP>
__global__ void __launch_bounds__(256)
fused_adam_kernel(float* __restrict__ master, P* __restrict__ param_lp, G* __restrict__ grad, float* __restrict__ m,
                  float* __restrict__ v, uint64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float bias1, float bias2_sqrt, const float* __restrict__ inv_scale_dev,
                  const float* __restrict__ found_inf_dev, bool zero_grad, const int* __restrict__ step_dev,
                  const float* __restrict__ lr_dev) {
    if (lr_dev) lr = __ldg(lr_dev);     // learning-rate schedule without re-capturing the graph (LambdaLR, main.py:258-261)
    const bool skip = found_inf_dev && (__ldg(found_inf_dev) != 0.f);
    const float inv_scale = inv_scale_dev ? __ldg(inv_scale_dev) : 1.f;
    if (step_dev) {     // step count kept on the device (CUDA-graph replay): bias corrections computed here, once per block
        __shared__ float s_bias[2];      // (two powf + a sqrtf per THREAD were ~6 % of the kernel's instructions)
        if (threadIdx.x == 0) {
            const float t = (float)max(__ldg(step_dev), 1);
            s_bias[0] = 1.f - powf(beta1, t);
            s_bias[1] = sqrtf(1.f - powf(beta2, t));
        }
        __syncthreads();
        bias1 = s_bias[0];
        bias2_sqrt = s_bias[1];
    }
    const float step_size = lr / bias1, inv_bias2_sqrt = 1.f / bias2_sqrt;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n4 = n / 4;      // the buffers are 16-byte aligned (checked by the host wrapper)
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const uint64_t i = q * 4;
        if (!skip) {
            const Vec4<G> g4 = load4(grad + i);
            Vec4<float> p4 = load4(master + i), m4 = load4(m + i), v4 = load4(v + i);
            Vec4<P> lp4;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float g = to_f32(g4.v[k]) * inv_scale;
                if (weight_decay != 0.f) g += weight_decay * p4.v[k];
                m4.v[k] = beta1 * m4.v[k] + (1.f - beta1) * g;
                v4.v[k] = beta2 * v4.v[k] + (1.f - beta2) * g * g;
                p4.v[k] -= adam_delta(m4.v[k], v4.v[k], step_size, bias2_sqrt, inv_bias2_sqrt, eps);
                if (HasLP) lp4.v[k] = from_f32<P>(p4.v[k]);
            }
            store4(m + i, m4); store4(v + i, v4); store4(master + i, p4);
            if (HasLP) store4(param_lp + i, lp4);
        }
        if (zero_grad) {
            Vec4<G> z;
#pragma unroll
            for (int k = 0; k < 4; k++) z.v[k] = from_f32<G>(0.f);
            store4(grad + i, z);
        }
    }
    // tail (n % 4 elements)
    for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!skip) {
            float g = to_f32(grad[i]) * inv_scale;
            float p = master[i];
            if (weight_decay != 0.f) g += weight_decay * p;
            const float mi = beta1 * m[i] + (1.f - beta1) * g;
            const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
            m[i] = mi;
            v[i] = vi;
            p -= adam_delta(mi, vi, step_size, bias2_sqrt, inv_bias2_sqrt, eps);
            master[i] = p;
            if (HasLP) param_lp[i] = from_f32<P>(p);
        }
        if (zero_grad) grad[i] = from_f32<G>(0.f);
    }
}

template <typename G>
__global__ void __launch_bounds__(256)
check_finite_kernel(const G* __restrict__ grad, uint64_t n, float* __restrict__ found_inf) {
    bool bad = false;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr uint32_t PER = 16 / sizeof(G);        // elements per 16-byte load
    const uint64_t nv = n / PER;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += stride) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(grad) + q);
        const G* e = reinterpret_cast<const G*>(&u);
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) bad |= !isfinite(to_f32(e[k]));
    }
    for (uint64_t i = nv * PER + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= !isfinite(to_f32(grad[i]));
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}


// ---- data parallel: reduce-scatter + Adam + all-gather in ONE kernel over NVLink peer memory ------------------------
// Every rank owns one contiguous shard [lo, hi) of the parameters (fp32 master, exp_avg, exp_avg_sq exist for the shard
// only).  For its shard it LOADS the gradients of all ranks straight from their buffers (peer-mapped symmetric memory,
// 16-byte loads over NVLink / NVSwitch), sums them in fp32 in rank order (the same order on every rank's shard, and only
// the owner computes: the result is written, not recomputed, so all ranks end up bit-identical), applies unscale + Adam,
// and STORES the updated low-precision parameters into the parameter buffer of every rank.  Compared with
// all-reduce -> replicated Adam this moves the same bytes over the links once, runs Adam on 1/world of the parameters
// and needs no staging buffer; the transfers overlap the arithmetic element by element.
// Cross-rank ordering (all gradients complete before the loads, all stores landed before the next forward) is provided
// by the caller's barriers on the same stream.
constexpr uint32_t kMaxPeers = 8;
struct PeerPtrs { const void* grad[kMaxPeers]; void* lp[kMaxPeers]; };

template <typename G, typename P>
__global__ void __launch_bounds__(256)
dp_fused_adam_kernel(const PeerPtrs peers, uint32_t world, uint32_t n_store, float* __restrict__ master, float* __restrict__ m,
                     float* __restrict__ v, uint64_t lo, uint64_t hi, float lr, float beta1, float beta2, float eps, float weight_decay,
                     const float* __restrict__ inv_scale_dev, const float* __restrict__ found_inf_dev, const int* __restrict__ step_dev,
                     const float* __restrict__ lr_dev, const float* __restrict__ flags, uint32_t n_flags) {
    if (found_inf_dev && __ldg(found_inf_dev) != 0.f) return;      // GradScaler: the whole step is skipped, on every rank alike
    // flags != nullptr: the ranks' inf / nan flags (published by peer stores before the barrier) are merged here and the step
    // count of THIS update is step_dev + 1 -- the counter itself is advanced by ngp_dp_finish after the kernel, so that no
    // block reads it while another writes it; saves the merge and counter launches on the critical side stream
    int step_now = __ldg(step_dev);
    if (flags) {
        bool bad = false;
        for (uint32_t r = 0; r < n_flags; r++) bad |= __ldcg(flags + r) != 0.f;
        if (bad) return;
        step_now += 1;
    }
    if (lr_dev) lr = __ldg(lr_dev);
    const float inv_scale = inv_scale_dev ? __ldg(inv_scale_dev) : 1.f;
    const float t = (float)max(step_now, 1);
    const float step_size = lr / (1.f - powf(beta1, t)), bias2_sqrt = sqrtf(1.f - powf(beta2, t));
    constexpr uint32_t PER = 16 / sizeof(G);           // elements per 16-byte gradient load
    const uint64_t n_vec = (hi - lo) / PER;            // the shard is a multiple of PER (host-checked)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_vec; q += stride) {
        const uint64_t i = lo + q * PER, s = q * PER;  // global element index, index inside the shard
        float g[PER];
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) g[k] = 0.f;
        uint4 u[kMaxPeers];
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; r++)       // all peer loads in flight before the first add
            if (r < world) u[r] = *reinterpret_cast<const uint4*>(reinterpret_cast<const G*>(peers.grad[r]) + i);
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; r++) {
            if (r < world) {
                const G* e = reinterpret_cast<const G*>(&u[r]);
#pragma unroll
                for (uint32_t k = 0; k < PER; k++) g[k] += to_f32(e[k]);
            }
        }
        P out[PER];
#pragma unroll
        for (uint32_t k0 = 0; k0 < PER; k0 += 4) {
            Vec4<float> p4 = load4(master + s + k0), m4 = load4(m + s + k0), v4 = load4(v + s + k0);
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                float gg = g[k0 + k] * inv_scale;
                if (weight_decay != 0.f) gg += weight_decay * p4.v[k];
                m4.v[k] = beta1 * m4.v[k] + (1.f - beta1) * gg;
                v4.v[k] = beta2 * v4.v[k] + (1.f - beta2) * gg * gg;
                p4.v[k] -= adam_delta(m4.v[k], v4.v[k], step_size, bias2_sqrt, 1.f / bias2_sqrt, eps);
                out[k0 + k] = from_f32<P>(p4.v[k]);
            }
            store4(m + s + k0, m4); store4(v + s + k0, v4); store4(master + s + k0, p4);
        }
        // PER low-precision parameters = PER * sizeof(P) bytes to every rank's copy (own copy included)
#pragma unroll
        for (uint32_t r = 0; r < kMaxPeers; r++) {
            if (r < n_store) {
                P* dst = reinterpret_cast<P*>(peers.lp[r]) + i;
                if constexpr (PER * sizeof(P) == 16) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
                else if constexpr (PER * sizeof(P) == 8) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(out);
                else { for (uint32_t k = 0; k < PER; k++) dst[k] = out[k]; }
            }
        }
    }
}

// found_inf of this rank -> slot `rank` of every rank's flag array (peer stores); merged after the barrier
__global__ void dp_publish_flag_kernel(const float* __restrict__ found_inf_local, PeerPtrs flags, uint32_t world, uint32_t rank) {
    if (threadIdx.x < world) reinterpret_cast<float*>(flags.lp[threadIdx.x])[rank] = *found_inf_local;
}
// end of a data-parallel update: merged flag -> found_inf (for the GradScaler update and the host), step counter, and the
// gradient buffers cleared for the next backward -- one launch instead of merge + counter + two memsets
__global__ void __launch_bounds__(256)
dp_finish_kernel(const float* __restrict__ flags, uint32_t world, float* __restrict__ found_inf, int* __restrict__ step_dev,
                 uint4* __restrict__ g0, uint64_t n0, uint4* __restrict__ g1, uint64_t n1) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float f = 0.f;
        for (uint32_t r = 0; r < world; r++) f = fmaxf(f, __ldcg(flags + r) != 0.f ? 1.f : 0.f);
        *found_inf = f;
        if (f == 0.f) *step_dev += 1;
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += stride) g0[i] = z;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) g1[i] = z;
}

__global__ void dp_merge_flags_kernel(const float* __restrict__ flags, uint32_t world, float* __restrict__ found_inf) {
    if (threadIdx.x == 0) {
        float f = 0.f;
        for (uint32_t r = 0; r < world; r++) f = fmaxf(f, flags[r] != 0.f ? 1.f : 0.f);
        *found_inf = f;
    }
}


// inf / nan check of up to four gradient buffers + GradScaler bookkeeping in ONE launch (the captured training step ran
// fill(found_inf) + check(table) + check(mlp) + step counter as four tiny kernels on the critical side stream).
// Blocks are split between the buffers in proportion to their size; every block publishes "bad" into scratch[1]; the last
// block to arrive (ticket scratch[0]) writes found_inf = 0 / 1 (overwrites: no zero-fill needed), counts the optimizer
// step when it is not skipped, and resets the scratch words for the next launch.
struct CheckBuffers { const void* p[4]; uint64_t n[4]; int dtype[4]; uint32_t first_block[5]; uint32_t count; };

template <typename G>
__device__ __forceinline__ bool block_has_nonfinite(const G* __restrict__ grad, uint64_t n, uint32_t block, uint32_t n_blocks) {
    bool bad = false;
    constexpr uint32_t PER = 16 / sizeof(G);
    const uint64_t nv = n / PER, stride = (uint64_t)n_blocks * blockDim.x;
    uint64_t q = (uint64_t)block * blockDim.x + threadIdx.x;
    // two 16-byte loads in flight per thread
    for (; q + stride < nv; q += 2 * stride) {
        const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(grad) + q), u1 = __ldg(reinterpret_cast<const uint4*>(grad) + q + stride);
        const G* e0 = reinterpret_cast<const G*>(&u0);
        const G* e1 = reinterpret_cast<const G*>(&u1);
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) bad |= !isfinite(to_f32(e0[k])) || !isfinite(to_f32(e1[k]));
    }
    for (; q < nv; q += stride) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(grad) + q);
        const G* e = reinterpret_cast<const G*>(&u);
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) bad |= !isfinite(to_f32(e[k]));
    }
    for (uint64_t i = nv * PER + (uint64_t)block * blockDim.x + threadIdx.x; i < n; i += stride) bad |= !isfinite(to_f32(grad[i]));
    return bad;
}

__global__ void __launch_bounds__(256)
check_finite_multi_kernel(const CheckBuffers b, float* __restrict__ found_inf, int* __restrict__ step_dev, uint32_t* __restrict__ scratch,
                          const PeerPtrs peer_flags, uint32_t world, uint32_t rank) {
    uint32_t which = 0;
    while (which + 1 < b.count && blockIdx.x >= b.first_block[which + 1]) which++;
    const uint32_t block = blockIdx.x - b.first_block[which], n_blocks = b.first_block[which + 1] - b.first_block[which];
    bool bad;
    if (b.dtype[which] == NGP_F32) bad = block_has_nonfinite(reinterpret_cast<const float*>(b.p[which]), b.n[which], block, n_blocks);
    else if (b.dtype[which] == NGP_F16) bad = block_has_nonfinite(reinterpret_cast<const __half*>(b.p[which]), b.n[which], block, n_blocks);
    else bad = block_has_nonfinite(reinterpret_cast<const __nv_bfloat16*>(b.p[which]), b.n[which], block, n_blocks);
    const bool any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (any_bad) atomicOr(scratch + 1, 1u);
        __threadfence();
        if (atomicAdd(scratch, 1u) == gridDim.x - 1) {          // last block: all flags are in
            __threadfence();
            const bool inf = *reinterpret_cast<volatile uint32_t*>(scratch + 1) != 0u;
            *found_inf = inf ? 1.0f : 0.0f;
            // data parallel: this rank's flag goes into slot `rank` of every rank's flag array (peer stores, ordered by the
            // barrier that follows on the stream)
            for (uint32_t r = 0; r < world; r++) reinterpret_cast<float*>(peer_flags.lp[r])[rank] = inf ? 1.0f : 0.0f;
            if (!inf && step_dev) *step_dev += 1;                 // a skipped GradScaler step is not counted
            scratch[0] = 0u; scratch[1] = 0u;
        }
    }
}

// A whole GradScaler + Adam step of a SMALL tensor (the se3 pose corrections: 600 values) in one single-block launch:
// inf / nan check, step count, unscale + Adam, gradient clear.  The pose update sits on the main stream in front of the ray
// generation of the next step, so its four tiny launches (fill, check, count, Adam) were on the critical path.
__global__ void __launch_bounds__(256)
small_adam_kernel(float* __restrict__ master, float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v, uint32_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int* __restrict__ step_dev, const float* __restrict__ lr_dev,
                  const float* __restrict__ inv_scale_dev, float* __restrict__ found_inf_out) {
    bool bad = false;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) bad |= !isfinite(grad[i]);
    const bool skip = __syncthreads_or(bad);
    __shared__ int s_step;
    if (threadIdx.x == 0) {
        if (!skip) *step_dev += 1;
        s_step = *step_dev;
        if (found_inf_out) *found_inf_out = skip ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (lr_dev) lr = __ldg(lr_dev);
    const float inv_scale = inv_scale_dev ? __ldg(inv_scale_dev) : 1.f;
    const float t = (float)max(s_step, 1);
    const float step_size = lr / (1.f - powf(beta1, t)), bias2_sqrt = sqrtf(1.f - powf(beta2, t));
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (!skip) {
            float g = grad[i] * inv_scale;
            float p = master[i];
            if (weight_decay != 0.f) g += weight_decay * p;
            const float mi = beta1 * m[i] + (1.f - beta1) * g;
            const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
            m[i] = mi; v[i] = vi;
            master[i] = p - step_size * (mi / (sqrtf(vi) / bias2_sqrt + eps));
        }
        grad[i] = 0.f;
    }
}

// The same for data parallel ranks over peer memory: the gradient is the SUM of every rank's buffer, read through the peer
// mappings (the caller's barrier has made them complete); every rank computes the identical update of its own replica, so no
// broadcast follows.  The buffers are NOT cleared here -- peers may still be reading them (ngp_dp_finish clears after the
// closing barrier).  Replaces an NCCL all-reduce of 600 floats in front of the ray generation.
struct SmallPeers { const float* grad[kMaxPeers]; };
__global__ void __launch_bounds__(256)
dp_small_adam_kernel(const SmallPeers peers, uint32_t world, float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                     uint32_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int* __restrict__ step_dev,
                     const float* __restrict__ lr_dev, const float* __restrict__ inv_scale_dev, float* __restrict__ found_inf_out) {
    bool bad = false;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        float g = 0.f;
        for (uint32_t r = 0; r < world; r++) g += __ldcg(peers.grad[r] + i);      // fixed rank order: bit-identical on every rank
        bad |= !isfinite(g);
    }
    const bool skip = __syncthreads_or(bad);
    __shared__ int s_step;
    if (threadIdx.x == 0) {
        if (!skip) *step_dev += 1;
        s_step = *step_dev;
        if (found_inf_out) *found_inf_out = skip ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (skip) return;
    if (lr_dev) lr = __ldg(lr_dev);
    const float inv_scale = inv_scale_dev ? __ldg(inv_scale_dev) : 1.f;
    const float t = (float)max(s_step, 1);
    const float step_size = lr / (1.f - powf(beta1, t)), bias2_sqrt = sqrtf(1.f - powf(beta2, t));
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        float g = 0.f;
        for (uint32_t r = 0; r < world; r++) g += __ldcg(peers.grad[r] + i);
        g *= inv_scale;
        const float p = master[i];
        if (weight_decay != 0.f) g += weight_decay * p;
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi; v[i] = vi;
        master[i] = p - step_size * (mi / (sqrtf(vi) / bias2_sqrt + eps));
    }
}

// GradScaler semantics for a device-side step counter: the optimizer step is counted only when it is not skipped
__global__ void adam_step_counter_kernel(int* __restrict__ step_dev, const float* __restrict__ found_inf_dev) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !(found_inf_dev && *found_inf_dev != 0.f)) *step_dev += 1;
}

// torch.amp.GradScaler.update() on the device (grad_scaler.py: _amp_update_scale_): state = {scale, inv_scale = 1 / (scale *
// world), growth tracker, skipped-step count}.  A step whose gradients held inf / nan halves the scale (backoff) and resets the
// tracker; `growth_interval` clean steps in a row double it.
__global__ void grad_scaler_update_kernel(float* __restrict__ scale, float* __restrict__ inv_scale, int* __restrict__ state,
                                          const float* __restrict__ found_a, const float* __restrict__ found_b, float growth, float backoff,
                                          int growth_interval, float world) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool bad = (found_a && *found_a != 0.f) || (found_b && *found_b != 0.f);
    float sc = *scale;
    if (bad) {
        sc *= backoff;
        state[0] = 0;
        state[1] += 1;                       // skipped steps, for the host to look at whenever it likes
    } else if (++state[0] >= growth_interval) {
        const float grown = sc * growth;
        if (isfinite(grown)) sc = grown;
        state[0] = 0;
    }
    *scale = sc;
    *inv_scale = 1.0f / (sc * world);
}

}  // namespace
}  // namespace ngp

using namespace ngp;

// Resident blocks per SM of the streaming optimizer kernels (grid-stride, 256 threads).  They run on the side stream beside the
// ray marcher of the next step.  Measured at configs[1] (NGP_ADAM_BLOCKS_PER_SM, step time between occupancy updates, round 2 with
// the 512-point marcher window): 8 -> 679.9 us, 6 -> 678.4, 5 -> 677.1, 4 -> 678.3, 3 -> 683.9.  With 8 x 256 threads the inf / nan
// check that opens the chain fills every thread slot of the SM and the marcher's first blocks wait ~9 us for it to drain; five
// blocks leave room for them and are still enough to stream at full HBM rate (Adam alone: 76 us with 5, 80 with 8).
static uint32_t adam_blocks_per_sm() {
    static const uint32_t v = [] {
        const char* e = getenv("NGP_ADAM_BLOCKS_PER_SM");
        const int x = e ? atoi(e) : 5;
        return (uint32_t)(x >= 1 && x <= 8 ? x : 5);
    }();
    return v;
}
#define kAdamBlocksPerSM adam_blocks_per_sm()

extern "C" int ngp_grad_scaler_update(float* scale_dev, float* inv_scale_dev, int32_t* state_dev, const float* found_inf_a,
                                      const float* found_inf_b, float growth_factor, float backoff_factor, int growth_interval,
                                      uint32_t world, ngp_stream_t stream) {
    if (!scale_dev || !inv_scale_dev || !state_dev) return NGP_ERR_NULL;
    if (!(growth_factor >= 1.f) || !(backoff_factor > 0.f && backoff_factor <= 1.f) || growth_interval < 1 || world == 0) return NGP_ERR_BAD_ARG;
    grad_scaler_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scale_dev, inv_scale_dev, state_dev, found_inf_a, found_inf_b, growth_factor,
                                                                 backoff_factor, growth_interval, (float)world);
    return finish_launch();
}

extern "C" int ngp_fused_adam(float* master, void* param_lp, int lp_dtype, void* grad, int grad_dtype, float* exp_avg,
                              float* exp_avg_sq, uint64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, uint32_t step, const int32_t* step_dev, const float* lr_dev,
                              const float* inv_scale_dev, const float* found_inf_dev, int zero_grad, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!master || !grad || !exp_avg || !exp_avg_sq) return NGP_ERR_NULL;
    if (step == 0 && !step_dev) return NGP_ERR_BAD_ARG;
    if (grad_dtype < NGP_F32 || grad_dtype > NGP_BF16) return NGP_ERR_BAD_DTYPE;
    if (param_lp && (lp_dtype != NGP_F16 && lp_dtype != NGP_BF16)) return NGP_ERR_BAD_DTYPE;
    if (!aligned(master, 16) || !aligned(grad, 16) || !aligned(exp_avg, 16) || !aligned(exp_avg_sq, 16) || (param_lp && !aligned(param_lp, 8)))
        return NGP_ERR_ALIGN;
    const float bias1 = 1.f - powf(beta1, (float)std::max(step, 1u));
    const float bias2_sqrt = sqrtf(1.f - powf(beta2, (float)std::max(step, 1u)));
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(div_up<uint64_t>(div_up<uint64_t>(n, 4), 256), (uint64_t)kNumSMs * kAdamBlocksPerSM);
    cudaStream_t st = (cudaStream_t)stream;
#define NGP_ADAM(G, P, HAS)                                                                                      \
    fused_adam_kernel<G, P, HAS><<<blocks, 256, 0, st>>>(master, (P*)param_lp, (G*)grad, exp_avg, exp_avg_sq, n, lr, \
                                                         beta1, beta2, eps, weight_decay, bias1, bias2_sqrt,      \
                                                         inv_scale_dev, found_inf_dev, zero_grad != 0, step_dev, lr_dev)
#define NGP_ADAM_G(G)                                                             \
    if (!param_lp) NGP_ADAM(G, __half, false);                                    \
    else if (lp_dtype == NGP_F16) NGP_ADAM(G, __half, true);                      \
    else NGP_ADAM(G, __nv_bfloat16, true)
    if (grad_dtype == NGP_F32) { NGP_ADAM_G(float); }
    else if (grad_dtype == NGP_F16) { NGP_ADAM_G(__half); }
    else { NGP_ADAM_G(__nv_bfloat16); }
#undef NGP_ADAM_G
#undef NGP_ADAM
    return finish_launch();
}


extern "C" int ngp_dp_fused_adam(const void* const* peer_grads, int grad_dtype, void* const* peer_params_lp, int lp_dtype,
                                 uint32_t world, uint32_t n_store, float* master_shard, float* exp_avg_shard,
                                 float* exp_avg_sq_shard, uint64_t lo, uint64_t hi, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, const int32_t* step_dev, const float* lr_dev, const float* inv_scale_dev,
                                 const float* found_inf_dev, const float* flags, uint32_t n_flags, ngp_stream_t stream) {
    if (hi <= lo) return NGP_OK;
    if (flags && (n_flags == 0 || n_flags > kMaxPeers)) return NGP_ERR_BAD_ARG;
    if (!peer_grads || !peer_params_lp || !master_shard || !exp_avg_shard || !exp_avg_sq_shard || !step_dev) return NGP_ERR_NULL;
    if (world == 0 || world > kMaxPeers || n_store > world) return NGP_ERR_BAD_ARG;
    if (grad_dtype != NGP_F32 && grad_dtype != NGP_F16) return NGP_ERR_BAD_DTYPE;
    if (lp_dtype != NGP_F32 && lp_dtype != NGP_F16) return NGP_ERR_BAD_DTYPE;
    const uint32_t per = grad_dtype == NGP_F16 ? 8u : 4u;
    if (lo % per || (hi - lo) % per) return NGP_ERR_ALIGN;
    PeerPtrs pp = {};
    for (uint32_t r = 0; r < world; r++) {
        if (!peer_grads[r] || !aligned(peer_grads[r], 16)) return NGP_ERR_ALIGN;
        pp.grad[r] = peer_grads[r];
        if (r < n_store) {
            if (!peer_params_lp[r] || !aligned(peer_params_lp[r], 16)) return NGP_ERR_ALIGN;
            pp.lp[r] = peer_params_lp[r];
        }
    }
    if (!aligned(master_shard, 16) || !aligned(exp_avg_shard, 16) || !aligned(exp_avg_sq_shard, 16)) return NGP_ERR_ALIGN;
    const uint64_t n_vec = (hi - lo) / per;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(div_up<uint64_t>(n_vec, 256), (uint64_t)kNumSMs * 4);
    cudaStream_t st = (cudaStream_t)stream;
#define NGP_DP(G, P) dp_fused_adam_kernel<G, P><<<blocks, 256, 0, st>>>(pp, world, n_store, master_shard, exp_avg_shard, exp_avg_sq_shard, \
                                                                        lo, hi, lr, beta1, beta2, eps, weight_decay, inv_scale_dev,     \
                                                                        found_inf_dev, step_dev, lr_dev, flags, n_flags)
    if (grad_dtype == NGP_F16 && lp_dtype == NGP_F16) NGP_DP(__half, __half);
    else if (grad_dtype == NGP_F32 && lp_dtype == NGP_F16) NGP_DP(float, __half);
    else if (grad_dtype == NGP_F32 && lp_dtype == NGP_F32) NGP_DP(float, float);
    else return NGP_ERR_UNSUPPORTED;
#undef NGP_DP
    return finish_launch();
}

extern "C" int ngp_dp_publish_flag(const float* found_inf_local, void* const* peer_flags, uint32_t world, uint32_t rank,
                                   ngp_stream_t stream) {
    if (!found_inf_local || !peer_flags) return NGP_ERR_NULL;
    if (world == 0 || world > kMaxPeers || rank >= world) return NGP_ERR_BAD_ARG;
    PeerPtrs pp = {};
    for (uint32_t r = 0; r < world; r++) { if (!peer_flags[r]) return NGP_ERR_NULL; pp.lp[r] = peer_flags[r]; }
    dp_publish_flag_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(found_inf_local, pp, world, rank);
    return finish_launch();
}

extern "C" int ngp_dp_merge_flags(const float* flags, uint32_t world, float* found_inf, ngp_stream_t stream) {
    if (!flags || !found_inf) return NGP_ERR_NULL;
    if (world == 0 || world > kMaxPeers) return NGP_ERR_BAD_ARG;
    dp_merge_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, world, found_inf);
    return finish_launch();
}

extern "C" int ngp_small_adam(float* master, float* grad, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int32_t* step_dev, const float* lr_dev,
                              const float* inv_scale_dev, float* found_inf_out, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!master || !grad || !exp_avg || !exp_avg_sq || !step_dev) return NGP_ERR_NULL;
    if (n > (1u << 20)) return NGP_ERR_BAD_ARG;      // one block: meant for small tensors
    small_adam_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(master, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                          step_dev, lr_dev, inv_scale_dev, found_inf_out);
    return finish_launch();
}

extern "C" int ngp_dp_small_adam(const void* const* peer_grads, uint32_t world, float* master, float* exp_avg, float* exp_avg_sq,
                                 uint32_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t* step_dev,
                                 const float* lr_dev, const float* inv_scale_dev, float* found_inf_out, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!peer_grads || !master || !exp_avg || !exp_avg_sq || !step_dev) return NGP_ERR_NULL;
    if (world == 0 || world > kMaxPeers || n > (1u << 20)) return NGP_ERR_BAD_ARG;
    SmallPeers pp = {};
    for (uint32_t r = 0; r < world; r++) {
        if (!peer_grads[r]) return NGP_ERR_NULL;
        if (!aligned(peer_grads[r], 4)) return NGP_ERR_ALIGN;
        pp.grad[r] = (const float*)peer_grads[r];
    }
    dp_small_adam_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pp, world, master, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                             step_dev, lr_dev, inv_scale_dev, found_inf_out);
    return finish_launch();
}

extern "C" int ngp_adam_step_counter(int32_t* step_dev, const float* found_inf_dev, ngp_stream_t stream) {
    if (!step_dev) return NGP_ERR_NULL;
    adam_step_counter_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_dev, found_inf_dev);
    return finish_launch();
}


extern "C" int ngp_check_finite_multi(const void* const* grads, const int* dtypes, const uint64_t* counts, uint32_t n_buffers,
                                      float* found_inf_dev, int32_t* step_dev, uint32_t* scratch, ngp_stream_t stream) {
    if (!grads || !dtypes || !counts || !found_inf_dev || !scratch) return NGP_ERR_NULL;
    if (n_buffers == 0 || n_buffers > 4) return NGP_ERR_BAD_ARG;
    CheckBuffers b = {};
    b.count = n_buffers;
    uint32_t blocks = 0;
    for (uint32_t i = 0; i < n_buffers; i++) {
        if (counts[i] && !grads[i]) return NGP_ERR_NULL;
        if (dtypes[i] < NGP_F32 || dtypes[i] > NGP_BF16) return NGP_ERR_BAD_DTYPE;
        if (!aligned(grads[i], 16)) return NGP_ERR_ALIGN;
        b.p[i] = grads[i]; b.n[i] = counts[i]; b.dtype[i] = dtypes[i];
        b.first_block[i] = blocks;
        blocks += (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(counts[i], 256 * 16), (uint64_t)kNumSMs * kAdamBlocksPerSM));
    }
    b.first_block[n_buffers] = blocks;
    check_finite_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b, found_inf_dev, step_dev, scratch, PeerPtrs{}, 0, 0);
    return finish_launch();
}

extern "C" int ngp_dp_check_publish(const void* const* grads, const int* dtypes, const uint64_t* counts, uint32_t n_buffers,
                                    float* found_inf_dev, uint32_t* scratch, void* const* peer_flags, uint32_t world, uint32_t rank,
                                    ngp_stream_t stream) {
    if (!grads || !dtypes || !counts || !found_inf_dev || !scratch || !peer_flags) return NGP_ERR_NULL;
    if (n_buffers == 0 || n_buffers > 4 || world == 0 || world > kMaxPeers || rank >= world) return NGP_ERR_BAD_ARG;
    CheckBuffers b = {};
    b.count = n_buffers;
    uint32_t blocks = 0;
    for (uint32_t i = 0; i < n_buffers; i++) {
        if (counts[i] && !grads[i]) return NGP_ERR_NULL;
        if (dtypes[i] < NGP_F32 || dtypes[i] > NGP_BF16) return NGP_ERR_BAD_DTYPE;
        if (!aligned(grads[i], 16)) return NGP_ERR_ALIGN;
        b.p[i] = grads[i]; b.n[i] = counts[i]; b.dtype[i] = dtypes[i];
        b.first_block[i] = blocks;
        blocks += (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(counts[i], 256 * 16), (uint64_t)kNumSMs * kAdamBlocksPerSM));
    }
    b.first_block[n_buffers] = blocks;
    PeerPtrs pf = {};
    for (uint32_t r = 0; r < world; r++) { if (!peer_flags[r]) return NGP_ERR_NULL; pf.lp[r] = peer_flags[r]; }
    check_finite_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b, found_inf_dev, nullptr, scratch, pf, world, rank);
    return finish_launch();
}

extern "C" int ngp_dp_finish(const float* flags, uint32_t world, float* found_inf_dev, int32_t* step_dev, void* grad0, uint64_t bytes0,
                             void* grad1, uint64_t bytes1, ngp_stream_t stream) {
    if (!flags || !found_inf_dev || !step_dev) return NGP_ERR_NULL;
    if (world == 0 || world > kMaxPeers || bytes0 % 16 || bytes1 % 16) return NGP_ERR_BAD_ARG;
    if ((bytes0 && !aligned(grad0, 16)) || (bytes1 && !aligned(grad1, 16))) return NGP_ERR_ALIGN;
    const uint64_t n0 = bytes0 / 16, n1 = bytes1 / 16;
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(div_up<uint64_t>(std::max(n0, n1), 256), (uint64_t)kNumSMs * 8));
    dp_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(flags, world, found_inf_dev, step_dev, (uint4*)grad0, n0, (uint4*)grad1, n1);
    return finish_launch();
}

extern "C" int ngp_check_finite(const void* grad, int grad_dtype, uint64_t n, float* found_inf_dev, ngp_stream_t stream) {
    if (n == 0) return NGP_OK;
    if (!grad || !found_inf_dev) return NGP_ERR_NULL;
    if (!aligned(grad, 16)) return NGP_ERR_ALIGN;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(div_up<uint64_t>(n, 256 * 8), (uint64_t)kNumSMs * 8);
    cudaStream_t st = (cudaStream_t)stream;
    switch (grad_dtype) {
        case NGP_F32: check_finite_kernel<float><<<blocks, 256, 0, st>>>((const float*)grad, n, found_inf_dev); break;
        case NGP_F16: check_finite_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)grad, n, found_inf_dev); break;
        case NGP_BF16: check_finite_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)grad, n, found_inf_dev); break;
        default: return NGP_ERR_BAD_DTYPE;
    }
    return finish_launch();
}
