// sh_encode.cu -- real spherical-harmonics encoder (degree 1..8) for view and light directions.
//
// Semantics of kernel_sh / kernel_sh_backward (shencoder/src/shencoder.cu:27-355, 358-382).  The basis
// comes from sh_basis.inc (generated, factored form); the backward pass recomputes the Jacobian from the
// 12-byte input instead of reading back a saved dy_dx [B, 3*deg^2] (192 B/sample at degree 4).
#include "common.cuh"

namespace ngp {
namespace {

constexpr uint32_t kShThreads = 128;

template <typename OutT, int DEG, bool Jac>
__global__ void __launch_bounds__(kShThreads)
sh_forward_kernel(const float* __restrict__ inputs, OutT* __restrict__ outputs, float* __restrict__ dy_dx, uint32_t B) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr int N = DEG * DEG;
    const float x = __ldg(inputs + (size_t)b * 3), y = __ldg(inputs + (size_t)b * 3 + 1), z = __ldg(inputs + (size_t)b * 3 + 2);
    const float zz = z * z;
    float val[N];
    float* jx = Jac ? dy_dx + (size_t)b * 3 * N : nullptr;
#define SH_TERM(i, v, ddx, ddy, ddz)                                           \
    val[i] = (v);                                                              \
    if (Jac) { jx[i] = (ddx); jx[N + i] = (ddy); jx[2 * N + i] = (ddz); }
#include "sh_basis.inc"
#undef SH_TERM
    OutT* out = outputs + (size_t)b * N;
    if constexpr (std::is_same<OutT, float>::value && (N % 4 == 0)) {
#pragma unroll
        for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(out + i) = make_float4(val[i], val[i + 1], val[i + 2], val[i + 3]);
    } else if constexpr (!std::is_same<OutT, float>::value && (N % 8 == 0)) {
#pragma unroll
        for (int i = 0; i < N; i += 8) {
            OutT t[8];
#pragma unroll
            for (int j = 0; j < 8; j++) t[j] = from_f32<OutT>(val[i + j]);
            *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(t);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = from_f32<OutT>(val[i]);
    }
}

template <typename GradT, int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_backward_kernel(const GradT* __restrict__ grad, const float* __restrict__ inputs, float* __restrict__ grad_inputs, uint32_t B) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr int N = DEG * DEG;
    const float x = __ldg(inputs + (size_t)b * 3), y = __ldg(inputs + (size_t)b * 3 + 1), z = __ldg(inputs + (size_t)b * 3 + 2);
    const float zz = z * z;
    float g[N];
    const GradT* gp = grad + (size_t)b * N;
    if constexpr (std::is_same<GradT, float>::value && (N % 4 == 0)) {
#pragma unroll
        for (int i = 0; i < N; i += 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(gp + i));
            g[i] = t.x; g[i + 1] = t.y; g[i + 2] = t.z; g[i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) g[i] = to_f32(gp[i]);
    }
    float gx = 0.f, gy = 0.f, gz = 0.f;
#define SH_TERM(i, v, ddx, ddy, ddz) \
    gx += g[i] * (ddx); gy += g[i] * (ddy); gz += g[i] * (ddz);
#include "sh_basis.inc"
#undef SH_TERM
    float* o = grad_inputs + (size_t)b * 3;
    o[0] += gx; o[1] += gy; o[2] += gz;
}

template <typename OutT, int DEG>
void launch_fwd(const float* in, OutT* out, float* jac, uint32_t B, cudaStream_t st) {
    const uint32_t blocks = div_up(B, kShThreads);
    if (jac) sh_forward_kernel<OutT, DEG, true><<<blocks, kShThreads, 0, st>>>(in, out, jac, B);
    else sh_forward_kernel<OutT, DEG, false><<<blocks, kShThreads, 0, st>>>(in, out, jac, B);
}

template <typename OutT>
int dispatch_fwd(const float* in, OutT* out, float* jac, uint32_t B, uint32_t degree, cudaStream_t st) {
    switch (degree) {
        case 1: launch_fwd<OutT, 1>(in, out, jac, B, st); break;
        case 2: launch_fwd<OutT, 2>(in, out, jac, B, st); break;
        case 3: launch_fwd<OutT, 3>(in, out, jac, B, st); break;
        case 4: launch_fwd<OutT, 4>(in, out, jac, B, st); break;
        case 5: launch_fwd<OutT, 5>(in, out, jac, B, st); break;
        case 6: launch_fwd<OutT, 6>(in, out, jac, B, st); break;
        case 7: launch_fwd<OutT, 7>(in, out, jac, B, st); break;
        case 8: launch_fwd<OutT, 8>(in, out, jac, B, st); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
    return finish_launch();
}

template <typename GradT>
int dispatch_bwd(const GradT* grad, const float* in, float* gin, uint32_t B, uint32_t degree, cudaStream_t st) {
    const uint32_t blocks = div_up(B, kShThreads);
    switch (degree) {
        case 1: sh_backward_kernel<GradT, 1><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 2: sh_backward_kernel<GradT, 2><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 3: sh_backward_kernel<GradT, 3><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 4: sh_backward_kernel<GradT, 4><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 5: sh_backward_kernel<GradT, 5><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 6: sh_backward_kernel<GradT, 6><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 7: sh_backward_kernel<GradT, 7><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        case 8: sh_backward_kernel<GradT, 8><<<blocks, kShThreads, 0, st>>>(grad, in, gin, B); break;
        default: return NGP_ERR_UNSUPPORTED;
    }
    return finish_launch();
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_sh_encode_forward(const float* inputs, void* outputs, uint32_t B, uint32_t degree, float* dy_dx,
                                     int out_dtype, ngp_stream_t stream) {
    if (degree < 1 || degree > 8) return NGP_ERR_UNSUPPORTED;
    if (B == 0) return NGP_OK;
    if (!inputs || !outputs) return NGP_ERR_NULL;
    if (!aligned(outputs, 16) || !aligned(inputs, 4) || (dy_dx && !aligned(dy_dx, 4))) return NGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_dtype) {
        case NGP_F32: return dispatch_fwd<float>(inputs, (float*)outputs, dy_dx, B, degree, st);
        case NGP_F16: return dispatch_fwd<__half>(inputs, (__half*)outputs, dy_dx, B, degree, st);
        case NGP_BF16: return dispatch_fwd<__nv_bfloat16>(inputs, (__nv_bfloat16*)outputs, dy_dx, B, degree, st);
        default: return NGP_ERR_BAD_DTYPE;
    }
}

extern "C" int ngp_sh_encode_backward(const void* grad, const float* inputs, uint32_t B, uint32_t degree,
                                      float* grad_inputs, int grad_dtype, ngp_stream_t stream) {
    if (degree < 1 || degree > 8) return NGP_ERR_UNSUPPORTED;
    if (B == 0) return NGP_OK;
    if (!grad || !inputs || !grad_inputs) return NGP_ERR_NULL;
    if (!aligned(grad, 16) || !aligned(inputs, 4) || !aligned(grad_inputs, 4)) return NGP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    switch (grad_dtype) {
        case NGP_F32: return dispatch_bwd<float>((const float*)grad, inputs, grad_inputs, B, degree, st);
        case NGP_F16: return dispatch_bwd<__half>((const __half*)grad, inputs, grad_inputs, B, degree, st);
        case NGP_BF16: return dispatch_bwd<__nv_bfloat16>((const __nv_bfloat16*)grad, inputs, grad_inputs, B, degree, st);
        default: return NGP_ERR_BAD_DTYPE;
    }
}


namespace ngp {
namespace {
__global__ void sh_dirs_backward_kernel(const __half* __restrict__ d_in2, uint32_t ld2, uint32_t col0, const float* __restrict__ dirs,
                                        uint32_t M, const int* __restrict__ m_dev, float* __restrict__ d_dirs) {
    if (m_dev) M = min(M, (uint32_t)__ldg(m_dev));
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    float gs[16];
#pragma unroll
    for (int k = 0; k < 16; k++) gs[k] = __half2float(d_in2[(size_t)row * ld2 + col0 + k]);
    const float d0 = __ldg(dirs + (size_t)row * 3), d1 = __ldg(dirs + (size_t)row * 3 + 1), d2 = __ldg(dirs + (size_t)row * 3 + 2);
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
    d_dirs[(size_t)row * 3] = (gx - u0 * t) * inv0;
    d_dirs[(size_t)row * 3 + 1] = (gy - u1 * t) * inv0;
    d_dirs[(size_t)row * 3 + 2] = (gz - u2 * t) * inv0;
}
}  // namespace
}  // namespace ngp

extern "C" int ngp_sh_dirs_backward(const void* d_in2, uint32_t ld2, uint32_t col0, const float* dirs, uint32_t M,
                                    const int32_t* m_dev, float* d_dirs, ngp_stream_t stream) {
    if (M == 0) return NGP_OK;
    if (!d_in2 || !dirs || !d_dirs) return NGP_ERR_NULL;
    if (col0 + 16 > ld2) return NGP_ERR_BAD_ARG;
    ngp::sh_dirs_backward_kernel<<<ngp::div_up(M, 256u), 256, 0, (cudaStream_t)stream>>>((const __half*)d_in2, ld2, col0, dirs, M, m_dev, d_dirs);
    return ngp::finish_launch();
}
