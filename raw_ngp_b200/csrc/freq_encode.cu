// freq_encode.cu -- frequency (positional) encoder, the remaining entry of encoding.get_encoder.
//
// Semantics of kernel_freq / kernel_freq_backward (freqencoder/src/freqencoder.cu:31-94):
//   out[b, d]                      = x[b, d]                                        d < D
//   out[b, D + (2f + s) D + d]     = __sinf(x[b, d] * 2^f + s * pi/2)               f < deg, s in {0 (sin), 1 (cos)}
//   d x[b, d] = g[b, d] + sum_f 2^f (g_sin * out_cos - g_cos * out_sin)             (the saved outputs are the derivatives)
// HBM bound: 4 D bytes in, 4 C bytes out per point (C = D + 2 D deg).  Forward: one thread per output element (coalesced
// stores, the point's D inputs come from L1); backward: one thread per (point, dim), every row of grad / outputs is read
// once.  The fast-math sine (__sinf) is part of the semantics: the same intrinsic gives bit-identical outputs.
#include "common.cuh"

namespace ngp {
namespace {

__global__ void __launch_bounds__(256)
freq_forward_kernel(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t C, float* __restrict__ outputs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * C) return;
    const uint32_t b = (uint32_t)(t / C), c = (uint32_t)(t - (uint64_t)b * C);
    const float* x = inputs + (size_t)b * D;
    if (c < D) {
        outputs[t] = __ldg(x + c);
    } else {
        const uint32_t col = c / D - 1, d = c % D;
        const float phase = (float)(col & 1u) * (3.141592653589793f / 2);
        outputs[t] = __sinf(scalbnf(__ldg(x + d), (int)(col >> 1)) + phase);
    }
}

__global__ void __launch_bounds__(256)
freq_backward_kernel(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                     float* __restrict__ grad_inputs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (uint64_t)b * D);
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float result = __ldg(g + d);
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        result += scalbnf(1.0f, (int)f) * (__ldg(g + d) * __ldg(o + D + d) - __ldg(g + D + d) * __ldg(o + d));
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = result;
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C, float* outputs,
                                       ngp_stream_t stream) {
    if (C != D + 2 * D * degree || D == 0) return NGP_ERR_BAD_ARG;
    if (B == 0) return NGP_OK;
    if (!inputs || !outputs) return NGP_ERR_NULL;
    const uint64_t n = (uint64_t)B * C;
    if (div_up<uint64_t>(n, 256) > 0x7FFFFFFFull) return NGP_ERR_BAD_ARG;
    freq_forward_kernel<<<(uint32_t)div_up<uint64_t>(n, 256), 256, 0, (cudaStream_t)stream>>>(inputs, B, D, C, outputs);
    return finish_launch();
}

extern "C" int ngp_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C,
                                        float* grad_inputs, ngp_stream_t stream) {
    if (C != D + 2 * D * degree || D == 0) return NGP_ERR_BAD_ARG;
    if (B == 0) return NGP_OK;
    if (!grad || !outputs || !grad_inputs) return NGP_ERR_NULL;
    const uint64_t n = (uint64_t)B * D;
    freq_backward_kernel<<<(uint32_t)div_up<uint64_t>(n, 256), 256, 0, (cudaStream_t)stream>>>(grad, outputs, B, D, degree, C, grad_inputs);
    return finish_launch();
}
