// diag.cu -- ceilings of the two units that bound the field kernels, measured on the running GPU (tools/l2_ceiling.py):
// the rate at which the SMs can pull random 4-byte rows out of an L2-resident table (what the forward's corner gathers do)
// and the rate at which L2 can apply random packed-fp16 reductions to one (what the backward's scatter does).
// Not used by any operator.
#include "common.cuh"

namespace ngp {
namespace {

__device__ __forceinline__ uint32_t mix(uint32_t x) {      // cheap integer hash: independent addresses per op
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// mode 0: 8 independent 4-byte gathers per thread and round (like one level of one sample); mode 1: 8 red.add.f16x2 to random rows;
// mode 2: 4 red.add.v2.f16x2 to random aligned row pairs (8 rows' worth of payload in 4 operations)
__global__ void __launch_bounds__(512)
l2_rate_kernel(uint32_t* __restrict__ table, uint32_t mask, uint32_t rounds, int mode, uint32_t* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t h = mix(tid * 2654435761u + r * 40503u);
        if (mode == 0) {
            uint32_t v[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) v[k] = __ldg(table + ((h + k * 0x9E3779B9u) & mask));
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) acc += v[k];
        } else if (mode == 1) {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) red_add_h2(reinterpret_cast<__half*>(table + ((h + k * 0x9E3779B9u) & mask)), 0u);
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) red_add_v2_h2(reinterpret_cast<__half*>(table + ((h + k * 0x9E3779B9u) & mask & ~1u)), 0u, 0u);
        }
    }
    if (acc == 0x12345678u) *sink = acc;     // keeps the loads alive
}

}  // namespace
}  // namespace ngp

using namespace ngp;

/* table: n_rows (a power of two) 4-byte rows; launches `blocks` x 512 threads, each doing `rounds` rounds of 8 row operations.
 * The reductions add +0.0, so the table is left unchanged. */
extern "C" int ngp_diag_l2_rate(void* table, uint32_t n_rows, uint32_t blocks, uint32_t rounds, int mode, void* sink, ngp_stream_t stream) {
    if (!table || !sink) return NGP_ERR_NULL;
    if (n_rows == 0 || (n_rows & (n_rows - 1)) || mode < 0 || mode > 2 || blocks == 0) return NGP_ERR_BAD_ARG;
    l2_rate_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>((uint32_t*)table, n_rows - 1, rounds, mode, (uint32_t*)sink);
    return finish_launch();
}
