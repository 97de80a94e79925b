// common.cuh -- shared helpers for libngp_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/ngp_b200.h"

#ifndef __CUDA_ARCH__
#define NGP_HOST_ONLY 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libngp_b200 is written for sm_100a (Blackwell B200) only"
#endif

namespace ngp {

constexpr int kNumSMs = 148;  // B200

void set_last_cuda_error(cudaError_t e);

// Collects the launch status the reference never looked at (SURVEY 8b "Errors").
static inline int finish_launch() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_last_cuda_error(e);
        (void)cudaGetLastError();  // clear the sticky-less error so later calls are not poisoned
        return NGP_ERR_CUDA;
    }
    return NGP_OK;
}

template <typename T>
__host__ __device__ __forceinline__ T div_up(T a, T b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: the largest size configured so far is
// cached per device (a process that drives several GPUs from one thread must set it on each of them).
constexpr int kMaxDevices = 64;
struct SmemCache { uint32_t bytes[kMaxDevices]; };
template <typename F>
static inline int ensure_dynamic_smem(F kernel, uint32_t smem_bytes, SmemCache& cache) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_last_cuda_error(cudaGetLastError()); return NGP_ERR_CUDA; }
    const bool cached = dev >= 0 && dev < kMaxDevices;
    if (cached && smem_bytes <= cache.bytes[dev]) return NGP_OK;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
        set_last_cuda_error(cudaGetLastError());
        return NGP_ERR_CUDA;
    }
    if (cached) cache.bytes[dev] = smem_bytes;
    return NGP_OK;
}

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- scalar conversions -------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- C-wide vector load through the read-only path (ld.global.nc) -------------------------
// One instruction for C*sizeof(T) in {2,4,8,16} bytes, two for 32 bytes.
template <typename T, uint32_t C>
__device__ __forceinline__ void load_row(const T* __restrict__ p, float (&v)[C]) {
    constexpr uint32_t BYTES = sizeof(T) * C;
    if constexpr (BYTES == 2) {
        unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
        T t; *reinterpret_cast<unsigned short*>(&t) = u;
        v[0] = to_f32(t);
    } else if constexpr (BYTES == 4) {
        uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
        const T* t = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) v[c] = to_f32(t[c]);
    } else if constexpr (BYTES == 8) {
        uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        const T* t = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) v[c] = to_f32(t[c]);
    } else if constexpr (BYTES == 16) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const T* t = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) v[c] = to_f32(t[c]);
    } else {
        static_assert(BYTES == 32, "unsupported row width");
        uint4 u0 = __ldg(reinterpret_cast<const uint4*>(p));
        uint4 u1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        const T* t0 = reinterpret_cast<const T*>(&u0);
        const T* t1 = reinterpret_cast<const T*>(&u1);
#pragma unroll
        for (uint32_t c = 0; c < C / 2; c++) { v[c] = to_f32(t0[c]); v[c + C / 2] = to_f32(t1[c]); }
    }
}

// ---- C-wide vector store --------------------------------------------------------------------
template <typename T, uint32_t C>
__device__ __forceinline__ void store_row(T* __restrict__ p, const T (&v)[C]) {
    constexpr uint32_t BYTES = sizeof(T) * C;
    if constexpr (BYTES == 2) {
        p[0] = v[0];
    } else if constexpr (BYTES == 4) {
        *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(v);
    } else if constexpr (BYTES == 8) {
        *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(v);
    } else if constexpr (BYTES == 16) {
        *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
    } else {
        static_assert(BYTES == 32, "unsupported row width");
        reinterpret_cast<uint4*>(p)[0] = reinterpret_cast<const uint4*>(v)[0];
        reinterpret_cast<uint4*>(p)[1] = reinterpret_cast<const uint4*>(v)[1];
    }
}

// ---- packed / vector reductions to global memory (REDG.*, no return value) ----------------
__device__ __forceinline__ void red_add_f32(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void red_add_h2(void* p, uint32_t v) {
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2_h2(void* p, uint32_t a, uint32_t b) {
    asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_h2(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_bf2(void* p, uint32_t v) {
    asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2_bf2(void* p, uint32_t a, uint32_t b) {
    asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_bf2(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Adds the C-vector `g` (fp32 values, already weighted) to the table row at p with ONE reduction
// instruction where the hardware has a wide enough form (C*sizeof(T) <= 16 bytes), two for 32 bytes.
// C == 1 with a 16-bit table falls back to a scalar 16-bit red (the reference drops these
// gradients altogether, gridencoder.cu:22-26,341-346).
template <typename T, uint32_t C>
__device__ __forceinline__ void red_add_row(T* p, const float (&g)[C]) {
    if constexpr (sizeof(T) == 4) {
        float* f = reinterpret_cast<float*>(p);
        if constexpr (C == 1) red_add_f32(f, g[0]);
        else if constexpr (C == 2) red_add_v2_f32(f, g[0], g[1]);
        else if constexpr (C == 4) red_add_v4_f32(f, g[0], g[1], g[2], g[3]);
        else { red_add_v4_f32(f, g[0], g[1], g[2], g[3]); red_add_v4_f32(f + 4, g[4], g[5], g[6], g[7]); }
    } else if constexpr (sizeof(T) == 2 && C == 1) {
        if constexpr (std::is_same<T, __half>::value) {
            __half h = __float2half_rn(g[0]);
            asm volatile("red.global.add.noftz.f16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&h)) : "memory");
        } else {
            __nv_bfloat16 h = __float2bfloat16_rn(g[0]);
            asm volatile("red.global.add.noftz.bf16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&h)) : "memory");
        }
    } else if constexpr (std::is_same<T, __half>::value) {
        if constexpr (C == 2) red_add_h2(p, pack_h2(g[0], g[1]));
        else if constexpr (C == 4) red_add_v2_h2(p, pack_h2(g[0], g[1]), pack_h2(g[2], g[3]));
        else red_add_v4_h2(p, pack_h2(g[0], g[1]), pack_h2(g[2], g[3]), pack_h2(g[4], g[5]), pack_h2(g[6], g[7]));
    } else {
        if constexpr (C == 2) red_add_bf2(p, pack_bf2(g[0], g[1]));
        else if constexpr (C == 4) red_add_v2_bf2(p, pack_bf2(g[0], g[1]), pack_bf2(g[2], g[3]));
        else red_add_v4_bf2(p, pack_bf2(g[0], g[1]), pack_bf2(g[2], g[3]), pack_bf2(g[4], g[5]), pack_bf2(g[6], g[7]));
    }
}

}  // namespace ngp
