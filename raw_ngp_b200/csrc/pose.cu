// pose.cu -- ray generation from refined camera poses and its backward: the step either side of
// march_rays_train when BARF pose refinement is on.
//
// Reference (per step, ~40 PyTorch kernels forward + their autograd):
//   CameraOptimizer.provide_refined_poses (barf/camera_optimizers.py:92-107):
//       pose_refine = lie.se3_to_SE3(se3_refine.weight[idx])          (barf/camera.py:93-105, Taylor series of order 10)
//       pose        = pose.compose([pose_refine, poses[idx, :3, :]])   (camera.py:47-63:  R = R_p R_r,  t = R_p t_r + t_p)
//   get_rays (nerf/train_utils.py:96-172):
//       rays_d = directions @ R^T  (camera-space pixel directions, NOT normalised),  rays_o = t
// Here: one thread per ray recomputes the refined pose of its camera (a few hundred flops) and writes the ray; the backward
// evaluates the same expression on forward-mode dual numbers, once per se3 component, and reduces the six directional
// derivatives into d se3[camera] with red.global.add.f32.  Nothing is saved between the two.
#include "common.cuh"

namespace ngp {
namespace {

struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }

template <typename T> __device__ __forceinline__ T lit(float c);
template <> __device__ __forceinline__ float lit<float>(float c) { return c; }
template <> __device__ __forceinline__ Dual lit<Dual>(float c) { return {c, 0.f}; }

// sum_{i=0..10} (-1)^i s^i / f_i with s = theta^2 (camera.py:124-153 evaluates x**(2i) of theta = |w|; only even powers
// occur, so the series is a polynomial in s = w.w and needs no square root -- which also makes the derivative at w = 0
// the finite value autograd produces there).  f_i: (2i+1)! for A = sin(x)/x, (2i+2)! for B = (1-cos x)/x^2, (2i+3)! for
// C = (x - sin x)/x^3.
template <typename T>
__device__ __forceinline__ void taylor_abc(T s, T& A, T& B, T& C) {
    A = lit<T>(0.f); B = lit<T>(0.f); C = lit<T>(0.f);
    T p = lit<T>(1.f);                   // (-s)^i
    float fa = 1.f, fb = 1.f, fc = 1.f;  // running factorial denominators
    for (int i = 0; i <= 10; i++) {
        if (i > 0) fa *= (float)((2 * i) * (2 * i + 1));
        fb *= (float)((2 * i + 1) * (2 * i + 2));
        fc *= (float)((2 * i + 2) * (2 * i + 3));
        A = A + p * lit<T>(1.f / fa);
        B = B + p * lit<T>(1.f / fb);
        C = C + p * lit<T>(1.f / fc);
        p = p * (-s);
    }
}

// Rt [3][4] = [R | V u] of se3 = (w, u)   (camera.py:93-105)
template <typename T>
__device__ __forceinline__ void se3_to_SE3(const T (&wu)[6], T (&Rt)[12]) {
    const T w0 = wu[0], w1 = wu[1], w2 = wu[2];
    T A, B, C;
    taylor_abc<T>(w0 * w0 + w1 * w1 + w2 * w2, A, B, C);
    const T O = lit<T>(0.f), I = lit<T>(1.f);
    const T wx[9] = {O, -w2, w1, w2, O, -w0, -w1, w0, O};
    T wx2[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) wx2[r * 3 + c] = wx[r * 3] * wx[c] + wx[r * 3 + 1] * wx[3 + c] + wx[r * 3 + 2] * wx[6 + c];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        T t = O;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const T eye = (r == c) ? I : O;
            Rt[r * 4 + c] = eye + A * wx[r * 3 + c] + B * wx2[r * 3 + c];
            t = t + (eye + B * wx[r * 3 + c] + C * wx2[r * 3 + c]) * wu[3 + c];
        }
        Rt[r * 4 + 3] = t;
    }
}

// refined ray of one pixel direction: R = R_p R_r, t = R_p t_r + t_p; rays_d = R dir, rays_o = t
template <typename T>
__device__ __forceinline__ void refined_ray(const T (&wu)[6], const float* __restrict__ P, const float (&dir)[3], T (&o)[3], T (&d)[3]) {
    T Rr[12];
    se3_to_SE3<T>(wu, Rr);
    T q[3];      // R_r dir
#pragma unroll
    for (int r = 0; r < 3; r++) q[r] = Rr[r * 4] * lit<T>(dir[0]) + Rr[r * 4 + 1] * lit<T>(dir[1]) + Rr[r * 4 + 2] * lit<T>(dir[2]);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const float p0 = __ldg(P + r * 4), p1 = __ldg(P + r * 4 + 1), p2 = __ldg(P + r * 4 + 2), p3 = __ldg(P + r * 4 + 3);
        d[r] = lit<T>(p0) * q[0] + lit<T>(p1) * q[1] + lit<T>(p2) * q[2];
        o[r] = lit<T>(p0) * Rr[3] + lit<T>(p1) * Rr[7] + lit<T>(p2) * Rr[11] + lit<T>(p3);
    }
}

__global__ void __launch_bounds__(128)
pose_rays_forward_kernel(const float* __restrict__ se3, const float* __restrict__ poses, uint32_t pose_stride,
                         const int* __restrict__ cam_idx, const float* __restrict__ dirs_cam, uint32_t N, float* __restrict__ rays_o,
                         float* __restrict__ rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t c = (uint32_t)__ldg(cam_idx + n);
    float wu[6];
#pragma unroll
    for (int k = 0; k < 6; k++) wu[k] = se3 ? __ldg(se3 + (size_t)c * 6 + k) : 0.f;
    const float dir[3] = {__ldg(dirs_cam + (size_t)n * 3), __ldg(dirs_cam + (size_t)n * 3 + 1), __ldg(dirs_cam + (size_t)n * 3 + 2)};
    float o[3], d[3];
    refined_ray<float>(wu, poses + (size_t)c * pose_stride, dir, o, d);
#pragma unroll
    for (int k = 0; k < 3; k++) { rays_o[(size_t)n * 3 + k] = o[k]; rays_d[(size_t)n * 3 + k] = d[k]; }
}

__global__ void __launch_bounds__(128)
pose_rays_backward_kernel(const float* __restrict__ d_rays_o, const float* __restrict__ d_rays_d, const float* __restrict__ se3,
                          const float* __restrict__ poses, uint32_t pose_stride, const int* __restrict__ cam_idx,
                          const float* __restrict__ dirs_cam, uint32_t N, float* __restrict__ d_se3) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t c = (uint32_t)__ldg(cam_idx + n);
    const float dir[3] = {__ldg(dirs_cam + (size_t)n * 3), __ldg(dirs_cam + (size_t)n * 3 + 1), __ldg(dirs_cam + (size_t)n * 3 + 2)};
    float go[3], gd[3];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        go[k] = __ldg(d_rays_o + (size_t)n * 3 + k);
        gd[k] = __ldg(d_rays_d + (size_t)n * 3 + k);
        any |= (go[k] != 0.f) || (gd[k] != 0.f);
    }
    if (!any) return;       // rays that hit nothing
    for (int j = 0; j < 6; j++) {
        Dual wu[6];
#pragma unroll
        for (int k = 0; k < 6; k++) wu[k] = {__ldg(se3 + (size_t)c * 6 + k), k == j ? 1.f : 0.f};
        Dual o[3], d[3];
        refined_ray<Dual>(wu, poses + (size_t)c * pose_stride, dir, o, d);
        const float g = go[0] * o[0].d + go[1] * o[1].d + go[2] * o[2].d + gd[0] * d[0].d + gd[1] * d[1].d + gd[2] * d[2].d;
        red_add_f32(d_se3 + (size_t)c * 6 + j, g);
    }
}

}  // namespace
}  // namespace ngp

using namespace ngp;

extern "C" int ngp_pose_rays_forward(const float* se3, const float* poses, uint32_t pose_stride, const int32_t* cam_idx,
                                     const float* dirs_cam, uint32_t N, uint32_t n_cameras, float* rays_o, float* rays_d,
                                     ngp_stream_t stream) {
    (void)n_cameras;
    if (N == 0) return NGP_OK;
    if (!poses || !cam_idx || !dirs_cam || !rays_o || !rays_d) return NGP_ERR_NULL;
    if (pose_stride < 12) return NGP_ERR_BAD_ARG;
    pose_rays_forward_kernel<<<div_up(N, 128u), 128, 0, (cudaStream_t)stream>>>(se3, poses, pose_stride, cam_idx, dirs_cam, N, rays_o, rays_d);
    return finish_launch();
}

extern "C" int ngp_pose_rays_backward(const float* d_rays_o, const float* d_rays_d, const float* se3, const float* poses,
                                      uint32_t pose_stride, const int32_t* cam_idx, const float* dirs_cam, uint32_t N,
                                      uint32_t n_cameras, float* d_se3, ngp_stream_t stream) {
    (void)n_cameras;
    if (N == 0) return NGP_OK;
    if (!d_rays_o || !d_rays_d || !se3 || !poses || !cam_idx || !dirs_cam || !d_se3) return NGP_ERR_NULL;
    if (pose_stride < 12) return NGP_ERR_BAD_ARG;
    pose_rays_backward_kernel<<<div_up(N, 128u), 128, 0, (cudaStream_t)stream>>>(d_rays_o, d_rays_d, se3, poses, pose_stride, cam_idx,
                                                                               dirs_cam, N, d_se3);
    return finish_launch();
}
