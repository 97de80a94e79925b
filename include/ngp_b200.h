/*
 * ngp_b200.h -- C ABI of libngp_b200.so: the sm_100a (B200) implementation of
 * raw_ngp's data-parallel NeRF hot path.
 *
 * Every entry point replaces one native function that the reference binds through
 * pybind11 (gridencoder/src/bindings.cpp:5-10, raymarching/src/bindings.cpp:5-20,
 * shencoder/src/bindings.cpp:5-8); the reference declaration it stands in for is cited
 * above each prototype as  <file>:<line>  relative to the reference tree.
 *
 * Conventions (differences from the reference's at::Tensor interface):
 *   - plain device pointers + sizes; the caller owns and allocates every buffer, the
 *     library allocates nothing and keeps no state (re-entrant, thread-safe);
 *   - every call takes the CUDA stream to launch on (the reference always used the
 *     legacy default stream) as an opaque pointer (cudaStream_t);
 *   - every call returns NGP_OK (0) or a negative NGP_ERR_* code and never throws;
 *     launch errors are collected with cudaPeekAtLastError() (the reference never checked);
 *   - `dtype` selects the element type of hash-table-typed buffers: NGP_F32 / NGP_F16 /
 *     NGP_BF16 (the reference dispatched float/double/half from the tensor; double is
 *     not provided, bf16 is new).  Ray-marching and SH buffers are always fp32: the
 *     reference wrappers cast them with custom_fwd(cast_inputs=float32)
 *     (raymarching/raymarching.py:34,254,336,399,450; shencoder/sphere_harmonics.py:16).
 *   - layouts are the *user-visible* layouts of the Python operators, so the wrappers do
 *     no permute copies: encoder outputs and their gradients are [B, L*C] (the reference
 *     kernel used [L,B,C] and the wrapper permuted, gridencoder/grid.py:49,63,81).
 */
#ifndef NGP_B200_H
#define NGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGP_B200_ABI_VERSION 4

typedef void* ngp_stream_t; /* cudaStream_t */

enum ngp_dtype { NGP_F32 = 0, NGP_F16 = 1, NGP_BF16 = 2 };

enum ngp_status {
    NGP_OK = 0,
    NGP_ERR_BAD_DTYPE = -1,   /* dtype not one of ngp_dtype                               */
    NGP_ERR_UNSUPPORTED = -2, /* D not in {2,3} / C not in {1,2,4,8} / degree not in 1..8  */
    NGP_ERR_NULL = -3,        /* required pointer is NULL                                 */
    NGP_ERR_ALIGN = -4,       /* pointer not aligned for the vector width of the kernel   */
    NGP_ERR_CUDA = -5,        /* cudaPeekAtLastError() != cudaSuccess after the launch    */
    NGP_ERR_BAD_ARG = -6      /* size / flag out of range                                 */
};

/* library identity */
int ngp_abi_version(void);
const char* ngp_status_string(int status);
/* last CUDA error string seen by this thread's most recent failing call ("" if none) */
const char* ngp_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Grid encoder  (reference: gridencoder/src/gridencoder.h:12-15)
 * ---------------------------------------------------------------------------------------- */

/* flags for ngp_grid_encode_forward / backward */
#define NGP_GRID_REF_ROUNDING 1u /* fp16 tables: accumulate in half exactly like the reference
                                    (gridencoder.cu:168,191); without it: fp32 accumulate, one rounding */
#define NGP_GRID_POINT_LEVEL_KERNELS 2u /* diagnostics: force the one-thread-per-(point, level) kernels (the reference's
                                           launch shape) instead of the tile kernels; results are identical */

/* replaces grid_encode_forward            gridencoder/src/gridencoder.h:12, gridencoder.cu:467-490
 * inputs      [B, D]       fp32 in [0,1]
 * embeddings  [sO, C]      dtype
 * offsets     [L+1]        int32
 * outputs     [B, L*C]     dtype   (levels >= max_level are written as zeros)
 * dy_dx       [B, L*D*C]   dtype or NULL  (same layout as the reference's dy_dx, gridencoder.cu:207)
 */
int ngp_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                            void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                            uint32_t max_level, float S, uint32_t H, void* dy_dx,
                            uint32_t gridtype, int align_corners, uint32_t interp,
                            int dtype, uint32_t flags, ngp_stream_t stream);

/* replaces grid_encode_backward           gridencoder/src/gridencoder.h:13, gridencoder.cu:492-522
 * grad             [B, L*C]  dtype
 * grad_embeddings  [sO, C]   dtype, must be zero-filled by the caller (as gridencoder/grid.py:83)
 * grad_inputs      [B, D]    fp32 or NULL; must be zero-filled by the caller.  It is recomputed
 *                            from the table (fuses kernel_input_backward, gridencoder.cu:352-378,
 *                            without the [B, L*D*C] dy_dx round trip).
 */
int ngp_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                             const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                             float* grad_inputs, uint32_t gridtype, int align_corners,
                             uint32_t interp, int dtype, uint32_t flags, ngp_stream_t stream);

/* reference-layout variant of the input gradient: grad_inputs[b,d] = sum_{l,c} grad[b,l,c]*dy_dx[b,l,d,c]
 * replaces kernel_input_backward          gridencoder.cu:352-378 (called from :421-441)
 * grad [B, L*C] dtype, dy_dx [B, L*D*C] dtype, grad_inputs [B, D] fp32 (overwritten) */
int ngp_grid_input_backward(const void* grad, const void* dy_dx, float* grad_inputs, uint32_t B,
                            uint32_t D, uint32_t C, uint32_t L, int dtype, ngp_stream_t stream);

/* replaces grad_total_variation           gridencoder/src/gridencoder.h:14, gridencoder.cu:525-668
 * inputs [B, D] dtype (the reference reads inputs as table dtype, gridencoder.cu:666), in [0,1];
 * grad [sO, C] dtype is accumulated in place. */
int ngp_grid_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                  const int32_t* offsets, float weight, uint32_t B, uint32_t D,
                                  uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                  int align_corners, int dtype, ngp_stream_t stream);

/* replaces grad_weight_decay              gridencoder/src/gridencoder.h:15, gridencoder.cu:670-713
 * n_entries = embeddings.shape[0] (the reference calls it B) */
int ngp_grid_grad_weight_decay(const void* embeddings, void* grad, const int32_t* offsets,
                               float weight, uint32_t n_entries, uint32_t C, uint32_t L, int dtype,
                               ngp_stream_t stream);

/* Per-level grid resolution exactly as the kernels evaluate it on the device, in fp32:
 * ceil(exp2f(level * S) * H)  (gridencoder.cu:133; the host computes table offsets in float64, grid.py:128,
 * and the two can disagree by one at the finest level).  out_dev uint32[L]. */
int ngp_grid_level_resolutions(uint32_t L, float S, uint32_t H, uint32_t* out_dev, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Spherical-harmonics encoder  (reference: shencoder/src/shencoder.h:9-10)
 * ---------------------------------------------------------------------------------------- */

/* replaces sh_encode_forward              shencoder/src/shencoder.h:9, shencoder.cu:400-417
 * inputs [B,3] fp32, outputs [B, degree^2] fp32, dy_dx [B, 3*degree^2] fp32 or NULL
 * out_dtype: NGP_F32 (reference behaviour) or NGP_F16/NGP_BF16 to emit the encoding directly in the
 * MLP's activation type (dy_dx stays fp32). */
int ngp_sh_encode_forward(const float* inputs, void* outputs, uint32_t B, uint32_t degree,
                          float* dy_dx, int out_dtype, ngp_stream_t stream);

/* replaces sh_encode_backward             shencoder/src/shencoder.h:10, shencoder.cu:419-439
 * grad [B, degree^2] (grad_dtype), inputs [B,3] fp32.  The derivative basis is recomputed from
 * `inputs` (no saved dy_dx); grad_inputs [B,3] fp32 is *accumulated into* like the reference
 * (shencoder.cu:377-379), so the caller zero-fills it (shencoder/sphere_harmonics.py:50). */
int ngp_sh_encode_backward(const void* grad, const float* inputs, uint32_t B, uint32_t degree,
                           float* grad_inputs, int grad_dtype, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Ray marching utilities  (reference: raymarching/src/raymarching.h:7-12)
 * ---------------------------------------------------------------------------------------- */

/* replaces near_far_from_aabb             raymarching/src/raymarching.h:7, raymarching.cu:91-156 */
int ngp_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                           float min_near, float* nears, float* fars, ngp_stream_t stream);

/* replaces sph_from_ray                   raymarching/src/raymarching.h:8, raymarching.cu:162-209 */
int ngp_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                     float* coords, ngp_stream_t stream);

/* replaces morton3D                       raymarching/src/raymarching.h:9, raymarching.cu:214-232 */
int ngp_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, ngp_stream_t stream);

/* replaces morton3D_invert                raymarching/src/raymarching.h:10, raymarching.cu:237-260 */
int ngp_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, ngp_stream_t stream);

/* replaces packbits                       raymarching/src/raymarching.h:11, raymarching.cu:267-300
 * grid [8*N] fp32, bitfield [N] uint8.  If thresh_dev != NULL the threshold is
 * min(*thresh_dev, density_thresh) read on the device (renderer.py:892 without the .item() sync). */
int ngp_packbits(const float* grid, uint32_t N, float density_thresh, const float* thresh_dev,
                 uint8_t* bitfield, ngp_stream_t stream);

/* replaces flatten_rays                   raymarching/src/raymarching.h:12, raymarching.cu:303-326 */
int ngp_flatten_rays(const int32_t* rays, uint32_t N, uint32_t M, int32_t* res, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training march / composite  (reference: raymarching/src/raymarching.h:14-16)
 * ---------------------------------------------------------------------------------------- */

/* Pass 1 of march_rays_train              raymarching/src/raymarching.h:14, raymarching.cu:337-508
 * (the reference calls its kernel with xyzs == nullptr, raymarching/raymarching.py:301).
 * Writes rays[n,1] = sample count of ray n, then rays[n,0] = exclusive prefix sum of the counts in
 * ray order (a legal outcome of the reference's atomicAdd(counter) and the only one its backward is
 * correct for, raymarching/raymarching.py:325), and counter[0] = M = total samples.
 * counter: int32[2] scratch, zero-filled by the caller ([1] is a block ticket).
 * t_scratch: fp32 [N * max_steps] workspace or NULL.  With it the march is warp-cooperative (one warp per ray,
 * 32 lattice points probed per instruction) and the t of every kept sample is left in t_scratch[n*max_steps + i]
 * for pass 2; without it the kernel is the one-thread-per-ray loop of the reference.  Same results either way. */
int ngp_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid,
                               float bound, int contract, float dt_gamma, uint32_t max_steps,
                               uint32_t N, uint32_t C, uint32_t H, const float* nears,
                               const float* fars, const float* noises, int32_t* rays,
                               int32_t* counter, float* t_scratch, ngp_stream_t stream);

/* Pass 2 of march_rays_train              raymarching.cu:337-508 with xyzs != nullptr
 * (raymarching/raymarching.py:311).  xyzs/dirs [M,3], ts [M,2], ldirs [M,3] or NULL (with rays_ldir).
 * t_scratch: the workspace filled by ngp_march_rays_train_count (samples are then written in parallel from their
 * stored t, no second march) or NULL (re-march like the reference; nears/fars/noises are only read then).
 * m_dev: NULL or device int32: rows available = min(M, *m_dev) (requires t_scratch). */
int ngp_march_rays_train_write(const float* rays_o, const float* rays_d, const float* rays_ldir,
                               const uint8_t* grid, float bound, int contract, float dt_gamma,
                               uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, const float* noises,
                               const int32_t* rays, uint32_t M, const int32_t* m_dev, const float* t_scratch,
                               float* xyzs, float* dirs, float* ts, float* ldirs, ngp_stream_t stream);

/* Pass 1 of march_rays_train with the renderer's differentiable near/far slab test (nerf/renderer.py:139-158,
 * torch semantics: (aabb - o) / (d + 1e-15), miss -> 1e9, near clamped to min_near) evaluated in the same kernel,
 * and a capacity: counter (int32[3], zero-filled once) receives [0] = total samples, [2] = samples of the longest
 * ray-ordered prefix whose rows fit in `cap` (what later stages use as the device-side M).  nears_out/fars_out
 * [N] or NULL.  t_scratch is required (warp-cooperative march).  No host synchronisation is needed between this
 * call and ngp_march_rays_train_write(..., M = cap, m_dev = counter + 2, ...). */
int ngp_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                    const uint8_t* grid, float bound, int contract, float dt_gamma,
                                    uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* noises,
                                    uint32_t cap, float* nears_out, float* fars_out, int32_t* rays,
                                    int32_t* counter, float* t_scratch, ngp_stream_t stream);

/* ngp_march_rays_train_count_aabb + the per-ray camera clip of run_cuda (nears = max(nears, cam_near_far[:, 0]), fars =
 * min(fars, cam_near_far[:, 1]); nerf/renderer.py:529-533; cam_near_far [N,2] or NULL) + a device-side live ray count for the
 * adaptive ray count of Trainer.train_step (nerf/train_utils.py:563-564; n_rays_dev int32[1] or NULL): rays at or beyond
 * *n_rays_dev get no samples. */
int ngp_march_rays_train_count_ex(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                  const float* cam_near_far, const int32_t* n_rays_dev, const uint8_t* grid, float bound,
                                  int contract, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                  const float* noises, uint32_t cap, float* nears_out, float* fars_out, int32_t* rays,
                                  int32_t* counter, float* t_scratch, ngp_stream_t stream);

/* n uniform [0, 1) floats from a counter-based generator (seed, *counter_dev, element index); *counter_dev advances by one per
 * call, on the device, so the call can sit inside a replayed CUDA graph and still produce a new stream every replay.  Stands in
 * for the `torch.rand(N)` of the marcher's per-ray jitter (raymarching/raymarching.py:295) and of the random background
 * (nerf/train_utils.py:496) inside the fused step; the operators themselves still take the noises as an input tensor. */
int ngp_uniform(float* out, uint32_t n, uint64_t seed, int32_t* counter_dev, ngp_stream_t stream);

/* Adaptive ray count on the device (nerf/train_utils.py:563-564, main.py:59-61): after a step that produced *m_dev samples from
 * *n_rays_dev rays, *n_rays_dev = clamp(round(target_points / m * n), 1, n_max) -- the reference computes the same on the
 * host after reading the sample count back. */
int ngp_adaptive_num_rays(int32_t* n_rays_dev, const int32_t* m_dev, uint32_t target_points, uint32_t n_max,
                          ngp_stream_t stream);

/* replaces composite_rays_train_forward   raymarching/src/raymarching.h:15, raymarching.cu:519-608
 * weights [M] must be zero-filled by the caller (raymarching/raymarching.py:356). */
int ngp_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ts,
                                     const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                     float* weights, float* weights_sum, float* depth, float* image,
                                     ngp_stream_t stream);

/* replaces composite_rays_train_backward  raymarching/src/raymarching.h:16, raymarching.cu:623-723
 * grad_sigmas [M], grad_rgbs [M,3] must be zero-filled by the caller (raymarching.py:382-383). */
int ngp_composite_rays_train_backward(const float* grad_weights, const float* grad_weights_sum,
                                      const float* grad_depth, const float* grad_image,
                                      const float* sigmas, const float* rgbs, const float* ts,
                                      const int32_t* rays, const float* weights_sum,
                                      const float* depth, const float* image, uint32_t M, uint32_t N,
                                      float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                      ngp_stream_t stream);

/* Optional inputs / outputs of ngp_composite_train_loss: the parts of Trainer.train_step (nerf/train_utils.py:481-568) around
 * the MSE / HDR loss that the reference evaluates with torch ops.  Every pointer may be NULL (feature off). */
typedef struct ngp_loss_opts {
    const float* bg_rays;        /* [N,3] per-ray background (background == 'random', train_utils.py:495-496); NULL: scalar bg_color */
    const float* target_alpha;   /* [N]   alpha of RGBA targets: gt = rgb * a + bg * (1 - a)   (train_utils.py:503-506) */
    const float* lossmult;       /* [N,3] Bayer mask of mosaiced raw data (train_utils.py:516-518) */
    const float* loss_weight;    /* [N,3] gaussian / planck / hanning weighting of the ground truth (train_utils.py:520-527) */
    const float* inv_norm_dev;   /* [1]   1 / sum(lossmult) on the device (train_utils.py:536); NULL: 1 / (3 * live rays) */
    const int32_t* n_rays_dev;   /* [1]   live rays of the batch, <= N (adaptive ray count, train_utils.py:563-564); NULL: N */
    float lambda_entropy;        /*       weight of the opacity entropy regulariser (train_utils.py:553-556); needs entropy_ray */
    float* entropy_ray;          /* [N]   scratch: per-ray entropy */
    float* weights_sum_out;      /* [N]   composited opacity (renderer results['weights_sum']) */
    float* depth_out;            /* [N]   composited depth (results['depth']) */
    float* parts_out;            /* [2]   data term and entropy term of loss_out */
    const float* loss_scale_dev; /* [1]   GradScaler scale kept on the device (ngp_grad_scaler_update); NULL: the loss_scale argument */
} ngp_loss_opts;

/* composite_rays_train forward + background blend + MSE loss + composite_rays_train backward for one training step
 * (raymarching.cu:519-723, nerf/renderer.py:553,672, nerf/train_utils.py:540-541) in one launch, one warp per ray:
 *   image = composite + (1 - weights_sum) * bg_color ;  loss = mean_n mean_c (image - target)^2
 *   grad_sigmas / grad_rgbs = d (loss_scale * loss) / d sigmas, rgbs   (every row of a ray that fits is written)
 * target [N,3]; image_out [N,3] or NULL; ray_loss [N] scratch; loss_out [1] (unscaled loss, summed in a fixed
 * order by the last block); ticket: int32[1] zero-filled once.  m_dev as in the field kernels.
 * loss_mode 0: the MSE above.  loss_mode 1: the raw/HDR loss of nerf/train_utils.py:529-536 -- c = min(1, image * exposure[n]),
 * loss = mean_n mean_c ((c - target) / (1e-3 + stop_grad(c)))^2; exposure [N] or NULL (= 1). */
int ngp_composite_train_mse(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                            uint32_t M, const int32_t* m_dev, uint32_t N, float T_thresh, float bg_color,
                            const float* target, float loss_scale, float* image_out, float* ray_loss,
                            float* loss_out, int32_t* ticket, float* grad_sigmas, float* grad_rgbs,
                            int loss_mode, const float* exposure, ngp_stream_t stream);

/* ngp_composite_train_mse with the optional terms of the ngp_loss_opts structure; `opts` may be NULL. */
int ngp_composite_train_loss(const float* sigmas, const float* rgbs, const float* ts, const int32_t* rays,
                             uint32_t M, const int32_t* m_dev, uint32_t N, float T_thresh, float bg_color,
                             const float* target, float loss_scale, float* image_out, float* ray_loss,
                             float* loss_out, int32_t* ticket, float* grad_sigmas, float* grad_rgbs, int loss_mode,
                             const float* exposure, const ngp_loss_opts* opts, ngp_stream_t stream);

/* Segmented sums of _march_rays_train.backward (raymarching/raymarching.py:319-329, which used
 * torch_scatter.segment_csr): dL/drays_o[n] = sum_seg dL/dxyz ; dL/drays_d[n] = sum_seg (dL/dxyz * t + dL/ddirs).
 * dL_ddirs may be NULL.  ts [M,2] (column 0 is used). */
int ngp_march_rays_train_backward(const float* dL_dxyzs, const float* dL_ddirs, const float* ts,
                                  const int32_t* rays, uint32_t N, uint32_t M, float* dL_drays_o,
                                  float* dL_drays_d, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Inference march / composite  (reference: raymarching/src/raymarching.h:18-19)
 * ---------------------------------------------------------------------------------------- */

/* replaces march_rays                     raymarching/src/raymarching.h:18, raymarching.cu:730-856
 * xyzs/dirs [n_alive*n_step,3], ts [n_alive*n_step,2]; the kernel zero-fills the unwritten tail of
 * every ray itself (the reference relied on torch.zeros, raymarching/raymarching.py:429-431). */
int ngp_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                   const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                   const float* nears, const float* fars, float* xyzs, float* dirs, float* ts,
                   const float* noises, ngp_stream_t stream);

/* replaces composite_rays                 raymarching/src/raymarching.h:19, raymarching.cu:859-950 */
int ngp_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                       float* rays_t, const float* sigmas, const float* rgbs, const float* ts,
                       float* weights_sum, float* depth, float* image, ngp_stream_t stream);

/* Device-side stream compaction of rays_alive (replaces `rays_alive[rays_alive >= 0]`,
 * nerf/renderer.py:612): writes the surviving ids, in order, to alive_out and their number to
 * n_out[0].  n_out int32[1].  workspace: NULL (one block does it all) or int32[ceil(n_alive / 4096)] scratch for the
 * two-pass multi-block version (a 1080p frame starts with 2 M ids). */
int ngp_compact_rays_alive(const int32_t* rays_alive, uint32_t n_alive, int32_t* alive_out,
                           int32_t* n_out, int32_t* workspace, ngp_stream_t stream);

/* The alive-ray loop of NeRFRenderer.run_cuda (nerf/renderer.py:588-616) without a host round trip per iteration: the reference
 * reads `rays_alive.shape[0]` after a boolean-mask compaction (a device -> host synchronisation) to size the next launches and to
 * pick n_step = max(min(N // n_alive, 8), 1).  Here that loop header lives in a device control block ctl = int32[4]
 * {n_alive, n_step, n_alive * n_step, step}; the three entry points below are ngp_march_rays / ngp_composite_rays /
 * ngp_compact_rays_alive reading their sizes from ctl, launched for `n_alive_bound` >= n_alive rays (any earlier value of the
 * count), and the compaction writes the control block of the NEXT iteration to ctl_next (n_alive = 0 once step >= max_steps;
 * n_step = max(min(row_cap // n_alive, step_cap), 1) with step_cap <= 16: row_cap = N, step_cap = 8 is the reference's schedule,
 * larger values render the same image -- the per-ray sample sequence does not depend on how it is cut into iterations -- in
 * fewer iterations; the sample buffers hold row_cap rows).
 * Field kernels take ctl + 2 as their m_dev.  Outputs are laid out exactly as by the host-driven entry points. */
int ngp_march_rays_dev(const int32_t* ctl, uint32_t n_alive_bound, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, int contract, float dt_gamma,
                       uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* fars, float* xyzs,
                       float* dirs, float* ts, const float* noises, ngp_stream_t stream);
int ngp_composite_rays_dev(const int32_t* ctl, uint32_t n_alive_bound, float T_thresh, int32_t* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* ts, float* weights_sum, float* depth,
                           float* image, ngp_stream_t stream);
int ngp_compact_rays_alive_dev(const int32_t* ctl, int32_t* ctl_next, uint32_t n_alive_bound, uint32_t row_cap,
                               uint32_t step_cap, uint32_t max_steps, const int32_t* rays_alive, int32_t* alive_out,
                               int32_t* n_out, int32_t* workspace, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Occupancy-grid update  (reference: nerf/renderer.py:811-897, Python loop over torch ops)
 * ---------------------------------------------------------------------------------------- */

/* Sample positions for one cascade of update_extra_state (renderer.py:824-848 full / :854-876 partial).
 * cell_indices == NULL : full update, cell n = n-th cell of the H^3 grid in x-major order (custom_meshgrid
 *                        order of renderer.py:835), indices_out[n] = morton3D(coords);
 * cell_indices != NULL : partial update, Morton indices given; coords = morton3D_invert(index).
 * noise [n,3] uniform [0,1).  xyzs_out[n,3] = (2c/(H-1)-1)*(bound-hgs) + (2*noise-1)*hgs, hgs = bound/H. */
int ngp_occ_sample_positions(const int32_t* cell_indices, const float* noise, uint32_t n, uint32_t H,
                             float bound, float* xyzs_out, int32_t* indices_out, ngp_stream_t stream);

/* EMA-max update of one cascade (renderer.py:848,883-885): for each i: j = indices[i];
 * tmp[j] = sigmas[i] (last writer wins on duplicates, as torch index_put);  then for every cell with
 * density_grid[j] >= 0 and tmp[j] >= 0: density_grid[j] = max(density_grid[j]*decay, tmp[j]).
 * tmp_grid [H3] scratch is filled with -1 by the caller; this call does the scatter. */
int ngp_occ_scatter_sigmas(const int32_t* indices, const float* sigmas, uint32_t n, float* tmp_grid,
                           ngp_stream_t stream);
/* density_grid/tmp_grid [n_cells]; mean_out[0] = mean(clamp(density_grid,min=0)) after the update
 * (renderer.py:887).  mean_out fp32[1]; accum fp64[2] scratch (running sum + block ticket), 16-byte aligned
 * and zero-filled by the caller. */
int ngp_occ_ema_update(float* density_grid, const float* tmp_grid, uint32_t n_cells, float decay,
                       double* accum, float* mean_out, ngp_stream_t stream);

/* The sample set of a PARTIAL occupancy update of one cascade (nerf/renderer.py:853-876: H^3/4 uniformly random cells + H^3/4
 * random occupied cells, jittered positions) without the reference's host sync on the size of the occupied list: the ids of the
 * cells with density > 0 are compacted from density_grid_cas [H^3] into occ_list [H^3] (count in occ_count[0]; workspace int32
 * [ceil(H^3 / 4096)]), then u [n, 6] uniforms in [0,1) (n even: first half uniform cells, second half occupied cells) become
 * xyzs_out [n, 3] and indices_out [n] (Morton ids).  Same distribution as the reference, different random stream. */
int ngp_occ_sample_partial(const float* density_grid_cas, uint32_t H, float bound, const float* u, uint32_t n,
                           int32_t* occ_list, int32_t* occ_count, int32_t* workspace, float* xyzs_out,
                           int32_t* indices_out, ngp_stream_t stream);

/* NeRFRenderer.mark_untrained_grid (nerf/renderer.py:716-809) as one launch: cell (cascade, Morton index) of density_grid
 * [cascade, H^3] is set to -1 unless its centre lies inside aabb [6] grown by half a cell AND inside the frustum of at least
 * one of the B cameras.  poses: camera-to-world, row-major [B, 3|4, 4] with `pose_stride` floats per camera; half_fov
 * [n_intr, 2] = (cx / fx, cy / fy) with n_intr = 1 (shared intrinsics) or B; cam_near [B] or NULL (then min_near for all);
 * grid_bound = NeRFRenderer.bound (cascade c spans min(2^c, grid_bound)). */
int ngp_mark_untrained_grid(float* density_grid, const float* poses, uint32_t pose_stride, uint32_t B,
                            const float* half_fov, uint32_t n_intr, const float* cam_near, float min_near,
                            const float* aabb, uint32_t H, uint32_t cascade, float grid_bound, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused MLP (reference: nerf/network.py:12-35, nn.Linear stack -> cuBLAS)
 * ---------------------------------------------------------------------------------------- */

#define NGP_MLP_MAX_LAYERS 4
enum ngp_activation { NGP_ACT_NONE = 0, NGP_ACT_RELU = 1 };

/* Bias-free MLP forward, all layers in one kernel, activations stay on chip.
 * x [M, dims[0]] f16 row-major (ldx elements per row), weights[l] [dims[l+1], dims[l]] f16 row-major
 * (nn.Linear layout), y [M, dims[n_layers]] f16.  hidden activation = `act` (network.py:30-34).
 * acts_out: optional [n_layers-1][M, dims[l+1]] f16 post-activation hidden states saved for backward
 * (pointer array on the host; entries may be NULL).  Dimensions must be multiples of 16 and <= 128 (callers
 * zero-pad, e.g. 31 -> 32 inputs, 3 -> 16 outputs).  Implementation: tcgen05.mma (M = 128 samples per CTA tile),
 * accumulators in TMEM, one persistent CTA loop over tiles. */
int ngp_mlp_forward(const void* x, uint32_t ldx, const void* const* weights, const uint32_t* dims,
                    uint32_t n_layers, uint32_t M, int act, void* y, uint32_t ldy,
                    void* const* acts_out, ngp_stream_t stream);

/* Backward: given dy [M, dims[n]] f16 and the saved hidden states, computes dx [M, dims[0]] f16 (or NULL)
 * and dW[l] [dims[l+1], dims[l]] fp32 (accumulated with atomics; zero-filled by the caller).  acts[l] is the
 * post-ReLU output of layer l as written by ngp_mlp_forward (row stride dims[l+1]). */
int ngp_mlp_backward(const void* dy, uint32_t lddy, const void* x, uint32_t ldx,
                     const void* const* weights, const void* const* acts, const uint32_t* dims,
                     uint32_t n_layers, uint32_t M, int act, void* dx, uint32_t lddx,
                     float* const* dweights, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused NeRF field (reference: nerf/network.py:74-143 -- GridEncoder -> grid_mlp -> trunc_exp / SHEncoder -> view_mlp ->
 * colour activation, each a separate PyTorch op there).  Four launches per training step:
 *   ngp_field_forward_density : xyz -> encode -> grid_mlp -> sigma, in2 = [feat(15), SH(dir), (SH(ldir)), 0]
 *   ngp_mlp_forward_rgb       : in2 -> view_mlp -> colour activation -> rgb
 *   ngp_mlp_backward_rgb      : d rgb -> view_mlp backward -> d in2
 *   ngp_field_backward_density: [d sigma, d in2[:, :15]] -> grid_mlp backward -> hash-table gradient (accumulated)
 * Table fp16 with F = 2, D = 3; MLP dims as ngp_mlp_forward (dims[0] = 2L, dims[n] = 16).
 * m_dev (all four): NULL, or a device int32 holding the live sample count; the kernels then process
 * min(M, *m_dev) rows, M being the capacity of the buffers.  This is what lets a whole training step run without the
 * reference's step_counter.item() host sync (raymarching/raymarching.py:303) and be captured in a CUDA graph.
 * ---------------------------------------------------------------------------------------- */

enum ngp_density_activation { NGP_DENSITY_EXP = 0, NGP_DENSITY_SOFTPLUS = 1 };           /* network.py:112-115 */
enum ngp_color_activation { NGP_COLOR_EXP = 1, NGP_COLOR_SIGMOID = 2, NGP_COLOR_CLAMPED_EXP = 3 }; /* network.py:131-138 */

/* xyzs [M,3] fp32 in [-bound, bound]; dirs [M,3] fp32 (any length, normalised inside like renderer.py:544 +
 * sphere_harmonics.py:81); ldirs [M,3] or NULL; feat_weights [2L] fp32 or NULL (BARF window, network.py:99-109).
 * Outputs: sigma_out [M] fp32; in2 [M, ld2] fp16 (ld2 >= 32, or 48 with ldirs); saved for backward: enc_out [M, 2L]
 * fp16 (or NULL) and acts_out[l] [M, dims[l+1]] fp16.  in2 == NULL computes the density only (NeRFNetwork.density). */
int ngp_field_forward_density(const float* xyzs, const float* dirs, const float* ldirs, const void* table,
                              const int32_t* offsets, const float* feat_weights, float bound, float S, uint32_t H,
                              uint32_t L, uint32_t gridtype, int align_corners, uint32_t interp,
                              const void* const* weights, const uint32_t* dims, uint32_t n_layers, uint32_t M,
                              const int32_t* m_dev, int density_act, float beta, void* enc_out, void* const* acts_out,
                              float* sigma_out, void* in2, uint32_t ld2, ngp_stream_t stream);

/* d_sigma, sigma [M] fp32; d_in2 [M, ld2] fp16 (columns 0..14 are used); enc / acts as saved by the forward;
 * grad_table [sO, 2] fp16 is ACCUMULATED into; dweights[l] fp32 accumulated with atomics.
 * Input gradients (BARF with the light-stage widths): dydx = the d enc / d x block that ngp_field_forward_full wrote
 * (dydx_out), d_xyzs [M, 3] fp32 out = d loss / d xyzs (kernel_input_backward, gridencoder.cu:352-378); both or neither. */
int ngp_field_backward_density(const float* xyzs, const float* d_sigma, const float* sigma, const void* d_in2,
                               uint32_t ld2, const void* enc, const void* dydx, const int32_t* offsets,
                               const float* feat_weights, float bound, float S, uint32_t H, uint32_t L,
                               uint32_t gridtype, int align_corners, uint32_t interp, const void* const* weights,
                               const void* const* acts, const uint32_t* dims, uint32_t n_layers, uint32_t M,
                               const int32_t* m_dev, int density_act, float beta, void* grad_table,
                               float* const* dweights, float* d_xyzs, ngp_stream_t stream);

/* d loss / d dirs [M, 3] from the gradient of the view_mlp input: d_in2 [M, ld2] fp16, columns [col0, col0 + 16) = d SH(dir)
 * (degree 4), through the SH Jacobian (shencoder.cu:130-350) and the two normalisations of the forward (renderer.py:544,
 * sphere_harmonics.py:81).  dirs = the marcher's directions.  The warp-specialised backward does this in its V0 group; this
 * entry point serves the kernel-pair backward. */
int ngp_sh_dirs_backward(const void* d_in2, uint32_t ld2, uint32_t col0, const float* dirs, uint32_t M,
                         const int32_t* m_dev, float* d_dirs, ngp_stream_t stream);

/* ngp_mlp_forward whose last epilogue applies the colour activation to output columns 0..2 and writes rgb_out [M,3] fp32 */
int ngp_mlp_forward_rgb(const void* x, uint32_t ldx, const void* const* weights, const uint32_t* dims,
                        uint32_t n_layers, uint32_t M, const int32_t* m_dev, int act, int color_act,
                        float* rgb_out, void* const* acts_out, ngp_stream_t stream);

/* ngp_mlp_backward whose incoming gradient is computed from d_rgb [M,3] fp32 and the activated rgb [M,3] fp32 */
int ngp_mlp_backward_rgb(const float* d_rgb, const float* rgb, int color_act, const void* x, uint32_t ldx,
                         const void* const* weights, const void* const* acts, const uint32_t* dims,
                         uint32_t n_layers, uint32_t M, const int32_t* m_dev, int act, void* dx, uint32_t lddx,
                         float* const* dweights, ngp_stream_t stream);

/* The whole field backward in ONE warp-specialised persistent kernel (csrc/field_bwd_ws.cu): two view groups (view_mlp
 * backward from d_rgb), two grid groups (grid_mlp backward from d_sigma and the view groups' d feat, handed over in shared
 * memory) and 16 scatter warps (hash-table gradient) run as a pipeline; saved tiles arrive by bulk async copies.
 * Inputs are what ngp_field_forward_full saved (tile-panel layout): enc, grid_acts[0..1], in2, view_acts[0..1]; plus
 * sigma [M], rgb [M,3] and the incoming d_sigma [M], d_rgb [M,3] (fp32).  grad_table [sO,2] fp16 and the fp32
 * grid_dweights[l] / view_dweights[l] ([dims[l+1], dims[l]]) are ACCUMULATED into.
 * Input gradients (BARF pose refinement: the positions and directions of the samples require grad, nerf/network.py with
 * rays_o / rays_d from refined poses): d_xyzs [M,3] and d_dirs [M,3] fp32, both NULL or both set, are WRITTEN (rows
 * < min(M, *m_dev)); they need dydx as written by ngp_field_forward_full (contracted with d enc like
 * kernel_input_backward, gridencoder.cu:352-378) and the un-normalised march directions dirs [M,3] (SH Jacobian + the
 * normalisations of renderer.py:544 and sphere_harmonics.py:81, shencoder.cu:358-382); view_dims[0] must be 32.
 * Otherwise pass NULLs. */
int ngp_field_backward_full(const float* xyzs, const float* d_sigma, const float* sigma, const float* d_rgb,
                            const float* rgb, const void* enc, const void* const* grid_acts, const void* in2,
                            const void* const* view_acts, const int32_t* offsets, const float* feat_weights,
                            float bound, float S, uint32_t H, uint32_t L, uint32_t gridtype, int align_corners,
                            uint32_t interp, const void* const* grid_weights, const uint32_t* grid_dims,
                            const void* const* view_weights, const uint32_t* view_dims, uint32_t M,
                            const int32_t* m_dev, int density_act, float beta, int color_act, void* grad_table,
                            float* const* grid_dweights, float* const* view_dweights, const void* dydx,
                            const float* dirs, float* d_xyzs, float* d_dirs, ngp_stream_t stream);

/* The whole field forward in ONE warp-specialised persistent kernel (csrc/field_ws.cu): gather warps encode into a ring of
 * shared-memory tiles while three groups of MLP warps run grid_mlp -> sigma / SH -> view_mlp -> colour on the tensor cores.
 * grid_dims = {2L, h, h, 16}, view_dims = {32 (48 with ldirs), h2, h2, 16}, all multiples of 16 and <= 128.
 * Outputs sigma_out [M], rgb_out [M,3] fp32.  Saved for the backward kernel (each may be NULL), all in the TILE-PANEL
 * layout: per 128-row tile the shared-memory image of the tile, i.e. [ceil(M/128)][128][width] fp16 row-major with the
 * 16-byte chunks of a row XOR-swizzled like the UMMA SWIZZLE_32B/64B/128B layouts (csrc/tile_sw.cuh); buffers hold
 * whole tiles; widths in {16, 32, 64} (else NGP_ERR_UNSUPPORTED: use the two-kernel path):
 * enc_out (width 2L), grid_acts_out[0..1] (h), in2_out (32/48), view_acts_out[0..1] (h2); ngp_field_backward_full
 * consumes exactly these.  view_weights == NULL: density-only query (grid_mlp only; dirs, ldirs,
 * view_dims, rgb_out and the view outputs are ignored) -- NeRFNetwork.density, the occupancy-grid update.
 * dydx_out (NULL, or ceil(M/128)*128 * 12L bytes, 16-byte aligned): d enc / d x in fp16 for the input gradients of the
 * backward kernel -- the dy_dx of gridencoder.cu:216-245 (calc_grad_inputs), stored per 128-row tile as
 * [4 level groups g][L/8 level pairs (g + 8p, g + 8p + 4)][2 levels x 3 dims][128 rows][2 channels]. */
int ngp_field_forward_full(const float* xyzs, const float* dirs, const float* ldirs, const void* table,
                           const int32_t* offsets, const float* feat_weights, float bound, float S, uint32_t H,
                           uint32_t L, uint32_t gridtype, int align_corners, uint32_t interp,
                           const void* const* grid_weights, const uint32_t* grid_dims,
                           const void* const* view_weights, const uint32_t* view_dims, uint32_t M,
                           const int32_t* m_dev, int density_act, float beta, int color_act, void* enc_out,
                           void* const* grid_acts_out, void* in2_out, void* const* view_acts_out,
                           float* sigma_out, float* rgb_out, void* dydx_out, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Frequency encoder (encoding.get_encoder('frequency'))
 * ---------------------------------------------------------------------------------------- */

/* replaces freq_encode_forward          freqencoder/src/freqencoder.h, freqencoder.cu:31-58,97-111
 * inputs [B, D] fp32 -> outputs [B, C] fp32, C = D + 2 D degree: the inputs, then per octave f < degree
 * __sinf(x 2^f) for all dims followed by __sinf(x 2^f + pi/2) for all dims. */
int ngp_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C, float* outputs,
                            ngp_stream_t stream);

/* replaces freq_encode_backward         freqencoder.cu:60-94,114-130
 * grad, outputs [B, C] (the forward's outputs carry the derivatives) -> grad_inputs [B, D] (written). */
int ngp_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C,
                             float* grad_inputs, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rays from refined camera poses (BARF pose refinement) -- SURVEY 8(f) row 2.
 * Replaces CameraOptimizer.provide_refined_poses (barf/camera_optimizers.py:92-107: lie.se3_to_SE3 of
 * barf/camera.py:93-153 composed with the dataset pose, camera.py:47-63) followed by get_rays
 * (nerf/train_utils.py:96-172) and the autograd of both.
 * ---------------------------------------------------------------------------------------- */

/* se3 [n_cameras, 6] fp32 (rotation w, translation u; NULL = identity refinement); poses: camera-to-world matrices, row
 * major, pose_stride floats apart (12 for [C,3,4], 16 for [C,4,4]; rows 0..2 are used); cam_idx [N] int32; dirs_cam [N,3]
 * camera-space pixel directions ((i+0.5-cx)/fx, -(j+0.5-cy)/fy, -1), not normalised.
 * rays_o[n] = R_p t_r + t_p, rays_d[n] = R_p R_r dirs_cam[n] with [R_r | t_r] = se3_to_SE3(se3[cam_idx[n]]). */
int ngp_pose_rays_forward(const float* se3, const float* poses, uint32_t pose_stride, const int32_t* cam_idx,
                          const float* dirs_cam, uint32_t N, uint32_t n_cameras, float* rays_o, float* rays_d,
                          ngp_stream_t stream);

/* d_se3 [n_cameras, 6] fp32 += sum over the rays of each camera of J^T [d_rays_o, d_rays_d] (ACCUMULATED with
 * reductions; the Jacobian is evaluated in forward mode from se3, nothing is saved by the forward). */
int ngp_pose_rays_backward(const float* d_rays_o, const float* d_rays_d, const float* se3, const float* poses,
                           uint32_t pose_stride, const int32_t* cam_idx, const float* dirs_cam, uint32_t N,
                           uint32_t n_cameras, float* d_se3, ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer over the flat parameter buffer (reference: torch.optim.Adam, main.py:245;
 * GradScaler unscale + inf check, nerf/train_utils.py:897-904) -- SURVEY 8(f) row 1.
 * ---------------------------------------------------------------------------------------- */

/* One Adam step (eps outside sqrt, no amsgrad, L2 weight_decay folded into grad like torch) over n
 * elements.  master/exp_avg/exp_avg_sq fp32; grad in grad_dtype, multiplied by *inv_scale_dev (device
 * fp32, e.g. 1/(loss_scale*world_size)); if *found_inf_dev != 0 the step is skipped (GradScaler semantics).
 * param_lp (optional, lp_dtype) receives the low-precision copy of the updated parameters; if
 * zero_grad != 0 the gradient buffer is cleared in the same pass.  step = 1-based step count. */
int ngp_fused_adam(float* master, void* param_lp, int lp_dtype, void* grad, int grad_dtype,
                   float* exp_avg, float* exp_avg_sq, uint64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, uint32_t step, const int32_t* step_dev, const float* lr_dev,
                   const float* inv_scale_dev, const float* found_inf_dev, int zero_grad, ngp_stream_t stream);

/* lr_dev (ngp_fused_adam): NULL, or a device fp32 that overrides `lr` -- the learning-rate schedule (LambdaLR,
 * main.py:258-261; ExponentialLR of the pose optimizer, barf/camera_optimizers.py:41-43) is then a 4-byte write between
 * replays of the captured step.
 * step_dev (ngp_fused_adam): NULL, or a device int32 holding the optimizer step count; the bias corrections are then
 * computed on the device from *step_dev (the host `step` is ignored), which lets the optimizer live inside a replayed
 * CUDA graph.  ngp_adam_step_counter increments *step_dev unless *found_inf_dev != 0 (a skipped GradScaler step is not
 * counted, as in torch). */
int ngp_adam_step_counter(int32_t* step_dev, const float* found_inf_dev, ngp_stream_t stream);

/* torch.amp.GradScaler.update() without a host round trip (nerf/train_utils.py:404, 897-904: the reference trains under a
 * dynamic GradScaler).  scale_dev [1], inv_scale_dev [1] = 1 / (scale * world) as the fused optimizers read it, state_dev
 * int32[2] = {clean steps since the last change, skipped steps so far}; found_inf_a / found_inf_b: the inf / nan flags of the
 * step (either may be NULL).  inf / nan: scale *= backoff_factor; growth_interval clean steps in a row: scale *= growth_factor. */
int ngp_grad_scaler_update(float* scale_dev, float* inv_scale_dev, int32_t* state_dev, const float* found_inf_a,
                           const float* found_inf_b, float growth_factor, float backoff_factor, int growth_interval,
                           uint32_t world, ngp_stream_t stream);

/* Data parallel over NVLink peer memory: reduce-scatter + Adam + all-gather in one kernel.  peer_grads[r] /
 * peer_params_lp[r] (host arrays of `world` device pointers, r = rank) are the SAME buffer on every rank -- gradient
 * (grad_dtype fp16 / fp32) and low-precision parameters (lp_dtype) -- mapped into this process (symmetric memory / CUDA
 * IPC).  This rank owns elements [lo, hi) (multiples of 8 for fp16 gradients, 4 for fp32): it sums that range of all
 * ranks' gradients in fp32, updates its fp32 master / exp_avg / exp_avg_sq SHARDS (hi - lo elements each) exactly like
 * ngp_fused_adam (step count and optional learning rate on the device, *inv_scale_dev, skip on *found_inf_dev) and writes
 * the updated parameters to peer_params_lp[0 .. n_store) (n_store = world: every rank's copy; n_store = 1 with
 * peer_params_lp[0] = the local copy and lo..hi = everything: replicated update of a small tensor).  Gradients are not
 * cleared (peers may still be reading them): the caller clears after its barrier.  The caller orders the ranks with
 * barriers on the stream before (all gradients and flags complete) and after (all parameter stores landed). */
int ngp_dp_fused_adam(const void* const* peer_grads, int grad_dtype, void* const* peer_params_lp, int lp_dtype,
                      uint32_t world, uint32_t n_store, float* master_shard, float* exp_avg_shard,
                      float* exp_avg_sq_shard, uint64_t lo, uint64_t hi, float lr, float beta1, float beta2, float eps,
                      float weight_decay, const int32_t* step_dev, const float* lr_dev, const float* inv_scale_dev,
                      const float* found_inf_dev, const float* flags, uint32_t n_flags,
                      ngp_stream_t stream);

/* The slim data-parallel update chain (4 launches + 2 barriers on the side stream instead of 8 + 2):
 *   ngp_dp_check_publish : ngp_check_finite_multi whose last block also stores this rank's inf / nan flag into slot `rank` of
 *                          every rank's flag array (peer stores; the caller's barrier orders them)
 *   ngp_dp_fused_adam    : with flags != NULL the ranks' flags are merged inside the kernel (any flag set: skip) and the update
 *                          uses step count *step_dev + 1, so no merge / counter launch precedes it
 *   ngp_dp_finish        : after the second barrier -- merged flag -> found_inf_dev, *step_dev += 1 unless skipped, and both
 *                          gradient buffers zero-filled (bytes multiples of 16) in one launch */
int ngp_dp_check_publish(const void* const* grads, const int* dtypes, const uint64_t* counts, uint32_t n_buffers,
                         float* found_inf_dev, uint32_t* scratch, void* const* peer_flags, uint32_t world, uint32_t rank,
                         ngp_stream_t stream);
int ngp_dp_finish(const float* flags, uint32_t world, float* found_inf_dev, int32_t* step_dev, void* grad0, uint64_t bytes0,
                  void* grad1, uint64_t bytes1, ngp_stream_t stream);

/* GradScaler flag across ranks without a collective: every rank stores its flag into slot `rank` of every rank's float[world]
 * array (peer_flags[r]); after the barrier ngp_dp_merge_flags takes the maximum of the local array. */
int ngp_dp_publish_flag(const float* found_inf_local, void* const* peer_flags, uint32_t world, uint32_t rank,
                        ngp_stream_t stream);
int ngp_dp_merge_flags(const float* flags, uint32_t world, float* found_inf, ngp_stream_t stream);

/* GradScaler + Adam step of a small fp32 tensor (n <= 2^20; the se3 pose corrections) in ONE single-block launch: inf / nan
 * check of grad, *step_dev += 1 unless skipped, unscale by *inv_scale_dev, Adam (torch semantics as ngp_fused_adam, learning
 * rate from *lr_dev if given), gradient cleared; found_inf_out (optional) receives 0 / 1. */
int ngp_small_adam(float* master, float* grad, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int32_t* step_dev, const float* lr_dev,
                   const float* inv_scale_dev, float* found_inf_out, ngp_stream_t stream);

/* ngp_small_adam for data-parallel ranks over peer memory: the gradient is the sum over peer_grads[0..world) (every rank's
 * buffer, mapped into this process; made complete by the caller's barrier), every rank applies the identical update to its own
 * replica.  The gradient buffers are NOT cleared (peers may still read them: ngp_dp_finish clears after the closing barrier).
 * Replaces the reference's DDP all-reduce of the pose optimizer's gradient (barf/camera_optimizers.py:40 under
 * nerf/train_utils.py:386, 900-904). */
int ngp_dp_small_adam(const void* const* peer_grads, uint32_t world, float* master, float* exp_avg, float* exp_avg_sq, uint32_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay, int32_t* step_dev, const float* lr_dev,
                      const float* inv_scale_dev, float* found_inf_out, ngp_stream_t stream);

/* The GradScaler bookkeeping of a step in one launch: found_inf_dev[0] = 1.0f if any element of any of the n_buffers (<= 4)
 * gradient buffers is inf / nan, else 0.0f (OVERWRITTEN, no zero-fill needed); *step_dev (optional) is incremented when
 * the step is not skipped (like ngp_adam_step_counter).  scratch: device uint32[2], zero before the first call; the
 * kernel leaves it zero. */
int ngp_check_finite_multi(const void* const* grads, const int* dtypes, const uint64_t* counts, uint32_t n_buffers,
                           float* found_inf_dev, int32_t* step_dev, uint32_t* scratch, ngp_stream_t stream);

/* found_inf_dev[0] = 1.0f if any element of grad is inf/nan (accumulates; caller zero-fills). */
int ngp_check_finite(const void* grad, int grad_dtype, uint64_t n, float* found_inf_dev,
                     ngp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics (tools/l2_ceiling.py): the random-row gather rate out of an L2-resident table and the random packed-fp16
 * reduction rate into one -- the ceilings of the forward's corner gathers and of the backward's scatter.  mode 0: 4-byte
 * gathers, 1: red.add.f16x2, 2: red.add.v2.f16x2 (aligned row pairs).  n_rows: power of two; every thread of
 * blocks x 512 does `rounds` x 8 row operations.  Adds +0.0: the table is unchanged.  Used by no operator.
 * ---------------------------------------------------------------------------------------- */
int ngp_diag_l2_rate(void* table, uint32_t n_rows, uint32_t blocks, uint32_t rounds, int mode, void* sink,
                     ngp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NGP_B200_H */
