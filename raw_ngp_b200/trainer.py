"""TrainStep -- one NeRF training step of the reference (nerf/train_utils.py:481-568, 863-930) on the B200 path:
march -> encode -> MLP -> composite -> loss -> backward -> (all-reduce) -> fused optimizer.

What is native here:
  * the hash table is held as fp32 master weights + a low-precision (fp16/bf16) working copy that the encoder
    reads; table gradients are accumulated by the backward kernel straight into one persistent buffer
    (GridEncoder.grad_sink): no zeros_like / memset per step, and the data-parallel all-reduce runs on that
    buffer without a pack copy;
  * GradScaler semantics (scale, unscale, skip on inf/nan, growth/backoff) with the check on the device;
  * Adam (eps=1e-15 like main.py:245) fused with the unscale, the low-precision copy and the gradient clear.
"""
import os

import torch
import torch.distributed as dist

from . import _lib, parallel


class _PeerMemory:
    """Symmetric (peer-mapped) device buffers of the ranks of one node: torch.distributed._symmetric_memory allocates and
    exchanges the handles; the kernels get plain arrays of device pointers, one per rank."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        self._sm, self.group, self.device = symm_mem, (group if group is not None else dist.group.WORLD), device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ptrs, self._handles = {}, {}

    def empty(self, n, dtype):
        return self._sm.empty(int(n), dtype=dtype, device=self.device)

    def empty_like(self, t, dtype):
        return self._sm.empty(*t.shape, dtype=dtype, device=self.device)

    def rendezvous(self, **tensors):
        import ctypes
        for name, t in tensors.items():
            h = self._sm.rendezvous(t, group=self.group)
            self._handles[name] = h
            self.ptrs[name] = (ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
        self._barrier_handle = self._handles[next(iter(tensors))]

    def barrier(self, channel):
        self._barrier_handle.barrier(channel=channel)      # a kernel on the current stream


class FusedAdam:
    """Adam over a list of (master fp32, low-precision copy or None, grad buffer) groups via ngp_fused_adam."""

    def __init__(self, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.0, device_step=None):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.lr_dev = None          # optional fp32[1] CUDA tensor overriding lr (a schedule under CUDA-graph replay)
        self.groups = []
        self.step_count = 0
        # optional int32[1] CUDA tensor: the step count lives on the device (bias corrections computed in the kernel), so that
        # the optimizer can be part of a replayed CUDA graph; incremented only for steps that are not skipped (GradScaler)
        self.device_step = device_step

    def add_group(self, master, grad, param_lp=None, lr_dev=None):
        assert master.dtype == torch.float32 and master.is_contiguous() and grad.is_contiguous()
        self.groups.append(dict(master=master, lp=param_lp, grad=grad, m=torch.zeros_like(master),
                                v=torch.zeros_like(master), lr_dev=lr_dev))

    def step(self, inv_scale, found_inf, zero_grad=True, lr=None, count_step=True):
        """count_step=False: the caller has already counted the step on the device (ngp_check_finite_multi does)."""
        self.step_count += 1
        _lib.weights_epoch += 1
        lr = self.lr if lr is None else lr
        st = _lib.stream()
        if self.device_step is not None and count_step:
            _lib.call("ngp_adam_step_counter", _lib.ptr(self.device_step), _lib.ptr(found_inf), st)
        for g in self.groups:
            lp = g["lp"]
            _lib.call("ngp_fused_adam", _lib.ptr(g["master"]), _lib.ptr(lp), _lib.dtype_id(lp.dtype) if lp is not None else 0,
                      _lib.ptr(g["grad"]), _lib.dtype_id(g["grad"].dtype), _lib.ptr(g["m"]), _lib.ptr(g["v"]),
                      g["master"].numel(), float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                      float(self.weight_decay), self.step_count, _lib.ptr(self.device_step),
                      _lib.ptr(g["lr_dev"] if g["lr_dev"] is not None else self.lr_dev), _lib.ptr(inv_scale),
                      _lib.ptr(found_inf), int(zero_grad), st)


class TrainStep:
    def __init__(self, model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, table_dtype=torch.float16, loss_scale=128.0,
                 dynamic_loss_scale=False, process_group=None, update_extra_interval=16, bg_color=1.0):
        self.model = model
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        dev = model.density_grid.device
        enc = model.grid_encoder

        # hash table: fp32 master + low-precision working copy (what the kernels read) + persistent gradient buffer
        self.table_master = enc.embeddings.data.float().contiguous()
        if table_dtype != torch.float32:
            enc.embeddings.data = self.table_master.to(table_dtype)
            table_lp = enc.embeddings.data
        else:
            enc.embeddings.data = self.table_master
            table_lp = None
        self.table_grad = torch.zeros_like(enc.embeddings.data)
        enc.grad_sink = self.table_grad

        # MLP weights: one flat fp32 buffer, parameters are views into it (so is their .grad)
        # the camera pose refinement has its own optimizer (barf/camera_optimizers.py:40: Adam(c_lr), default betas): not in this group
        mlp_params = [p for n, p in model.named_parameters() if not n.startswith(("grid_encoder.", "pose_optimizer."))]
        n_mlp = sum(p.numel() for p in mlp_params)
        self.mlp_flat = torch.empty(n_mlp, device=dev, dtype=torch.float32)
        self.mlp_grad = torch.zeros(n_mlp, device=dev, dtype=torch.float32)
        o = 0
        for p in mlp_params:
            n = p.numel()
            self.mlp_flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.mlp_flat[o:o + n].view_as(p)
            p.grad = self.mlp_grad[o:o + n].view_as(p)
            o += n
        self.mlp_params = mlp_params

        self.opt = FusedAdam(lr=lr, betas=betas, eps=eps)
        self.opt.add_group(self.table_master, self.table_grad, table_lp)
        self.opt.add_group(self.mlp_flat, self.mlp_grad, None)

        self.loss_scale = float(loss_scale) if table_dtype == torch.float16 or model.opt.fp16 else 1.0
        self.dynamic = dynamic_loss_scale
        self.growth_interval, self._good_steps = 2000, 0
        self.inv_scale = torch.empty(1, device=dev, dtype=torch.float32)
        self.found_inf = torch.zeros(1, device=dev, dtype=torch.float32)
        self.update_extra_interval = update_extra_interval
        self.bg_color = bg_color
        self.global_step = 0
        self.last_num_points = 0

    def _all_reduce_grads(self):
        parallel.all_reduce_gradients([self.table_grad, self.mlp_grad], None, self.pg)

    def step(self, rays_o, rays_d, target_rgb, rays_ldir=None, update_grid=True):
        """One optimisation step on N rays; returns the (unscaled) loss as a 0-d device tensor."""
        model = self.model
        if update_grid and self.global_step % self.update_extra_interval == 0:
            model.update_extra_state()
        model.train()
        out = model.render(rays_o, rays_d, rays_ldir=rays_ldir, bg_color=self.bg_color, perturb=True)
        self.last_num_points = out["num_points"]
        loss = torch.nn.functional.mse_loss(out["image"], target_rgb, reduction="none").mean(-1).mean()
        (loss * self.loss_scale).backward()

        self._all_reduce_grads()
        st = _lib.stream()
        self.found_inf.zero_()
        _lib.call("ngp_check_finite", _lib.ptr(self.table_grad), _lib.dtype_id(self.table_grad.dtype),
                  self.table_grad.numel(), _lib.ptr(self.found_inf), st)
        _lib.call("ngp_check_finite", _lib.ptr(self.mlp_grad), _lib.NGP_F32, self.mlp_grad.numel(), _lib.ptr(self.found_inf), st)
        if self.world > 1:
            dist.all_reduce(self.found_inf, op=dist.ReduceOp.MAX, group=self.pg)
        self.inv_scale.fill_(parallel.unscale_factor(self.loss_scale, self.world))
        self.opt.step(self.inv_scale, self.found_inf, zero_grad=True)
        if self.dynamic:  # GradScaler growth / backoff (one sync, only in dynamic mode)
            if self.found_inf.item() != 0:
                self.loss_scale *= 0.5
                self._good_steps = 0
            else:
                self._good_steps += 1
                if self._good_steps % self.growth_interval == 0:
                    self.loss_scale *= 2.0
        self.global_step += 1
        return loss.detach()


class FusedTrainStep:
    """The same training step as TrainStep with nothing left on the host: ~10 kernels, no synchronisation, replayed as
    a CUDA graph.

    What the reference does per step and where it went:
      * near_far_from_aabb (renderer.py:139-158, ~10 torch kernels)      -> inside the marching kernel
      * march pass 1, step_counter.item() HOST SYNC, allocation, pass 2   -> the sample count stays on the device
        (raymarching.py:301-311)                                             (m_dev), buffers have a fixed capacity
      * dirs normalisation, encoder, MLPs, activations (~60 kernels)      -> ngp_field_forward_full
      * composite fwd, bg blend, MSE, autograd of all that, composite bwd -> ngp_composite_train_mse
      * MLP / encoder backward incl. zeros_like(embeddings)               -> ngp_field_backward_full
      * GradScaler.unscale_/inf check/Adam/zero_grad                      -> ngp_check_finite + ngp_fused_adam
    Numerically it is the autograd path of TrainStep (same kernels for march / field / MLP / Adam, same loss); the
    equality is tested in tests/test_gpu_trainstep.py.  Eligibility = the fused-field conditions of NeRFNetwork plus an
    MSE loss and a scalar background."""

    def __init__(self, model, n_rays, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, loss_scale=128.0, max_samples=None,
                 process_group=None, update_extra_interval=16, bg_color=1.0, perturb=True, use_graph=True, loss="mse",
                 ray_grads=False, pose_optimizer=None, poses=None, pose_lr=1e-3, pose_betas=(0.9, 0.999), pose_eps=1e-8,
                 rgba_targets=False, lossmult=False, loss_weight=False, cam_near_far=False, lambda_entropy=0.0,
                 adaptive_num_rays=False, num_points=2 ** 18, dynamic_loss_scale=True, growth_interval=2000,
                 allow_nccl_fallback=False):
        """The options after pose_eps switch on the remaining pieces of Trainer.train_step (nerf/train_utils.py:481-568), each a
        static device buffer read by the captured kernels (fill them through set_rays):
          bg_color="random"   per-ray random background, redrawn on the device every step (:495-496)
          rgba_targets        targets carry alpha: gt = rgb * a + bg * (1 - a) (:503-506)
          lossmult, loss_weight  [N, 3] Bayer mask / ground-truth weighting of the HDR loss, normalised by sum(lossmult) (:515-536)
          cam_near_far        [N, 2] per-ray camera clip of near / far (renderer.py:529-533)
          lambda_entropy      opacity entropy regulariser (:553-556)
          adaptive_num_rays   the number of live rays of the next batch follows num_points / samples of this one (:563-564);
                              n_rays is then the capacity, the live count stays on the device (self.n_rays_dev)
          dynamic_loss_scale  torch.amp.GradScaler semantics on the device (train_utils.py:404, 897-904): the scale starts at
                              loss_scale, halves after a step with inf / nan gradients (that step is skipped), doubles after
                              growth_interval clean steps; self.skipped_steps() reads the device-side count of skipped steps"""
        import ctypes
        from . import field as _field
        from .ffmlp import _pad16
        self._ct = ctypes
        self.model, self.N = model, int(n_rays)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        enc = model.grid_encoder
        dev = model.density_grid.device
        self.dev = dev
        opt = model.opt
        if not (enc.level_dim == 2 and enc.input_dim == 3 and enc.num_levels % 4 == 0 and opt.internal_activation == "relu"
                and opt.density_activation in ("clamped_exp", "softplus") and opt.color_activation in _field.COLOR_ACT
                and opt.pose_opt in ("none", "barf") and model.grid_mlp.num_layers == 3 and model.view_mlp.num_layers == 3):
            raise RuntimeError("FusedTrainStep: configuration outside the fused field path; use TrainStep")
        self.rfield = bool(opt.rfield)
        if loss not in ("mse", "hdr"):
            raise ValueError("loss must be 'mse' or 'hdr' (nerf/train_utils.py:512-541)")
        self.loss_mode = 0 if loss == "mse" else 1
        self.perturb = perturb
        self.random_bg = isinstance(bg_color, str)
        if self.random_bg and bg_color != "random":
            raise ValueError("bg_color must be a number or 'random'")
        self.bg_color = 0.0 if self.random_bg else float(bg_color)
        self.lambda_entropy = float(lambda_entropy)
        self.adaptive, self.num_points = bool(adaptive_num_rays), int(num_points)
        self.loss_scale = float(loss_scale)
        self.update_extra_interval = update_extra_interval
        self.global_step = 0
        N = self.N
        self.cap = int(max_samples) if max_samples else N * int(opt.max_steps)
        cap_t = (self.cap + 127) // 128 * 128      # saved activations are stored as whole 128-row tiles (tile-panel layout)

        # ---- parameters: fp32 master + fp16 working copy + fp32/fp16 gradient buffers -------------------------------
        self.table_master = enc.embeddings.data.float().contiguous()
        # Data parallel over NVLink peer memory (csrc/optim.cu: dp_fused_adam_kernel): the gradient buffers and the fp16 table
        # live in symmetric memory, every rank owns a shard of the optimizer state.  Needs CUDA + NCCL ranks on one node;
        # NGP_DP_PEER=0 (or a failing rendezvous) selects the NCCL all-reduce + replicated Adam path.
        self.peer = None
        self.d1 = [model.grid_mlp.net[0].weight.shape[1]] + [l.weight.shape[0] for l in model.grid_mlp.net]
        self.d2 = [model.view_mlp.net[0].weight.shape[1]] + [l.weight.shape[0] for l in model.view_mlp.net]
        self.p1, self.p2 = [_pad16(d) for d in self.d1], [_pad16(d) for d in self.d2]
        shapes = [(self.p1[i + 1], self.p1[i]) for i in range(3)] + [(self.p2[i + 1], self.p2[i]) for i in range(3)]
        n_w = sum(a * b for a, b in shapes)
        if self.world > 1 and dev.type == "cuda" and self.world <= 8 and os.environ.get("NGP_DP_PEER", "1") != "0" \
                and dist.get_backend(process_group) == "nccl":
            try:
                pm = _PeerMemory(process_group, dev)
                table_lp = pm.empty_like(self.table_master, torch.float16).copy_(self.table_master)
                table_grad = pm.empty_like(self.table_master, torch.float16).zero_()
                # the se3 gradient of the pose optimizer rides behind the MLP gradients in the same symmetric buffer (one peer
                # mapping, cleared together by ngp_dp_finish)
                n_se3 = pose_optimizer.se3_refine.weight.numel() if pose_optimizer is not None else 0
                w_grad_all = pm.empty(n_w + (n_se3 + 3) // 4 * 4, torch.float32).zero_()
                w_grad = w_grad_all[:n_w]
                flags = pm.empty(8, torch.float32).zero_()
                pm.rendezvous(grad=table_grad, table=table_lp, w_grad=w_grad_all, flags=flags)
                self.peer = pm
            except Exception as e:       # no symmetric memory on this system (all ranks fail alike)
                if not allow_nccl_fallback:
                    raise RuntimeError("FusedTrainStep: the peer-memory data-parallel path (symmetric memory over NVLink) is unavailable: "
                                       f"{e!r}.  The NCCL all-reduce + replicated Adam path is ~15 % slower per step; select it explicitly "
                                       "with NGP_DP_PEER=0 or allow_nccl_fallback=True.") from e
                import warnings
                warnings.warn(f"FusedTrainStep: peer-memory data parallel path unavailable ({e!r}); using NCCL all-reduce")
        if self.peer is not None:
            enc.embeddings.data, self.table_grad, self.w_grad, self.flags = table_lp, table_grad, w_grad, flags
            self._w_grad_all, self._n_w = w_grad_all, n_w
        else:
            enc.embeddings.data = self.table_master.half()
            self.table_grad = torch.zeros_like(enc.embeddings.data)
            self.w_grad = torch.zeros(n_w, device=dev, dtype=torch.float32)
        enc.grad_sink = self.table_grad
        layers = [l for l in model.grid_mlp.net] + [l for l in model.view_mlp.net]
        # widths in {16, 32, 64}: the warp-specialised one-kernel forward / backward; otherwise (rfield: 48-wide view input,
        # 80-wide hidden layers) the density-field + view-MLP kernel pairs, with the same buffers and the same graph
        self.ws = _field._ws_ok(self.p1, self.p2)
        # any width that is a multiple of 16: the warp-specialised FORWARD still applies (row-major saves for the pair backward)
        self.ws_fwd = self.ws or _field._ws_fwd_ok(self.p1, self.p2, enc.num_levels)
        self.ray_grads = bool(ray_grads) or pose_optimizer is not None
        # ray_grads: also produce dL/d rays_o and dL/d rays_d (self.d_rays_o / self.d_rays_d, scaled by loss_scale like every
        # gradient of the step) for pose refinement (BARF, --pose_opt barf: rays come from refined poses and require grad)
        # pose_optimizer (raw_ngp_b200.pose.CameraOptimizer) + poses [C, 3|4, 4]: the rays of the step are generated on the
        # device from (camera index, pixel direction) through the refined poses, and se3_refine is trained by its own Adam
        # (torch defaults like barf/camera_optimizers.py:40) with its own inf check (GradScaler checks per optimizer)
        self.pose = pose_optimizer
        if self.ray_grads and not self.ws_fwd:
            raise RuntimeError("FusedTrainStep: ray gradients need the warp-specialised forward (layer widths multiples of 16 up to 128)")
        self.w_master = torch.zeros(n_w, device=dev, dtype=torch.float32)
        self.w_lp = torch.zeros(n_w, device=dev, dtype=torch.float16)
        self._w_master_views, self._w_lp_views, self._w_grad_views = [], [], []
        o = 0
        for lin, (n, k) in zip(layers, shapes):
            mv = self.w_master[o:o + n * k].view(n, k)
            mv[:lin.weight.shape[0], :lin.weight.shape[1]].copy_(lin.weight.data)
            lin.weight.data = mv[:lin.weight.shape[0], :lin.weight.shape[1]]      # the module keeps seeing the trained weights
            self._w_master_views.append(mv)
            self._w_lp_views.append(self.w_lp[o:o + n * k].view(n, k))
            self._w_grad_views.append(self.w_grad[o:o + n * k].view(n, k))
            o += n * k
        self.w_lp.copy_(self.w_master)
        self._w_lp_local = (ctypes.c_void_p * 1)(self.w_lp.data_ptr())

        self.opt_step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.opt = FusedAdam(lr=lr, betas=betas, eps=eps, device_step=self.opt_step_dev)
        self.lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)      # see set_lr()
        self.opt.lr_dev = self.lr_dev
        if self.peer is None:
            self.opt.add_group(self.table_master, self.table_grad, enc.embeddings.data)
            self.opt.add_group(self.w_master, self.w_grad, self.w_lp)
        else:
            # shard [lo, hi) of the flattened table: fp32 master + Adam moments exist for the shard only
            n = self.table_master.numel()
            per = (n + 8 * self.world - 1) // (8 * self.world) * 8
            rank = dist.get_rank(process_group)
            self.shard = (min(rank * per, n), min((rank + 1) * per, n))
            lo, hi = self.shard
            self.table_master_shard = self.table_master.view(-1)[lo:hi].clone()
            self.table_master = None                         # see gather_table_master()
            self.shard_m, self.shard_v = torch.zeros_like(self.table_master_shard), torch.zeros_like(self.table_master_shard)
            self.w_m, self.w_v = torch.zeros_like(self.w_master), torch.zeros_like(self.w_master)
        self._pending = False          # gradients of the last step are waiting for their optimizer update
        if self.pose is not None:
            if poses is None:
                raise ValueError("FusedTrainStep: pose_optimizer needs the dataset poses [C, 3|4, 4]")
            self.poses = poses.to(dev).float().contiguous()
            self.pose_stride = self.poses.shape[-2] * self.poses.shape[-1]
            w = self.pose.se3_refine.weight
            w.data = w.data.to(dev).float().contiguous()
            self.se3 = w.data
            if self.se3.shape[0] != self.poses.shape[0]:
                raise ValueError("FusedTrainStep: one se3 row per dataset pose")
            if self.peer is not None:
                # summed across ranks by ngp_dp_small_adam through the peer mappings (no NCCL call in front of the ray generation)
                self.se3_grad = self._w_grad_all[self._n_w:self._n_w + self.se3.numel()].view_as(self.se3)
                self._se3_peer_ptrs = (ctypes.c_void_p * self.world)(*[int(p) + 4 * self._n_w for p in self.peer.ptrs["w_grad"]])
            else:
                self.se3_grad = torch.zeros_like(self.se3)
            self.cam_idx = torch.zeros(N, device=dev, dtype=torch.int32)
            self.dirs_cam = torch.zeros(N, 3, device=dev, dtype=torch.float32)
            self.pose_found_inf = torch.zeros(1, device=dev, dtype=torch.float32)
            self.pose_lr_dev = torch.full((1,), float(pose_lr), device=dev, dtype=torch.float32)
            self.pose_step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
            self.pose_opt = FusedAdam(lr=pose_lr, betas=pose_betas, eps=pose_eps, device_step=self.pose_step_dev)
            self.pose_opt.add_group(self.se3, self.se3_grad, None, lr_dev=self.pose_lr_dev)
        self.inv_scale = torch.full((1,), parallel.unscale_factor(self.loss_scale, self.world), device=dev, dtype=torch.float32)
        self.found_inf = torch.zeros(1, device=dev, dtype=torch.float32)
        self.dynamic = bool(dynamic_loss_scale)
        self.growth_interval = int(growth_interval)
        self.scale_dev = torch.full((1,), self.loss_scale, device=dev, dtype=torch.float32)
        self.scaler_state = torch.zeros(2, device=dev, dtype=torch.int32)      # clean steps since the last change, skipped steps

        # ---- static buffers --------------------------------------------------------------------------------------------
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        cap = self.cap
        self.rays_o, self.rays_d, self.target = torch.zeros(N, 3, **f32), torch.zeros(N, 3, **f32), torch.zeros(N, 3, **f32)
        self.rays_ldir = torch.zeros(N, 3, **f32) if self.rfield else None
        self.noises = torch.zeros(N, **f32)
        self._rng_seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF      # follows torch.manual_seed
        self._rng_counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self.exposure = torch.ones(N, **f32)          # per-ray exposure of the HDR loss (train_utils.py:514)
        self.bg_rays = torch.zeros(N, 3, **f32) if self.random_bg else None
        self.target_alpha = torch.ones(N, **f32) if rgba_targets else None
        self.lossmult = torch.ones(N, 3, **f32) if lossmult else None
        self.loss_weight = torch.ones(N, 3, **f32) if loss_weight else None
        self.inv_norm = torch.full((1,), 1.0 / (3 * N), **f32) if lossmult else None
        self.cam_near_far = torch.tensor([0.0, 1e9], **f32).repeat(N, 1).contiguous() if cam_near_far else None
        self.n_rays_dev = torch.full((1,), N, device=dev, dtype=torch.int32) if self.adaptive else None
        self.entropy_ray = torch.zeros(N, **f32) if self.lambda_entropy > 0 else None
        self.weights_sum, self.depth = torch.zeros(N, **f32), torch.zeros(N, **f32)
        self.loss_parts = torch.zeros(2, **f32)
        P0 = lambda t: None if t is None else t.data_ptr()
        self._loss_opts = _lib.LossOpts(P0(self.bg_rays), P0(self.target_alpha), P0(self.lossmult), P0(self.loss_weight), P0(self.inv_norm),
                                        P0(self.n_rays_dev), self.lambda_entropy, P0(self.entropy_ray), P0(self.weights_sum),
                                        P0(self.depth), P0(self.loss_parts), P0(self.scale_dev) if self.dynamic else None)
        self.rays = torch.zeros(N, 2, device=dev, dtype=torch.int32)
        self.counter = torch.zeros(4, device=dev, dtype=torch.int32)
        self.ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        self.t_scratch = torch.empty(N * int(opt.max_steps), **f32)
        self.xyzs, self.dirs, self.ts = torch.empty(cap, 3, **f32), torch.empty(cap, 3, **f32), torch.empty(cap, 2, **f32)
        self.ldirs = torch.empty(cap, 3, **f32) if self.rfield else None
        self.enc_buf = torch.empty(cap_t, self.p1[0], **f16)
        self.acts1 = [torch.empty(cap_t, self.p1[l + 1], **f16) for l in range(2)]
        self.acts2 = [torch.empty(cap_t, self.p2[l + 1], **f16) for l in range(2)]
        self.in2 = torch.empty(cap_t, self.p2[0], **f16)
        self.d_in2 = None if self.ws else torch.empty(cap_t, self.p2[0], **f16)
        self.sigma, self.rgb = torch.empty(cap, **f32), torch.empty(cap, 3, **f32)
        self.d_sigma, self.d_rgb = torch.empty(cap, **f32), torch.empty(cap, 3, **f32)
        self.dydx = torch.empty(cap_t, 6 * enc.num_levels, **f16) if self.ray_grads else None   # d enc / d x, saved by the forward
        self.d_xyzs = torch.empty(cap, 3, **f32) if self.ray_grads else None
        self.d_dirs = torch.empty(cap, 3, **f32) if self.ray_grads else None
        self.d_rays_o = torch.zeros(N, 3, **f32) if self.ray_grads else None
        self.d_rays_d = torch.zeros(N, 3, **f32) if self.ray_grads else None
        self.image, self.ray_loss, self.loss = torch.zeros(N, 3, **f32), torch.zeros(N, **f32), torch.zeros(1, **f32)
        self.feat_weights = torch.ones(2 * enc.num_levels, **f32) if opt.pose_opt == "barf" else None
        self._feat_host = torch.ones(2 * enc.num_levels, dtype=torch.float32).pin_memory() if self.feat_weights is not None else None
        self._feat_annealing = None
        self._density_act = model._density_act()
        self._color_act = _field.COLOR_ACT[opt.color_activation]
        self._grid_scalars = _field._grid_scalars(enc)
        self.use_graph = use_graph
        self._graph = None
        self.graph_kernels = self.pipe_kernels = self.kernels_replayed = 0
        self._m_dev = self.counter.data_ptr() + 8   # counter[2]: samples of the rays that fit in `cap`

    # ------------------------------------------------------------------------------------------------------------------
    def _ptrs(self, tensors):
        arr = (self._ct.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = t.data_ptr()
        return arr

    def _launch_forward_backward(self, join=None):
        """march -> field -> composite+loss -> backward, all on the current stream (graph-capturable).  `join`: a stream whose
        work (the previous step's optimizer update) must finish before the field kernels read the weights."""
        self._launch_march()
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)
        self._launch_field()

    def _launch_pose_update(self):
        """inf check + Adam of se3_refine for the PREVIOUS step's gradients: on the main stream, because the rays of this step
        are generated from the updated poses (the table / MLP update runs beside the march on the side stream).
        Peer-memory data parallel: the gradient is summed over the ranks' buffers inside the kernel; the caller has put a
        barrier in front (every rank's backward has finished) and ngp_dp_finish clears the buffers afterwards."""
        g = self.pose_opt.groups[0]
        b1, b2 = self.pose_opt.betas
        self.pose_opt.step_count += 1
        _lib.weights_epoch += 1
        if self.peer is not None:
            _lib.call("ngp_dp_small_adam", self._se3_peer_ptrs, self.world, _lib.ptr(self.se3), _lib.ptr(g["m"]), _lib.ptr(g["v"]),
                      self.se3.numel(), float(self.pose_opt.lr), float(b1), float(b2), float(self.pose_opt.eps),
                      float(self.pose_opt.weight_decay), _lib.ptr(self.pose_step_dev), _lib.ptr(self.pose_lr_dev), _lib.ptr(self.inv_scale),
                      _lib.ptr(self.pose_found_inf), _lib.stream())
            return
        _lib.call("ngp_small_adam", _lib.ptr(self.se3), _lib.ptr(self.se3_grad), _lib.ptr(g["m"]), _lib.ptr(g["v"]), self.se3.numel(),
                  float(self.pose_opt.lr), float(b1), float(b2), float(self.pose_opt.eps), float(self.pose_opt.weight_decay),
                  _lib.ptr(self.pose_step_dev), _lib.ptr(self.pose_lr_dev), _lib.ptr(self.inv_scale), _lib.ptr(self.pose_found_inf),
                  _lib.stream())

    def _launch_march(self):
        m, opt, N, cap = self.model, self.model.opt, self.N, self.cap
        st = _lib.stream()
        P = _lib.ptr
        if self.pose is not None:      # provide_refined_poses + get_rays (barf/camera_optimizers.py:92-107, train_utils.py:150-165)
            _lib.call("ngp_pose_rays_forward", P(self.se3), P(self.poses), self.pose_stride, P(self.cam_idx), P(self.dirs_cam), N,
                      self.se3.shape[0], P(self.rays_o), P(self.rays_d), st)
        # per-ray jitter (raymarching.py:295) and random background (train_utils.py:496) from the library's counter-based generator:
        # torch's own costs two eager int64 fills (Philox seed / offset) in front of every graph replay
        if self.perturb:
            _lib.call("ngp_uniform", P(self.noises), N, self._rng_seed, P(self._rng_counter), st)
        if self.random_bg:
            _lib.call("ngp_uniform", P(self.bg_rays), 3 * N, self._rng_seed ^ 0x5DEECE66D, P(self._rng_counter), st)
        aabb = m.aabb_train
        _lib.call("ngp_march_rays_train_count_ex", P(self.rays_o), P(self.rays_d), P(aabb), float(m.min_near), P(self.cam_near_far),
                  P(self.n_rays_dev), P(m.density_bitfield), float(m.real_bound), int(bool(opt.contract)), float(opt.dt_gamma),
                  int(opt.max_steps), N, int(m.cascade), int(m.grid_size), P(self.noises), cap, None, None, P(self.rays),
                  P(self.counter), P(self.t_scratch), st)
        _lib.call("ngp_march_rays_train_write", P(self.rays_o), P(self.rays_d), P(self.rays_ldir), P(m.density_bitfield),
                  float(m.real_bound), int(bool(opt.contract)), float(opt.dt_gamma), int(opt.max_steps), N, int(m.cascade),
                  int(m.grid_size), None, None, None, P(self.rays), cap, self._m_dev, P(self.t_scratch), P(self.xyzs),
                  P(self.dirs), P(self.ts), P(self.ldirs), st)

    def _launch_field(self):
        m, opt, N, cap, ct = self.model, self.model.opt, self.N, self.cap, self._ct
        st = _lib.stream()
        P = _lib.ptr
        S, H, L, gt, ac, ip = self._grid_scalars
        enc = m.grid_encoder
        c1 = (ct.c_uint32 * 4)(*self.p1)
        c2 = (ct.c_uint32 * 4)(*self.p2)
        w1, w2 = self._ptrs(self._w_lp_views[:3]), self._ptrs(self._w_lp_views[3:])
        a1, a2 = self._ptrs(self.acts1), self._ptrs(self.acts2)
        def composite():
            _lib.call("ngp_composite_train_loss", P(self.sigma), P(self.rgb), P(self.ts), P(self.rays), cap, self._m_dev, N,
                      float(opt.T_thresh), self.bg_color, P(self.target), self.loss_scale, P(self.image), P(self.ray_loss),
                      P(self.loss), P(self.ticket), P(self.d_sigma), P(self.d_rgb), self.loss_mode, P(self.exposure),
                      ct.byref(self._loss_opts), st)
            if self.adaptive:      # the live ray count of the NEXT batch (train_utils.py:563-564), from this step's sample count
                _lib.call("ngp_adaptive_num_rays", P(self.n_rays_dev), self.counter.data_ptr(), self.num_points, N, st)

        dw1, dw2 = self._ptrs(self._w_grad_views[:3]), self._ptrs(self._w_grad_views[3:])
        if self.ws:
            _lib.call("ngp_field_forward_full", P(self.xyzs), P(self.dirs), P(self.ldirs), P(enc.embeddings), P(enc.offsets),
                      P(self.feat_weights), float(m.bound), S, H, L, gt, ac, ip, w1, c1, w2, c2, cap, self._m_dev, self._density_act,
                      float(opt.beta), self._color_act, P(self.enc_buf), a1, P(self.in2), a2, P(self.sigma), P(self.rgb),
                      P(self.dydx), st)
            composite()
            _lib.call("ngp_field_backward_full", P(self.xyzs), P(self.d_sigma), P(self.sigma), P(self.d_rgb), P(self.rgb),
                      P(self.enc_buf), a1, P(self.in2), a2, P(enc.offsets), P(self.feat_weights), float(m.bound), S, H, L, gt, ac, ip,
                      w1, c1, w2, c2, cap, self._m_dev, self._density_act, float(opt.beta), self._color_act, P(self.table_grad),
                      dw1, dw2, P(self.dydx), P(self.dirs) if self.ray_grads else None,
                      P(self.d_xyzs), P(self.d_dirs), st)
            self._launch_ray_backward()
            return
        ld2 = self.p2[0]
        if self.ws_fwd:       # one warp-specialised forward (16-column panels for the 48 / 80-wide layers), plain-row saves
            _lib.call("ngp_field_forward_full", P(self.xyzs), P(self.dirs), P(self.ldirs), P(enc.embeddings), P(enc.offsets),
                      P(self.feat_weights), float(m.bound), S, H, L, gt, ac, ip, w1, c1, w2, c2, cap, self._m_dev, self._density_act,
                      float(opt.beta), self._color_act, P(self.enc_buf), a1, P(self.in2), a2, P(self.sigma), P(self.rgb),
                      P(self.dydx), st)
        else:
            _lib.call("ngp_field_forward_density", P(self.xyzs), P(self.dirs), P(self.ldirs), P(enc.embeddings), P(enc.offsets),
                      P(self.feat_weights), float(m.bound), S, H, L, gt, ac, ip, w1, c1, 3, cap, self._m_dev, self._density_act,
                      float(opt.beta), P(self.enc_buf), a1, P(self.sigma), P(self.in2), ld2, st)
            _lib.call("ngp_mlp_forward_rgb", P(self.in2), ld2, w2, c2, 3, cap, self._m_dev, 1, self._color_act, P(self.rgb), a2, st)
        composite()
        _lib.call("ngp_mlp_backward_rgb", P(self.d_rgb), P(self.rgb), self._color_act, P(self.in2), ld2, w2, a2, c2, 3, cap,
                  self._m_dev, 1, P(self.d_in2), ld2, dw2, st)
        if self.ray_grads:      # d dirs from d SH(dir) = d in2[:, 15:31] (the warp-specialised backward does this in its V0 group)
            _lib.call("ngp_sh_dirs_backward", P(self.d_in2), ld2, 15, P(self.dirs), cap, self._m_dev, P(self.d_dirs), st)
        _lib.call("ngp_field_backward_density", P(self.xyzs), P(self.d_sigma), P(self.sigma), P(self.d_in2), ld2, P(self.enc_buf),
                  P(self.dydx), P(enc.offsets), P(self.feat_weights), float(m.bound), S, H, L, gt, ac, ip, w1, a1, c1, 3, cap, self._m_dev,
                  self._density_act, float(opt.beta), P(self.table_grad), dw1, P(self.d_xyzs), st)
        self._launch_ray_backward()

    def _launch_ray_backward(self):
        """dL/d rays_o = sum_seg dL/dxyz, dL/d rays_d = sum_seg (dL/dxyz * t + dL/ddirs)  (raymarching.py:319-329), then d se3."""
        P, st, N = _lib.ptr, _lib.stream(), self.N
        if self.ray_grads:
            _lib.call("ngp_march_rays_train_backward", P(self.d_xyzs), P(self.d_dirs), P(self.ts), P(self.rays), N, self.cap,
                      P(self.d_rays_o), P(self.d_rays_d), st)
        if self.pose is not None:
            _lib.call("ngp_pose_rays_backward", P(self.d_rays_o), P(self.d_rays_d), P(self.se3), P(self.poses), self.pose_stride,
                      P(self.cam_idx), P(self.dirs_cam), N, self.se3.shape[0], P(self.se3_grad), st)

    def _launch_check(self, count_step=False):
        """inf / nan check of both gradient buffers -> self.found_inf (overwritten) in one launch; count_step: the same kernel
        advances the device-side Adam step counter unless the step will be skipped."""
        ct = self._ct
        if not hasattr(self, "_chk_args"):
            self._chk_scratch = torch.zeros(2, device=self.dev, dtype=torch.int32)
            self._chk_args = ((ct.c_void_p * 2)(self.table_grad.data_ptr(), self.w_grad.data_ptr()), (ct.c_int * 2)(_lib.NGP_F16, _lib.NGP_F32),
                              (ct.c_uint64 * 2)(self.table_grad.numel(), self.w_grad.numel()))
        g, d, n = self._chk_args
        _lib.call("ngp_check_finite_multi", g, d, n, 2, _lib.ptr(self.found_inf), _lib.ptr(self.opt_step_dev) if count_step else None,
                  _lib.ptr(self._chk_scratch), _lib.stream())

    def _launch_optimizer(self):
        """inf check + step count (one kernel) + fused Adam of the table and of the MLP weights."""
        if self.world == 1:
            self._launch_check(count_step=True)
        self.opt.step(self.inv_scale, self.found_inf, zero_grad=True, count_step=self.world != 1)

    def _launch_scaler_update(self):
        """GradScaler.update() on the device, after every optimizer of the step has consumed inv_scale and before the next loss
        reads the scale (main stream, behind the join with the side stream)."""
        if not self.dynamic:
            return
        _lib.call("ngp_grad_scaler_update", _lib.ptr(self.scale_dev), _lib.ptr(self.inv_scale), _lib.ptr(self.scaler_state),
                  _lib.ptr(self.found_inf), _lib.ptr(self.pose_found_inf) if self.pose is not None else None, 2.0, 0.5,
                  self.growth_interval, self.world, _lib.stream())

    def skipped_steps(self):
        """Optimizer steps skipped so far because of inf / nan gradients (one 4-byte read; the step itself never syncs)."""
        return int(self.scaler_state[1].item())

    def _launch_pipelined(self):
        """[optimizer update of the PREVIOUS step]  ||  [march of this step]  ->  field forward -> composite -> backward.
        The marcher depends on the rays and the occupancy bitfield only, and it is latency bound (one warp per ray) while
        Adam streams 366 MB through HBM: on two streams they overlap almost completely."""
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        ev = None
        if self.pose is not None:
            self._launch_pose_update()
            ev = torch.cuda.Event()
            ev.record(main)
        with torch.cuda.stream(self._side):
            self._launch_optimizer()
            # GradScaler.update() behind every consumer of inv_scale (both optimizers), off the main stream: the marcher is
            # the longer branch, so this launch costs nothing there
            if ev is not None:
                self._side.wait_event(ev)
            self._launch_scaler_update()
        self._launch_march()
        main.wait_stream(self._side)
        self._launch_field()

    def _capture(self):
        # one eager pass first: sets the dynamic shared-memory attributes and warms the allocator outside the capture
        s = torch.cuda.Stream()
        self._side = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._launch_forward_backward()
            self._launch_check()
            self.table_grad.zero_()
            self.w_grad.zero_()
            if self.pose is not None:
                self.se3_grad.zero_()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph_fb, self._graph_pipe, self._graph_chk = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        c0 = _lib.launch_count
        with torch.cuda.graph(self._graph_fb):
            self._launch_forward_backward()
        self.graph_kernels = _lib.launch_count - c0
        n_adam = self.opt.step_count
        n_pose = self.pose_opt.step_count if self.pose is not None else 0
        c0 = _lib.launch_count
        with torch.cuda.graph(self._graph_pipe):
            self._launch_pipelined()
        self.pipe_kernels = _lib.launch_count - c0
        self.opt.step_count = n_adam        # capturing is not stepping
        if self.pose is not None:
            self.pose_opt.step_count = n_pose
        with torch.cuda.graph(self._graph_chk):
            self._launch_check()
        if self.world > 1:      # data parallel: the collectives stay outside the graphs, between a march graph and a field graph
            self._graph_march, self._graph_field = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            c0 = _lib.launch_count
            with torch.cuda.graph(self._graph_march):
                self._launch_march()
            self.march_kernels = _lib.launch_count - c0
            c0 = _lib.launch_count
            with torch.cuda.graph(self._graph_field):
                self._launch_field()
            self.field_kernels = _lib.launch_count - c0
            # the peer-memory update chain (4 launches + 2 symmetric-memory barriers) as a graph of its own: replayed on the side
            # stream it costs one launch of host time and no gaps between its small kernels.  NGP_DP_GRAPH=0 keeps it eager.
            # The peer-memory update chain (4 launches + 2 symmetric-memory barriers) and the march as ONE graph with two branches:
            #   side:  check + publish | barrier 0 | Adam(table) | Adam(MLPs) | [wait pose] | barrier 1 | finish
            #   main:  [barrier 2 | pose Adam over the peers' se3 gradients] | march                      -> join
            # one launch of host time per step and no gaps between the small kernels.  NGP_DP_GRAPH=0 keeps everything eager.
            self._graph_update = self._graph_um = None
            if self.peer is not None and os.environ.get("NGP_DP_GRAPH", "1") != "0":
                n_adam, epoch = self.opt.step_count, _lib.weights_epoch
                n_pose = self.pose_opt.step_count if self.pose is not None else 0
                try:
                    if self.pose is None:        # the chain alone, for tools/dp_timeline.py
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._peer_update()
                        self._graph_update = g
                    g = torch.cuda.CUDAGraph()
                    c0 = _lib.launch_count
                    with torch.cuda.graph(g):
                        main = torch.cuda.current_stream()
                        self._side.wait_stream(main)
                        ev = None
                        if self.pose is not None:
                            self.peer.barrier(2)                   # every rank's se3 gradient is complete
                            self._launch_pose_update()
                            ev = torch.cuda.Event()
                            ev.record(main)
                        with torch.cuda.stream(self._side):
                            self._peer_update(pose_event=ev)
                            self._launch_scaler_update()
                        self._launch_march()
                        main.wait_stream(self._side)
                    self.um_kernels = _lib.launch_count - c0 + 2 + (1 if self.pose is not None else 0)
                    self.update_kernels = self.um_kernels - self.march_kernels
                    self._graph_um = g
                except Exception as e:       # barrier not capturable in this torch build: eager launches, same result
                    import warnings
                    warnings.warn(f"FusedTrainStep: the peer update could not be captured ({e!r}); launching it eagerly")
                    self._graph_update = self._graph_um = None
                self.opt.step_count, _lib.weights_epoch = n_adam, epoch
                if self.pose is not None:
                    self.pose_opt.step_count = n_pose
        self.table_grad.zero_()
        self.w_grad.zero_()
        if self.pose is not None:
            self.se3_grad.zero_()

    def set_lr(self, lr=None, pose_lr=None):
        """Learning-rate schedule (LambdaLR of main.py:258-261; ExponentialLR of the pose optimizer): the captured Adam kernels
        read the rate from device memory, so a schedule is one 4-byte fill per step and no re-capture."""
        if lr is not None:
            self.lr_dev.fill_(float(lr))
        if pose_lr is not None and self.pose is not None:
            self.pose_lr_dev.fill_(float(pose_lr))

    def _reduce_and_update(self, pose_inline=False):
        """Data parallel: all-reduce of the two gradient buffers, inf check, MAX of the flag, fused Adam -- on the current
        (side) stream, so that it overlaps the ray marching of the next step on the main stream."""
        if self.peer is not None:
            return self._peer_update(pose_inline=pose_inline)
        parallel.all_reduce_gradients([self.table_grad, self.w_grad], None, self.pg)
        # The inf / nan check runs on the REDUCED buffers, which are bit-identical on every rank (a non-finite value of any
        # rank survives the SUM), so all ranks take the same skip decision without a second collective for the flag.
        self._launch_check(count_step=True)
        self.opt.step(self.inv_scale, self.found_inf, zero_grad=True, count_step=False)

    def _peer_update(self, pose_inline=False, pose_event=None):
        """reduce-scatter + Adam + all-gather of the table as ONE kernel over peer memory (and a replicated variant of the same
        kernel for the 57 KB of MLP weights); the GradScaler flag travels through peer stores; two symmetric-memory barriers
        order the ranks.  No NCCL call, no staging buffer, Adam on 1 / world of the table.  Four launches + two barriers:
        check + flag publish | barrier | Adam(table) | Adam(MLPs) | barrier | finish (merged flag, step count, gradient clear).
        The pose optimizer's se3 gradient lives behind the MLP gradients in the same symmetric buffer: its update
        (ngp_dp_small_adam, which sums the ranks' buffers itself) runs either inside this chain behind the first barrier
        (pose_inline, the eager path) or on the main stream behind a barrier of its own, in front of the ray generation; in that
        case pose_event (recorded behind it) keeps this rank out of the closing barrier until it has read every peer's buffer."""
        st, P, pm = _lib.stream(), _lib.ptr, self.peer
        world, rank = self.world, pm.rank
        _lib.weights_epoch += 1
        self.opt.step_count += 1
        b1, b2 = self.opt.betas
        if not hasattr(self, "_chk_args"):
            self._launch_check()                                   # builds the argument arrays (and is harmless)
        g, d, n = self._chk_args
        _lib.call("ngp_dp_check_publish", g, d, n, 2, P(self.found_inf), P(self._chk_scratch), pm.ptrs["flags"], world, rank, st)
        pm.barrier(0)                                              # every rank's gradients and flags are complete
        if pose_inline and self.pose is not None:
            self._launch_pose_update()
        lo, hi = self.shard
        _lib.call("ngp_dp_fused_adam", pm.ptrs["grad"], _lib.NGP_F16, pm.ptrs["table"], _lib.NGP_F16, world, world,
                  P(self.table_master_shard), P(self.shard_m), P(self.shard_v), lo, hi, float(self.opt.lr), float(b1), float(b2),
                  float(self.opt.eps), float(self.opt.weight_decay), P(self.opt_step_dev), P(self.lr_dev), P(self.inv_scale),
                  None, P(self.flags), world, st)
        _lib.call("ngp_dp_fused_adam", pm.ptrs["w_grad"], _lib.NGP_F32, self._w_lp_local, _lib.NGP_F16, world, 1, P(self.w_master),
                  P(self.w_m), P(self.w_v), 0, self.w_master.numel(), float(self.opt.lr), float(b1), float(b2), float(self.opt.eps),
                  float(self.opt.weight_decay), P(self.opt_step_dev), P(self.lr_dev), P(self.inv_scale), None, P(self.flags), world, st)
        if pose_event is not None:
            torch.cuda.current_stream().wait_event(pose_event)
        pm.barrier(1)                                              # all parameter stores have landed, all gradient loads are done
        _lib.call("ngp_dp_finish", P(self.flags), world, P(self.found_inf), P(self.opt_step_dev), P(self.table_grad),
                  self.table_grad.numel() * self.table_grad.element_size(), P(self._w_grad_all), self._w_grad_all.numel() * 4, st)

    def gather_table_master(self):
        """fp32 master copy of the whole table (checkpoints).  In peer mode every rank holds only its shard: all-gather them."""
        if self.peer is None:
            return self.table_master
        n = self.table_grad.numel()
        per = (n + 8 * self.world - 1) // (8 * self.world) * 8
        mine = torch.zeros(per, device=self.dev, dtype=torch.float32)
        mine[:self.table_master_shard.numel()] = self.table_master_shard
        full = torch.empty(per * self.world, device=self.dev, dtype=torch.float32)
        dist.all_gather_into_tensor(full, mine, group=self.pg)
        return full[:n].view_as(self.table_grad)

    # ---- checkpoints (nerf/train_utils.py:1141-1299 saves model.state_dict() + optimizer + scaler) -------------------------------
    def _full(self, shard):
        """Peer mode: the flat fp32 tensor whose contiguous shards live on the ranks (all-gather); otherwise the tensor itself."""
        if self.peer is None:
            return shard
        n = self.table_grad.numel()
        per = (n + 8 * self.world - 1) // (8 * self.world) * 8
        mine = torch.zeros(per, device=self.dev, dtype=torch.float32)
        mine[:shard.numel()] = shard.reshape(-1)
        full = torch.empty(per * self.world, device=self.dev, dtype=torch.float32)
        dist.all_gather_into_tensor(full, mine, group=self.pg)
        return full[:n].view_as(self.table_grad)

    def _table_state(self):
        if self.peer is None:
            g = self.opt.groups[0]
            return self.table_master, g["m"], g["v"]
        return self.table_master_shard, self.shard_m, self.shard_v

    def _w_state(self):
        if self.peer is None:
            g = self.opt.groups[1]
            return g["m"], g["v"]
        return self.w_m, self.w_v

    def model_state_dict(self):
        """model.state_dict() with the hash table as the fp32 master copy -- the dtype of the reference's checkpoints
        (gridencoder/grid.py:43-46 keeps the embeddings in fp32); the module itself holds the fp16 working copy."""
        self.flush()
        sd = dict(self.model.state_dict())
        sd["grid_encoder.embeddings"] = self._full(self._table_state()[0]).detach().clone().view_as(self.table_grad)
        return sd

    def state_dict(self):
        """Everything needed to resume: fp32 masters, Adam moments, device-side step counters, the loss scale."""
        self.flush()
        tm, tmm, tv = self._table_state()
        wm, wv = self._w_state()
        sd = {"table_master": self._full(tm).detach().clone(), "table_exp_avg": self._full(tmm).detach().clone(),
              "table_exp_avg_sq": self._full(tv).detach().clone(), "w_master": self.w_master.clone(), "w_exp_avg": wm.clone(),
              "w_exp_avg_sq": wv.clone(), "opt_step": self.opt_step_dev.clone(), "loss_scale": self.scale_dev.clone(),
              "scaler_state": self.scaler_state.clone(), "lr": self.lr_dev.clone(), "global_step": self.global_step,
              "rng": (self._rng_seed, self._rng_counter.clone())}
        if self.pose is not None:
            g = self.pose_opt.groups[0]
            sd.update({"se3": self.se3.clone(), "se3_exp_avg": g["m"].clone(), "se3_exp_avg_sq": g["v"].clone(),
                       "pose_step": self.pose_step_dev.clone(), "pose_lr": self.pose_lr_dev.clone()})
        return sd

    def _scatter_table(self, full, dst):
        full = full.to(self.dev, torch.float32).reshape(-1)
        if self.peer is None:
            dst.view(-1).copy_(full)
        else:
            lo, hi = self.shard
            dst.copy_(full[lo:hi])

    def load_state_dict(self, sd):
        self.flush()
        tm, tmm, tv = self._table_state()
        wm, wv = self._w_state()
        self._scatter_table(sd["table_master"], tm)
        self._scatter_table(sd["table_exp_avg"], tmm)
        self._scatter_table(sd["table_exp_avg_sq"], tv)
        self.w_master.copy_(sd["w_master"])
        wm.copy_(sd["w_exp_avg"])
        wv.copy_(sd["w_exp_avg_sq"])
        self.opt_step_dev.copy_(sd["opt_step"])
        self.opt.step_count = int(sd["opt_step"].item())
        self.scale_dev.copy_(sd["loss_scale"])
        self.scaler_state.copy_(sd["scaler_state"])
        self.inv_scale.copy_(1.0 / (self.scale_dev * self.world))
        self.lr_dev.copy_(sd["lr"])
        self.global_step = int(sd["global_step"])
        if "rng" in sd:      # the jitter / background stream continues where the checkpoint left it
            self._rng_seed = int(sd["rng"][0])
            self._rng_counter.copy_(sd["rng"][1])
            self._graph = None      # the seed is a launch argument of the captured kernels
        if self.pose is not None and "se3" in sd:
            g = self.pose_opt.groups[0]
            self.se3.copy_(sd["se3"])
            g["m"].copy_(sd["se3_exp_avg"])
            g["v"].copy_(sd["se3_exp_avg_sq"])
            self.pose_step_dev.copy_(sd["pose_step"])
            self.pose_lr_dev.copy_(sd["pose_lr"])
        self._refresh_working_copies(sd["table_master"])

    def _refresh_working_copies(self, table_full):
        """fp16 working copies (what the kernels read) from the masters."""
        self.model.grid_encoder.embeddings.data.copy_(table_full.to(self.dev).view_as(self.table_grad))
        self.w_lp.copy_(self.w_master)
        _lib.weights_epoch += 1

    def load_model_state_dict(self, sd, strict=True):
        """model.load_state_dict() for a checkpoint in the reference layout (fp32 hash table, train_utils.py:1283-1299): the
        module's parameters are written in place (the captured kernels keep their buffers), the fp32 table goes into the master
        copy without passing through fp16, and the fp16 working copies are refreshed."""
        self.flush()
        result = self.model.load_state_dict(sd, strict=strict)
        self._scatter_table(sd["grid_encoder.embeddings"], self._table_state()[0])
        self.w_lp.copy_(self.w_master)
        _lib.weights_epoch += 1
        return result

    def sync_from_model(self):
        """Call after writing the module's parameters behind the trainer's back (model.load_state_dict(...), manual edits): the
        fp32 masters are re-read from the module -- the table from grid_encoder.embeddings (fp16 working copy or a loaded fp32
        tensor), the MLP weights are views of the master already -- and the fp16 working copies are refreshed.  Adam moments are
        kept; use load_state_dict() to restore them."""
        self.flush()
        emb = self.model.grid_encoder.embeddings.data
        tm = self._table_state()[0]
        self._scatter_table(emb.float(), tm)
        if emb.dtype != torch.float16:      # a fp32 table was assigned: put the fp16 working copy back (same storage as before)
            raise RuntimeError("FusedTrainStep.sync_from_model: grid_encoder.embeddings.data was replaced; copy into it in place "
                               "(load_state_dict does) so that the captured kernels keep reading the same buffer")
        self.w_lp.copy_(self.w_master)
        _lib.weights_epoch += 1

    def flush(self):
        """Applies the optimizer update that is still pending (the update of step k normally runs at the start of step
        k + 1, overlapped with its ray marching).  Call before using the model outside of step()."""
        if not self._pending:
            return
        peer_pose = self.peer is not None and self.pose is not None
        if self.world > 1:
            if self.pose is not None and not peer_pose:
                parallel.all_reduce_gradients([self.se3_grad], None, self.pg)
            self._reduce_and_update(pose_inline=peer_pose)
        else:
            self._launch_optimizer()
        if self.pose is not None and not peer_pose:
            self._launch_pose_update()
        self._launch_scaler_update()
        self._pending = False

    def profile_kernels(self, iters=10):
        """Average device time (ms, CUDA events on the launching stream) of every kernel of the step, launched eagerly in
        step order on the current inputs (these are real optimisation steps: the model trains `iters` steps).  Returns {entry point: ms}.
        With world > 1 the update is a collective (peer-memory kernel or NCCL all-reduce), so only the rank-local kernels are
        timed: gradients are discarded and no optimizer step is taken or counted -- a single rank may call this and the replicas
        stay identical."""
        self.flush()
        local_only = self.world > 1
        names, events = [], []
        real_call = _lib.call

        def timed_call(name, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real_call(name, *args)
            e1.record()
            names.append(name)
            events.append((e0, e1))

        try:
            _lib.call = timed_call
            for _ in range(iters):
                self._launch_forward_backward()
                self._launch_check(count_step=not local_only)
                if local_only:
                    self.table_grad.zero_()
                    self.w_grad.zero_()
                    if self.pose is not None:
                        self.se3_grad.zero_()
                    continue
                self.opt.step(self.inv_scale, self.found_inf, zero_grad=True, count_step=False)
                if self.pose is not None:
                    self._launch_pose_update()
                self._launch_scaler_update()
        finally:
            _lib.call = real_call
        torch.cuda.synchronize()
        out = {}
        for n, (e0, e1) in zip(names, events):
            out.setdefault(n, []).append(e0.elapsed_time(e1))
        return {n: sum(v) / iters for n, v in out.items()}

    def set_rays(self, rays_o, rays_d, target_rgb, rays_ldir=None, exposure=None, **extra):
        """Copies the step's inputs into the static buffers (pinned host tensors are copied asynchronously).  With rgba_targets
        target_rgb is [N, 4] (or pass target_alpha=[N]); extra: lossmult, loss_weight [N, 3], cam_near_far [N, 2]."""
        if target_rgb is not None and target_rgb.shape[-1] == 4:
            if self.target_alpha is None:
                raise RuntimeError("FusedTrainStep: RGBA targets need rgba_targets=True")
            extra.setdefault("target_alpha", target_rgb[..., 3])
            target_rgb = target_rgb[..., :3]
        for dst, src in ((self.rays_o, rays_o), (self.rays_d, rays_d), (self.target, target_rgb), (self.rays_ldir, rays_ldir),
                         (self.exposure, exposure)):
            if dst is None or src is None or (src.is_cuda and src.data_ptr() == dst.data_ptr()):
                continue
            dst.copy_(src.reshape(dst.shape), non_blocking=True)
        self._set_extra(extra)

    def _set_extra(self, extra):
        for name, src in extra.items():
            if name not in ("target_alpha", "lossmult", "loss_weight", "cam_near_far"):
                raise TypeError(f"FusedTrainStep.set_rays: unknown input {name!r}")
            dst = getattr(self, name)
            if src is None:
                continue
            if dst is None:
                raise RuntimeError(f"FusedTrainStep: {name} was not enabled at construction")
            src = torch.as_tensor(src)
            if name in ("lossmult", "loss_weight"):      # scalars and [N, 1] broadcast like torch.broadcast_to (train_utils.py:518)
                src = torch.broadcast_to(src.to(dst.device, torch.float32), dst.shape)
            dst.copy_(src.reshape(dst.shape), non_blocking=True)
            if name == "lossmult":                       # the normaliser lossmult_tensor.sum() (train_utils.py:536), on the device
                self.inv_norm.copy_(self.lossmult.sum().reciprocal().reshape(1))

    def set_camera_rays(self, cam_idx, dirs_cam, target_rgb, exposure=None, **extra):
        """Pose mode: the step's rays as (camera index [N], camera-space pixel direction [N, 3]); rays_o / rays_d are produced
        on the device from the refined poses inside the captured step."""
        for dst, src in ((self.cam_idx, cam_idx), (self.dirs_cam, dirs_cam), (self.target, target_rgb), (self.exposure, exposure)):
            if src is None or (src.is_cuda and src.data_ptr() == dst.data_ptr()):
                continue
            dst.copy_(src.reshape(dst.shape), non_blocking=True)
        self._set_extra(extra)

    @property
    def last_num_points(self):
        return int(self.counter[0].item())

    def step(self, rays_o=None, rays_d=None, target_rgb=None, rays_ldir=None, update_grid=True, exposure=None, cam_idx=None,
             dirs_cam=None, **extra):
        """One optimisation step; returns the (unscaled) loss as a 1-element device tensor (overwritten by the next step).
        The parameter update of this step is applied at the start of the next call (or by flush()).
        Pose mode (pose_optimizer given): pass cam_idx / dirs_cam instead of rays_o / rays_d."""
        model = self.model
        if update_grid and self.global_step % self.update_extra_interval == 0:
            self.flush()                       # the density queries of the occupancy update see the updated weights
            model.update_extra_state()
        if rays_o is not None:
            self.set_rays(rays_o, rays_d, target_rgb, rays_ldir, exposure, **extra)
        if cam_idx is not None:
            self.set_camera_rays(cam_idx, dirs_cam, target_rgb, exposure, **extra)
        if self.feat_weights is not None and self._feat_annealing != model.annealing:
            # BARF window: computed on the host, one 128-byte copy (the torch expression is ~10 elementwise kernels)
            self._feat_host.copy_(torch.from_numpy(model._feat_weights_host()))
            self.feat_weights.copy_(self._feat_host, non_blocking=True)
            self._feat_annealing = model.annealing
        if self.use_graph:
            sig = (model.density_bitfield.data_ptr(), model.aabb_train.data_ptr(), model.grid_encoder.embeddings.data_ptr())
            if self._graph != sig:      # first step, or a captured buffer was re-allocated (update_aabb, load_state_dict)
                self.flush()
                self._capture()
                self._graph = sig
            if self.world > 1:
                # [all-reduce + optimizer update of the previous step, side stream]  ||  [march graph]  ->  field graph
                main = torch.cuda.current_stream()
                scaler_done = False
                if self._pending and getattr(self, "_graph_um", None) is not None:
                    _lib.weights_epoch += 1
                    self.opt.step_count += 1
                    if self.pose is not None:
                        self.pose_opt.step_count += 1
                    self._graph_um.replay()                    # update chain + GradScaler update || [pose update] march, joined
                    self.kernels_replayed += self.um_kernels
                    scaler_done = True
                else:
                    if self._pending:
                        peer_pose = self.peer is not None and self.pose is not None
                        if self.pose is not None and not peer_pose:      # tiny, and the march needs the updated poses: main stream, first
                            parallel.all_reduce_gradients([self.se3_grad], None, self.pg)
                        self._side.wait_stream(main)
                        with torch.cuda.stream(self._side):
                            self._reduce_and_update(pose_inline=peer_pose)
                        if peer_pose:
                            main.wait_stream(self._side)           # eager fallback: no overlap with the march
                        elif self.pose is not None:
                            self._launch_pose_update()
                    self._graph_march.replay()
                    self.kernels_replayed += self.march_kernels
                main.wait_stream(self._side)
                if self._pending and not scaler_done:
                    self._launch_scaler_update()
                self._graph_field.replay()
                self.kernels_replayed += self.field_kernels
            elif self._pending:
                _lib.weights_epoch += 1          # the replayed graph contains the optimizer update
                self._graph_pipe.replay()
                self.opt.step_count += 1
                if self.pose is not None:
                    self.pose_opt.step_count += 1
                self.kernels_replayed += self.pipe_kernels
            else:
                self._graph_fb.replay()
                self.kernels_replayed += self.graph_kernels
        else:
            self.flush()
            self._launch_forward_backward()
        self._pending = True
        self.global_step += 1
        return self.loss
