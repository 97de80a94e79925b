"""NeRFRenderer -- the cuda_ray rendering path of the reference (nerf/renderer.py) on the B200 operators.

In scope: __init__ (:161-198), update_aabb (:211-217), render/run_cuda (:374-377, :515-676),
mark_untrained_grid (:716-809), update_extra_state (:811-897), the differentiable near_far_from_aabb (:139-158) and the
proposal-network path `run` (:405-513, `cuda_ray=False`; helpers in nerf/proposal.py).  Mesh export is not provided.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, raymarching


def default_opt(**overrides):
    """The hot-path relevant defaults of the reference CLI (main.py:31-92)."""
    opt = SimpleNamespace(
        bound=2, contract=False, grid_size=128, min_near=0.05, density_thresh=10, cuda_ray=True, dt_gamma=0,
        max_steps=1024, T_thresh=1e-8, fp16=True, hashmap_size=19, hashgrid_resolution=2048, rfield=False,
        pose_opt="none", internal_activation="relu", beta=1.0, density_activation="clamped_exp",
        color_activation="clamped_exp", start_annealing=0.0, end_annealing=0.5, lambda_orientation=0,
        compute_normals=False, device="cuda", num_cameras=0, update_extra_interval=16,
        num_steps=[256, 96, 48], background="black", lambda_proposal=1, lambda_distort=0, max_ray_batch=4096 * 4,
    )
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


@torch.amp.autocast("cuda", enabled=False)
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.05):
    """Differentiable slab test, returns near [N,1], far [N,1] (renderer.py:139-158)."""
    tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
    tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
    far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
    mask = far < near
    near = torch.where(mask, torch.full_like(near, 1e9), near)
    far = torch.where(mask, torch.full_like(far, 1e9), far)
    near = torch.clamp(near, min=min_near)
    return near, far


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.real_bound = opt.bound                       # bound for ray marching (world space)
        self.bound = 2 if self.opt.contract else opt.bound  # bound for grid querying
        self.cascade = 1 + math.ceil(math.log2(self.bound))
        self.grid_size = opt.grid_size
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh

        aabb_train = torch.FloatTensor([-self.real_bound] * 3 + [self.real_bound] * 3)
        self.register_buffer("aabb_train", aabb_train)
        self.register_buffer("aabb_infer", aabb_train.clone())

        self.cuda_ray = opt.cuda_ray
        if not self.cuda_ray:           # proposal-network sampling (run): no occupancy grid
            self._mean_density = 0.0
            return
        self.register_buffer("density_grid", torch.zeros([self.cascade, self.grid_size ** 3]))
        self.register_buffer("density_bitfield", torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        self._mean_density = 0.0        # python float or a 1-element device tensor (synced lazily)
        self.iter_density = 0

    # mean_density stays a plain attribute for checkpoints (train_utils.py:1156-1157) but is produced on the device;
    # reading it is the only host sync of update_extra_state.
    @property
    def mean_density(self):
        if torch.is_tensor(self._mean_density):
            self._mean_density = float(self._mean_density.item())
        return self._mean_density

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density = v

    def forward(self, x, d, **kwargs):
        raise NotImplementedError()

    def density(self, x, **kwargs):
        raise NotImplementedError()

    def update_aabb(self, aabb):
        if not torch.is_tensor(aabb):
            aabb = torch.from_numpy(np.asarray(aabb)).float()
        self.aabb_train = aabb.clamp(-self.real_bound, self.real_bound).to(self.aabb_train.device)
        self.aabb_infer = self.aabb_train.clone()

    def render(self, rays_o, rays_d, **kwargs):
        if self.cuda_ray:
            return self.run_cuda(rays_o, rays_d, **kwargs)
        if self.training:
            return self.run(rays_o, rays_d, **kwargs)
        # staged inference (renderer.py:378-402): ray batches of opt.max_ray_batch
        N, B = rays_o.shape[0], int(self.opt.max_ray_batch)
        parts = [self.run(rays_o[h:h + B], rays_d[h:h + B], **kwargs) for h in range(0, N, B)]
        return {k: torch.cat([p[k] for p in parts], dim=0) for k in ("depth", "image", "weights_sum")}

    def run(self, rays_o, rays_d, bg_color=None, perturb=False, cam_near_far=None, shading="full", update_proposal=True, **kwargs):
        """Hierarchical sampling with the proposal networks (renderer.py:405-513): level 0 samples opt.num_steps[0] uniform bins in
        the contracted ray parameter, every further level resamples opt.num_steps[i] bins from the previous level's weights; the
        last level queries the NeRF field and is volume rendered.  rays [N, 3] -> image [N, 3], depth [N], weights_sum [N]."""
        from . import proposal as P
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        N, device = rays_o.shape[0], rays_o.device
        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer, self.min_near)
        if cam_near_far is not None:
            nears = torch.maximum(nears, cam_near_far[:, [0]])
            fars = torch.minimum(fars, cam_near_far[:, [1]])
        if bg_color is None:
            bg_color = 1
        s_near, s_far = P.spacing(nears), P.spacing(fars)
        steps = list(self.opt.num_steps)
        history = []                      # (bins, weights) of every level, for the inter-level loss
        bins = weights = rgbs = mid_t = None
        for level, T in enumerate(steps):
            if level == 0:
                bins = torch.linspace(0, 1, T + 1, device=device).unsqueeze(0).expand(N, -1)
                if perturb:
                    bins = (bins + (torch.rand_like(bins) - 0.5) / T).clamp(0, 1)
            else:
                bins = P.resample_bins(bins, weights, T + 1, perturb).detach()
            real_bins = P.spacing_inv(s_near * (1 - bins) + s_far * bins)          # [N, T + 1] in [near, far]
            mid_t = (real_bins[..., 1:] + real_bins[..., :-1]) / 2
            xyzs = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * mid_t.unsqueeze(2)
            pts = P.contract(xyzs) if self.opt.contract else xyzs
            if level != len(steps) - 1:
                with torch.set_grad_enabled(update_proposal and torch.is_grad_enabled()):
                    sigmas = self.density(pts, proposal=level)["sigma"]
            else:
                dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
                dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
                out = self(pts, dirs, None, shading=shading)
                sigmas, rgbs = out["sigma"], out["color"]
            weights = P.weights_from_sigmas(real_bins, sigmas, opaque_last=self.opt.background == "last_sample")
            if self.training:
                history.append((bins, weights))
        results = {}
        weights_sum = weights.sum(dim=-1)
        depth = (weights * mid_t).sum(dim=-1)
        image = (weights.unsqueeze(-1) * rgbs).sum(dim=-2)
        if self.training:
            results["num_points"] = xyzs.shape[0] * xyzs.shape[1]
            results["weights"] = weights
            if self.opt.lambda_proposal > 0 and update_proposal:
                results["proposal_loss"] = P.interlevel_loss([b for b, _ in history], [w for _, w in history])
            if self.opt.lambda_distort > 0:
                results["distort_loss"] = P.distortion_loss(bins, weights)
        results["image"] = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results["weights_sum"] = weights_sum
        results["depth"] = depth
        return results

    def run_cuda(self, rays_o, rays_d, rays_ldir=None, bg_color=None, perturb=False, cam_near_far=None,
                 update_proposal=True, shading="full", **kwargs):
        # rays_o, rays_d [N, 3] -> image [N, 3], depth [N]
        rays_o = rays_o.contiguous()
        rays_d = rays_d.contiguous()
        N = rays_o.shape[0]
        device = rays_o.device

        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer,
                                         self.min_near)
        if cam_near_far is not None:
            nears = torch.maximum(nears, cam_near_far[:, 0])
            fars = torch.minimum(fars, cam_near_far[:, 1])
        if bg_color is None:
            bg_color = 0
        results = {}
        amp = torch.amp.autocast("cuda", enabled=self.opt.fp16)

        if self.training:
            xyzs, dirs, ts, rays, ldirs = raymarching.march_rays_train(
                rays_o, rays_d, rays_ldir, self.real_bound, self.opt.contract, self.density_bitfield, self.cascade,
                self.grid_size, nears, fars, perturb, self.opt.dt_gamma, self.opt.max_steps)
            dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
            with amp:
                outputs = self(xyzs, dirs, ldirs, shading=shading)
                sigmas = outputs["sigma"]
                rgbs = outputs["color"]
            weights, weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, ts, rays,
                                                                                  self.opt.T_thresh)
            results["num_points"] = xyzs.shape[0]
            results["weights"] = weights
            results["weights_sum"] = weights_sum
            if self.opt.lambda_orientation > 0:
                pos = xyzs.clone().requires_grad_(True)
                normals = torch.autograd.grad(self(pos, dirs, shading=shading)["sigma"], pos,
                                              grad_outputs=torch.ones_like(sigmas), retain_graph=True)[0]
                normals = (-torch.nn.functional.normalize(normals, dim=-1) + 1) / 2
                zero = torch.tensor(0.0, dtype=torch.float32, device=device)
                n_dot_v = (normals * -dirs).sum(dim=-1)
                results["orientation_loss"] = torch.mean((weights * torch.min(zero, n_dot_v) ** 2).sum(dim=-1))
        else:
            dtype = torch.float32
            weights_sum = torch.zeros(N, dtype=dtype, device=device)
            depth = torch.zeros(N, dtype=dtype, device=device)
            image = torch.zeros(N, 3, dtype=dtype, device=device)
            fast = self._fast_infer_args(rays_ldir, shading) if self.FAST_INFER else None
            if fast is not None:
                self._march_composite_loop_fast(N, rays_o, rays_d, nears, fars, perturb, weights_sum, depth, image, fast)
            else:
                self._march_composite_loop(N, rays_o, rays_d, rays_ldir, nears, fars, perturb, shading, amp, weights_sum,
                                           depth, image, normals=False)
            if getattr(self.opt, "compute_normals", False):
                ws_n = torch.zeros(N, dtype=dtype, device=device)
                depth_n = torch.zeros(N, dtype=dtype, device=device)
                normalmap = torch.zeros(N, 3, dtype=dtype, device=device)
                self._march_composite_loop(N, rays_o, rays_d, rays_ldir, nears, fars, perturb, shading, amp, ws_n,
                                           depth_n, normalmap, normals=True)
                results["normals"] = normalmap + (1 - weights_sum).unsqueeze(-1) * bg_color

        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results["depth"] = depth
        results["image"] = image
        return results

    FAST_INFER = True

    def _fast_infer_args(self, rays_ldir, shading):
        """Subclasses whose field is one fused kernel return its launch arguments (see NeRFNetwork); None = generic loop."""
        return None

    LOOKAHEAD = 4          # iterations the host may run ahead of the last alive-ray count it has seen
    # Steps per ray and iteration: the reference takes n_step = max(min(N // n_alive, 8), 1) (renderer.py:598) with sample buffers
    # of N rows.  How a ray's sample sequence is cut into iterations does not change the image (composite_rays continues from
    # rays_t and the accumulated weights, sample by sample), only the number of iterations: with buffers of INFER_ROWS * N rows
    # and up to INFER_MAX_STEP steps a 1080p frame takes ~14 iterations instead of ~44.  (1, 8) is the reference's schedule.
    INFER_ROWS, INFER_MAX_STEP = 4, 16

    def _march_composite_loop_fast(self, N, rays_o, rays_d, nears, fars, perturb, weights_sum, depth, image, field):
        """The alive-ray loop of renderer.py:588-616 driven from the device: four launches per iteration (march, fused field,
        composite, two-pass compaction) on buffers allocated once per frame, with the loop header -- n_alive, n_step =
        max(min(N // n_alive, 8), 1), step += n_step, stop at max_steps -- kept in a device control block that the compaction
        kernel advances (csrc/raymarch.cu: ngp_*_dev).  The reference reads the surviving count back every iteration to size
        the next launches; here the host only polls, LOOKAHEAD iterations late and without ever idling the GPU, the counts that
        earlier iterations left in pinned memory: they bound the launch sizes and tell it when to stop enqueuing.  n_alive *
        n_step never exceeds N, the marcher writes every row of its outputs (finished rays as zeros), and the field kernel
        normalises the directions itself -- so the per-iteration zero fills, the torch normalisation and the autograd /
        autocast plumbing of the generic loop disappear as well."""
        from .. import _lib
        P, st = _lib.ptr, _lib.stream()
        dev = rays_o.device
        f32 = dict(dtype=torch.float32, device=dev)
        R = N * int(self.INFER_ROWS)              # rows of the sample buffers
        xyzs, dirs, ts = torch.empty(R, 3, **f32), torch.empty(R, 3, **f32), torch.empty(R, 2, **f32)
        sigmas, rgbs = torch.empty(R, **f32), torch.empty(R, 3, **f32)
        alive = torch.arange(N, dtype=torch.int32, device=dev)
        alive_next = torch.empty_like(alive)
        rays_t = nears.clone().view(-1).contiguous()
        fars = fars.view(-1).contiguous()
        zeros = torch.zeros(N, **f32)
        noises = torch.rand(N, **f32) if perturb else zeros
        n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty((N + 4095) // 4096 + 1, dtype=torch.int32, device=dev)
        max_steps = int(self.opt.max_steps)
        # first iteration: one step per ray (most rays of a frame never meet an occupied cell and leave the loop right there)
        ctl = torch.tensor([[N, 1, N, 0], [0, 1, 0, 0]], dtype=torch.int32, device=dev)      # ping-pong control blocks
        L = self.LOOKAHEAD
        if getattr(self, "_infer_poll", None) is None or self._infer_poll[0].shape[0] != L:
            self._infer_poll = (torch.zeros(L, 4, dtype=torch.int32).pin_memory(), [torch.cuda.Event() for _ in range(L)])
        seen, events = self._infer_poll
        bound = N                     # upper bound of n_alive known to the host
        it = 0
        while True:
            cur, nxt = ctl[it & 1], ctl[(it + 1) & 1]
            _lib.call("ngp_march_rays_dev", P(cur), bound, P(alive), P(rays_t), P(rays_o), P(rays_d), float(self.real_bound),
                      int(bool(self.opt.contract)), float(self.opt.dt_gamma), max_steps, int(self.cascade), int(self.grid_size),
                      P(self.density_bitfield), P(fars), P(xyzs), P(dirs), P(ts), P(noises if it == 0 else zeros), st)
            field(xyzs, dirs, R, sigmas, rgbs, st, m_dev=cur.data_ptr() + 8)
            _lib.call("ngp_composite_rays_dev", P(cur), bound, float(self.opt.T_thresh), P(alive), P(rays_t), P(sigmas), P(rgbs), P(ts),
                      P(weights_sum), P(depth), P(image), st)
            _lib.call("ngp_compact_rays_alive_dev", P(cur), P(nxt), bound, R, int(self.INFER_MAX_STEP), max_steps, P(alive), P(alive_next),
                      P(n_out), P(ws), st)
            alive, alive_next = alive_next, alive
            seen[it % L].copy_(nxt, non_blocking=True)
            events[it % L].record()
            it += 1
            if it >= L:               # the control block written L - 1 iterations ago: its copy has long landed, the GPU still has work queued
                j = (it - L) % L
                events[j].synchronize()
                bound = int(seen[j, 0])
                if bound == 0:
                    break
            if it > max_steps + L:    # cannot happen: step grows by >= 1 per iteration
                raise RuntimeError("inference loop did not terminate")

    def _march_composite_loop(self, N, rays_o, rays_d, rays_ldir, nears, fars, perturb, shading, amp, weights_sum,
                              depth, image, normals):
        """The alive-ray loop of renderer.py:588-616 (:626-668 for normals).  Compaction of the surviving ray ids
        happens on the device; only the surviving count comes back to the host (it sizes the next launch)."""
        device = rays_o.device
        n_alive = N
        rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
        rays_t = nears.clone().view(-1).contiguous()
        step = 0
        while step < self.opt.max_steps:
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, ts = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.real_bound,
                                                    self.opt.contract, self.density_bitfield, self.cascade,
                                                    self.grid_size, nears, fars, perturb if step == 0 else False,
                                                    self.opt.dt_gamma, self.opt.max_steps)
            dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
            with amp:
                ldirs = rays_ldir.repeat(xyzs.shape[0], 1) if self.opt.rfield else None
                outputs = self(xyzs, dirs, ldirs, shading=shading)
                sigmas = outputs["sigma"]
                if normals:
                    with torch.enable_grad():
                        pos = xyzs.clone().requires_grad_(True)
                        nrm = torch.autograd.grad(self(pos, dirs, ldirs, shading=shading)["sigma"], pos,
                                                  grad_outputs=torch.ones_like(sigmas), retain_graph=True)[0]
                        rgbs = (-torch.nn.functional.normalize(nrm, dim=-1) + 1) / 2
                else:
                    rgbs = outputs["color"]
            raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image,
                                       self.opt.T_thresh)
            rays_alive, count = raymarching.compact_rays_alive(rays_alive, n_alive)
            n_alive = int(count.item())
            step += n_step

    @torch.no_grad()
    def mark_untrained_grid(self, dataset, S=64):
        """Cells seen by no training camera, or outside the training AABB, get density -1 and never become occupied
        (renderer.py:716-809).  `dataset` provides poses [B, 3|4, 4] (camera-to-world), intrinsics (fx, fy, cx, cy) -- one
        4-vector or [B, 4] -- and optionally cam_near_far [B, 2].  One kernel: a thread per (cascade, cell), cameras staged
        through shared memory, early exit at the first camera that sees the cell (csrc/occupancy.cu); `S`, the chunk size of
        the reference's Python loops, has no meaning here and is accepted for signature compatibility."""
        dev = self.density_grid.device
        poses = torch.as_tensor(dataset.poses).to(dev).float().contiguous()
        B = poses.shape[0]
        intr = dataset.intrinsics
        if torch.is_tensor(intr):
            # renderer.py:731,776-777: fp32 tensor division, per camera when intrinsics is [B, 4]
            intr = intr.to(dev).float().reshape(-1, 4)
            half_fov = torch.stack([intr[:, 2] / intr[:, 0], intr[:, 3] / intr[:, 1]], dim=-1)
        else:
            # renderer.py:729,779-780,784-785: numpy scalars divide in float64, the quotient is cast when it meets the fp32 tensor
            fx, fy, cx, cy = [float(v) for v in np.asarray(intr).reshape(-1)[:4]]
            half_fov = torch.tensor([[cx / fx, cy / fy]], dtype=torch.float64).float().to(dev)
        half_fov = half_fov.contiguous()
        if half_fov.shape[0] not in (1, B):
            raise ValueError("mark_untrained_grid: intrinsics must be one (fx, fy, cx, cy) or one per pose")
        cnf = getattr(dataset, "cam_near_far", None)
        cam_near = None if cnf is None else torch.as_tensor(cnf).to(dev).float()[:, 0].contiguous()
        _lib.call("ngp_mark_untrained_grid", _lib.ptr(self.density_grid), _lib.ptr(poses), poses.shape[-2] * poses.shape[-1], B,
                  _lib.ptr(half_fov), half_fov.shape[0], _lib.ptr(cam_near), float(self.opt.min_near), _lib.ptr(self.aabb_train),
                  int(self.grid_size), int(self.cascade), float(self.bound), _lib.stream())

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """EMA-max update of the density grid + bitfield (renderer.py:811-897).

        Full update (first 16 calls): every cell of every cascade is queried once at a jittered position.  Partial
        update: H^3/4 uniform cells + H^3/4 occupied cells per cascade.  The per-cascade torch op chains of the
        reference are replaced by three kernels (sample positions, sigma scatter, fused EMA-max + clamped mean) and
        packbits reads min(mean, density_thresh) on the device."""
        if not self.cuda_ray:
            return
        H = self.grid_size
        H3 = H ** 3
        dev = self.density_grid.device
        st = _lib.stream()
        tmp_grid = torch.full_like(self.density_grid, -1.0)
        amp = torch.amp.autocast("cuda", enabled=self.opt.fp16)

        for cas in range(self.cascade):
            bound = float(min(2 ** cas, self.bound))
            if self.iter_density < 16:
                n = H3
                noise = torch.rand(n, 3, device=dev)
                xyzs = torch.empty(n, 3, device=dev)
                indices = torch.empty(n, dtype=torch.int32, device=dev)
                _lib.call("ngp_occ_sample_positions", None, _lib.ptr(noise), n, H, bound, _lib.ptr(xyzs), _lib.ptr(indices), st)
            else:
                # H^3/4 uniform cells + H^3/4 random OCCUPIED cells (renderer.py:853-876 does randint / nonzero / randint /
                # morton3D_invert / cat, with a host sync on the size of the occupied list): the occupied ids are compacted on the
                # device, one kernel turns 6 uniforms per sample into (cell, jittered position); the count never leaves the GPU.
                # With no occupied cell the second half samples uniform cells as well.
                n = 2 * (H3 // 4)
                u = torch.rand(n, 6, device=dev)
                xyzs = torch.empty(n, 3, device=dev)
                indices = torch.empty(n, dtype=torch.int32, device=dev)
                occ_list = torch.empty(H3, dtype=torch.int32, device=dev)
                occ_count = torch.empty(1, dtype=torch.int32, device=dev)
                ws = torch.empty((H3 + 4095) // 4096, dtype=torch.int32, device=dev)
                _lib.call("ngp_occ_sample_partial", _lib.ptr(self.density_grid[cas]), H, bound, _lib.ptr(u), n, _lib.ptr(occ_list),
                          _lib.ptr(occ_count), _lib.ptr(ws), _lib.ptr(xyzs), _lib.ptr(indices), st)
            with amp:
                sigmas = self.density(xyzs)["sigma"].reshape(-1).detach().float().contiguous()
            _lib.call("ngp_occ_scatter_sigmas", _lib.ptr(indices), _lib.ptr(sigmas), n, _lib.ptr(tmp_grid[cas]), st)

        accum = torch.zeros(2, dtype=torch.float64, device=dev)
        mean = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.call("ngp_occ_ema_update", _lib.ptr(self.density_grid), _lib.ptr(tmp_grid), self.cascade * H3, float(decay),
                  _lib.ptr(accum), _lib.ptr(mean), st)
        self._mean_density = mean
        self.iter_density += 1
        self.density_bitfield = raymarching.packbits(self.density_grid.detach(), (mean, float(self.density_thresh)),
                                                     self.density_bitfield)
        return None
