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
import torch
import torch.distributed as dist

from . import _lib, parallel


class FusedAdam:
    """Adam over a list of (master fp32, low-precision copy or None, grad buffer) groups via ngp_fused_adam."""

    def __init__(self, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.0):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.groups = []
        self.step_count = 0

    def add_group(self, master, grad, param_lp=None):
        assert master.dtype == torch.float32 and master.is_contiguous() and grad.is_contiguous()
        self.groups.append(dict(master=master, lp=param_lp, grad=grad, m=torch.zeros_like(master),
                                v=torch.zeros_like(master)))

    def step(self, inv_scale, found_inf, zero_grad=True, lr=None):
        self.step_count += 1
        lr = self.lr if lr is None else lr
        st = _lib.stream()
        for g in self.groups:
            lp = g["lp"]
            _lib.call("ngp_fused_adam", _lib.ptr(g["master"]), _lib.ptr(lp), _lib.dtype_id(lp.dtype) if lp is not None else 0,
                      _lib.ptr(g["grad"]), _lib.dtype_id(g["grad"].dtype), _lib.ptr(g["m"]), _lib.ptr(g["v"]),
                      g["master"].numel(), float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                      float(self.weight_decay), self.step_count, _lib.ptr(inv_scale), _lib.ptr(found_inf),
                      int(zero_grad), st)


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
        mlp_params = [p for n, p in model.named_parameters() if not n.startswith("grid_encoder.")]
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
