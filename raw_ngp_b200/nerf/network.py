"""NeRFNetwork -- hash grid + tiny MLPs (nerf/network.py:12-184) on the B200 operators.

Same module / parameter names as the reference so that checkpoints keep their keys:
grid_encoder.embeddings, grid_encoder.offsets, grid_mlp.net.{0,1,2}.weight, view_mlp.net.{0,1,2}.weight
(SURVEY.md 5.4).  The camera pose optimizer (barf/) is outside the hot path; the BARF / BAA-NGP feature annealing
that lives inside common_forward (network.py:77-109) is kept.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..activation import trunc_exp
from ..encoding import get_encoder
from ..ffmlp import fused_mlp
from .. import field as _field
from .renderer import NeRFRenderer


class MLP(nn.Module):
    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, opt, bias=True):
        super().__init__()
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.dim_hidden = dim_hidden
        self.num_layers = num_layers
        self.opt = opt
        self.net = nn.ModuleList([
            nn.Linear(self.dim_in if l == 0 else self.dim_hidden,
                      self.dim_out if l == num_layers - 1 else self.dim_hidden, bias=bias)
            for l in range(num_layers)])

    def _fusable(self, x):
        # The tcgen05 kernel computes in fp16 with fp32 accumulation, i.e. what the nn.Linear stack does under autocast
        # (renderer.py:546).  fp32 runs (--fp16 off) and softplus hidden activations use the library path below.
        return (x.is_cuda and x.dim() == 2 and self.opt.internal_activation == "relu" and self.num_layers <= 4
                and all(l.bias is None for l in self.net) and (x.dtype == torch.float16 or torch.is_autocast_enabled("cuda"))
                and max(self.dim_in, self.dim_hidden, self.dim_out) <= 128)

    def forward(self, x):
        if self._fusable(x):
            return fused_mlp(x, *[l.weight for l in self.net])
        for l in range(self.num_layers):
            x = self.net[l](x)
            if l != self.num_layers - 1:
                if self.opt.internal_activation == "relu":
                    x = F.relu(x, inplace=True)
                if self.opt.internal_activation == "softplus":
                    x = F.softplus(x, beta=self.opt.beta, threshold=20)
        return x


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt):
        super().__init__(opt)
        self.annealing = 0.0
        if opt.pose_opt != "none":      # network.py:42-44
            if getattr(opt, "pose_optimizer_factory", None) is not None:
                self.pose_optimizer = opt.pose_optimizer_factory(opt)
            elif getattr(opt, "num_cameras", 0) > 0:
                from ..pose import CameraOptimizer
                self.pose_optimizer = CameraOptimizer(num_cameras=opt.num_cameras, device=opt.device, opt=opt)

        self.level_dim = 2
        self.grid_encoder, self.grid_in_dim = get_encoder(
            "hashgrid", input_dim=3, level_dim=self.level_dim, num_levels=16, log2_hashmap_size=self.opt.hashmap_size,
            desired_resolution=self.opt.hashgrid_resolution * self.bound)
        self.grid_mlp = MLP(self.grid_in_dim, 16, 64, 3, opt, bias=False)

        self.view_encoder, self.view_in_dim = get_encoder("sh", input_dim=3, degree=4)
        ldir_dim = self.view_in_dim if self.opt.rfield else 0
        self.view_mlp = MLP(15 + self.view_in_dim + ldir_dim, 3, 64 + ldir_dim, 3, opt, bias=False)

        if not self.opt.cuda_ray:       # the two proposal fields of the sampling path (network.py:59-72): 5-level grids + 2-layer MLPs
            self.prop_encoders = nn.ModuleList()
            self.prop_mlp = nn.ModuleList()
            for res in (128, 256):
                enc, dim = get_encoder("hashgrid", input_dim=3, level_dim=2, num_levels=5, log2_hashmap_size=17, desired_resolution=res)
                self.prop_encoders.append(enc)
                self.prop_mlp.append(MLP(dim, 1, 16, 2, opt, bias=False))

    def _annealing_window(self, L, device):
        start, end = self.opt.start_annealing, self.opt.end_annealing
        k = torch.arange(L, dtype=torch.float32, device=device)
        if end == 0:
            end = 1e-12
        alpha = (self.annealing - start) / (end - start) * L
        return (1 - (alpha - k).clamp_(min=0, max=1).mul_(np.pi).cos_()) / 2

    # ---- fused path (csrc/field.cu): same maths as the op-by-op path below, four kernels instead of ~60 ops ----------
    FUSED = True

    def _fused_eligible(self, *tensors):
        enc = self.grid_encoder
        return (self.FUSED and enc.embeddings.is_cuda and enc.embeddings.dtype == torch.float16 and enc.level_dim == 2
                and enc.input_dim == 3 and enc.num_levels % 8 == 0 and self.opt.internal_activation == "relu"
                and self.opt.density_activation in ("clamped_exp", "softplus")
                and self.opt.color_activation in _field.COLOR_ACT and self.opt.pose_opt in ("none", "barf")
                and torch.is_autocast_enabled("cuda") and self.grid_mlp.num_layers == 3 and self.view_mlp.num_layers == 3
                and not any(t is not None and t.requires_grad for t in tensors))

    def _fast_infer_args(self, rays_ldir, shading):
        """Inference loop of the renderer: the whole field as one launch per iteration, no tensors allocated (renderer.py)."""
        import ctypes
        from .. import _lib
        from ..ffmlp import _pad16, _ptr_array
        enc = self.grid_encoder
        if not (self.FUSED and self.opt.fp16 and (rays_ldir is not None) == bool(self.opt.rfield) and enc.embeddings.is_cuda
                and enc.embeddings.dtype == torch.float16 and enc.level_dim == 2 and enc.input_dim == 3 and enc.num_levels % 8 == 0
                and self.opt.internal_activation == "relu" and self.opt.density_activation in ("clamped_exp", "softplus")
                and self.opt.color_activation in _field.COLOR_ACT and self.opt.pose_opt in ("none", "barf")
                and self.grid_mlp.num_layers == 3 and self.view_mlp.num_layers == 3):
            return None
        gw, vw = [l.weight for l in self.grid_mlp.net], [l.weight for l in self.view_mlp.net]
        p1 = [_pad16(d) for d in [gw[0].shape[1]] + [w.shape[0] for w in gw]]
        p2 = [_pad16(d) for d in [vw[0].shape[1]] + [w.shape[0] for w in vw]]
        if not _field._ws_fwd_ok(p1, p2, enc.num_levels):
            return None
        w1 = [_field._pad_weight(w, p1[i + 1], p1[i]) for i, w in enumerate(gw)]
        w2 = [_field._pad_weight(w, p2[i + 1], p2[i]) for i, w in enumerate(vw)]
        a1, a2 = _ptr_array(w1), _ptr_array(w2)
        c1, c2 = (ctypes.c_uint32 * 4)(*p1), (ctypes.c_uint32 * 4)(*p2)
        S, H, L, gt, ac, ip = _field._grid_scalars(enc)
        fw = self._feat_weights(enc.embeddings.device)
        dens, col, beta, bound = self._density_act(), _field.COLOR_ACT[self.opt.color_activation], float(self.opt.beta), float(self.bound)
        keep = (w1, w2, fw)      # the launch arguments below hold raw pointers into these
        # light stage (rfield): ONE light direction per frame, repeated for every sample (renderer.py:605)
        ld_vec = rays_ldir.reshape(-1, 3)[:1].float() if self.opt.rfield else None
        ld_rows = {}

        def field(xyzs, dirs, M, sigmas, rgbs, st, m_dev=None, _keep=keep):
            # m_dev: device address of the live row count (<= M), see NeRFRenderer._march_composite_loop_fast
            ldirs = None
            if ld_vec is not None:
                if M not in ld_rows:
                    ld_rows[M] = ld_vec.to(xyzs.device).expand(M, 3).contiguous()
                ldirs = ld_rows[M]
            _lib.call("ngp_field_forward_full", _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ldirs), _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets),
                      _lib.ptr(fw), bound, S, H, L, gt, ac, ip, a1, c1, a2, c2, M, m_dev, dens, beta, col, None, None, None, None,
                      _lib.ptr(sigmas), _lib.ptr(rgbs), None, st)
        return field

    def _feat_weights_host(self):
        """The BARF window of _feat_weights as a numpy array (fp32 arithmetic like the torch expression): lets the captured
        training step refresh it with one small host-to-device copy instead of ten elementwise kernels."""
        if self.opt.pose_opt != "barf":
            return None
        L = self.grid_mlp.dim_out
        start, end = np.float32(self.opt.start_annealing), np.float32(self.opt.end_annealing)
        if end == 0:
            end = np.float32(1e-12)
        k = np.arange(L, dtype=np.float32)
        alpha = np.float32((np.float32(self.annealing) - start) / (end - start) * np.float32(L))
        w = (np.float32(1) - np.cos(np.clip(alpha - k, 0, 1).astype(np.float32) * np.float32(np.pi), dtype=np.float32)) / np.float32(2)
        w = np.repeat(w.astype(np.float32), self.level_dim)
        w[0:2] = 1
        return w

    def _feat_weights(self, device):
        if self.opt.pose_opt != "barf":
            return None
        w = self._annealing_window(self.grid_mlp.dim_out, device).repeat_interleave(self.level_dim)
        w[0:2] = 1
        return w.contiguous()

    def _density_act(self):
        return _field.DENSITY_ACT["clamped_exp" if self.opt.density_activation == "clamped_exp" else "softplus"]

    def common_forward(self, x):
        f = self.grid_encoder(x, bound=self.bound)
        if self.opt.pose_opt == "baangp":      # network.py:77-97
            L = self.grid_mlp.dim_out - 1
            weight = self._annealing_window(L, f.device)
            weights = torch.cat([torch.ones(self.level_dim, device=f.device), weight.repeat_interleave(self.level_dim)])
            assert f.shape[-1] == len(weights)
            available_features = f[..., weights > 0]
            assert len(available_features) > 0, "no features are selected!"
            coarse_features = available_features[..., -self.level_dim:]
            coarse_f = coarse_features.repeat(1, L + 1)
            weights[0:2] = 1
            f = f * weights + coarse_f * (1 - weights)
        if self.opt.pose_opt == "barf":        # network.py:99-109
            L = self.grid_mlp.dim_out
            weights = self._annealing_window(L, f.device).repeat_interleave(self.level_dim)
            weights[0:2] = 1
            f = f * weights

        f = self.grid_mlp(f)
        if self.opt.density_activation == "clamped_exp":
            sigma = trunc_exp(f[..., 0])
        else:
            sigma = F.softplus(f[..., 0], beta=self.opt.beta, threshold=20)
        return sigma, f[..., 1:]

    def forward(self, x, d, ld=None, **kwargs):
        # x [N, 3] in [-bound, bound], d [N, 3] unit view directions, ld [N, 3] unit light directions (rfield)
        if x.dim() == 3:        # [N, T, 3] samples of the proposal path: one flat batch through the same kernels
            out = self.forward(x.reshape(-1, 3), d.reshape(-1, 3), None if ld is None else ld.reshape(-1, 3), **kwargs)
            return {"sigma": out["sigma"].view(*x.shape[:-1]), "color": out["color"].view(*x.shape[:-1], 3)}
        if x.dim() == 2 and self._fused_eligible(x, d, ld) and (ld is not None) == bool(self.opt.rfield):
            sigma, color = _field.fused_field(
                x, d, ld if self.opt.rfield else None, self.grid_encoder, [l.weight for l in self.grid_mlp.net],
                [l.weight for l in self.view_mlp.net], self.bound, self._density_act(), self.opt.beta,
                _field.COLOR_ACT[self.opt.color_activation], self._feat_weights(x.device))
            return {"sigma": sigma, "color": color}
        sigma, feat = self.common_forward(x)
        d = self.view_encoder(d)
        if self.opt.rfield:
            ld = self.view_encoder(ld)
            color = self.view_mlp(torch.cat([feat, d, ld], dim=-1))
        else:
            color = self.view_mlp(torch.cat([feat, d], dim=-1))
        if self.opt.color_activation == "exp":
            color = torch.exp(color - 5.0)
        if self.opt.color_activation == "sigmoid":
            color = torch.sigmoid(color)
        if self.opt.color_activation == "clamped_exp":
            color = torch.clamp(torch.exp(color - 5.0), max=5.0)
        return {"sigma": sigma, "color": color}

    def density(self, x, proposal=-1):
        if proposal >= 0 and not self.opt.cuda_ray and proposal < len(self.prop_encoders):      # network.py:146-149
            f = self.prop_encoders[proposal](x, bound=self.bound)
            lead = f.shape[:-1]
            return {"sigma": trunc_exp(self.prop_mlp[proposal](f.reshape(-1, f.shape[-1])).view(*lead))}
        if x.dim() == 2 and not torch.is_grad_enabled() and self._fused_eligible(x):
            return {"sigma": _field.density_only(self.grid_encoder, [l.weight for l in self.grid_mlp.net], x, self.bound,
                                                 self._density_act(), self.opt.beta, self._feat_weights(x.device))}
        sigma, _ = self.common_forward(x)
        return {"sigma": sigma}

    def apply_total_variation(self, w):
        self.grid_encoder.grad_total_variation(w)

    def apply_weight_decay(self, w):
        self.grid_encoder.grad_weight_decay(w)

    def update_annealing(self, new_value):
        self.annealing = new_value

    def get_params(self, lr):
        groups = [
            {"params": self.grid_encoder.parameters(), "lr": lr},
            {"params": self.grid_mlp.parameters(), "lr": lr},
            {"params": self.view_mlp.parameters(), "lr": lr},
        ]
        if not self.opt.cuda_ray:
            groups += [{"params": self.prop_encoders.parameters(), "lr": lr}, {"params": self.prop_mlp.parameters(), "lr": lr}]
        return groups
