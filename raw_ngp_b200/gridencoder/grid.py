"""GridEncoder -- multiresolution hash / tiled grid encoding, B200 backend.

Host-side mirror of the reference operator surface (gridencoder/grid.py:24-211): same class, function and
argument names, same autograd / autocast contract.  The native work goes through the C ABI of
libngp_b200.so (include/ngp_b200.h); there is no other backend.

Differences that are deliberate:
  * the kernel writes / reads the [B, L*C] layout directly (no [L,B,C] staging + permute copies,
    reference grid.py:49,63,81);
  * no dy_dx [B, L*D*C] tensor is saved for backward; the input gradient is recomputed from the table inside
    the backward kernel (reference grid.py:54-55, gridencoder.cu:352-378).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}

# fp16 tables: reproduce the reference's half-precision accumulation bit for bit (gridencoder.cu:168,191).
# Set to False for fp32 accumulation with a single final rounding (more accurate, not bit-identical).
REFERENCE_ROUNDING = True


def _flags():
    return _lib.NGP_GRID_REF_ROUNDING if REFERENCE_ROUNDING else 0


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0, max_level=None, grad_sink=None):
        # inputs [B, D] float in [0, 1]; embeddings [sO, C]; offsets [L + 1] int32  ->  [B, L * C] in table dtype
        # grad_sink (extension, optional): persistent [sO, C] buffer the table gradient is ACCUMULATED into instead of
        # returning a fresh zero-filled tensor (no 24-48 MB memset + allocation per step; see raw_ngp_b200/trainer.py)
        _lib.require_cuda(inputs, embeddings, offsets)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        embeddings = embeddings.contiguous()
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")
        offsets = offsets.contiguous()

        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        max_level = L if max_level is None else min(int(max_level), L)
        dt = _lib.dtype_id(embeddings.dtype)

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=embeddings.dtype)
        _lib.call("ngp_grid_encode_forward", _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                  _lib.ptr(outputs), B, D, C, L, max_level, S, H, None, int(gridtype), int(bool(align_corners)),
                  int(interpolation), dt, _flags(), _lib.stream())

        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dims = (B, D, C, L, S, H, int(gridtype), int(interpolation), max_level)
        ctx.align_corners = bool(align_corners)
        ctx.calc_grad_inputs = bool(calc_grad_inputs)
        ctx.grad_sink = grad_sink
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, max_level = ctx.dims

        grad = grad.contiguous()
        if grad.dtype != embeddings.dtype:
            grad = grad.to(embeddings.dtype)
        sink = ctx.grad_sink
        if sink is not None:
            if sink.shape != embeddings.shape or sink.dtype != embeddings.dtype or not sink.is_contiguous():
                raise RuntimeError("grad_sink must be a contiguous tensor with the table's shape and dtype")
            grad_embeddings = sink
        else:
            grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros(B, D, device=inputs.device, dtype=torch.float32) if ctx.calc_grad_inputs else None

        _lib.call("ngp_grid_encode_backward", _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings),
                  _lib.ptr(offsets), _lib.ptr(grad_embeddings), B, D, C, L, max_level, S, H, _lib.ptr(grad_inputs),
                  gridtype, int(ctx.align_corners), interpolation, _lib.dtype_id(embeddings.dtype), _flags(),
                  _lib.stream())
        return grad_inputs, (None if sink is not None else grad_embeddings), None, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def grid_encode_with_jacobian(inputs, embeddings, offsets, per_level_scale, base_resolution, gridtype=0,
                              align_corners=False, interpolation=0, max_level=None):
    """Forward that also materialises dy_dx [B, L*D*C] in the reference's layout (gridencoder.cu:207).
    Not used by the autograd path; kept so the kernel can be compared with the reference's dy_dx."""
    inputs = inputs.contiguous().float()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    max_level = L if max_level is None else min(int(max_level), L)
    outputs = torch.empty(B, L * C, device=inputs.device, dtype=embeddings.dtype)
    dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
    _lib.call("ngp_grid_encode_forward", _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(outputs),
              B, D, C, L, max_level, S, int(base_resolution), _lib.ptr(dy_dx), int(gridtype), int(bool(align_corners)),
              int(interpolation), _lib.dtype_id(embeddings.dtype), _flags(), _lib.stream())
    return outputs, dy_dx


def level_table_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size):
    """Row offset of every level inside the embedding table, plus the total as last entry.

    Level i has resolution ceil(base * scale^i) (float64 on the host, as the reference computes it, grid.py:124-134) and
    min(2^log2_hashmap_size, resolution^D) rows, rounded up to a multiple of 8 so that every level starts 32-byte aligned."""
    cap = 2 ** log2_hashmap_size
    rows = []
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        rows.append(int(np.ceil(min(cap, res ** input_dim) / 8) * 8))
    return [int(v) for v in np.concatenate([[0], np.cumsum(rows, dtype=np.int64)])]


class GridEncoder(nn.Module):
    """Multiresolution hash / tiled grid with trainable fp32 / fp16 / bf16 embeddings (GridEncoder of the reference)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:      # geometric progression from base_resolution up to desired_resolution
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.log2_hashmap_size, self.max_params = log2_hashmap_size, 2 ** log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, _gridtype_to_id[gridtype]
        self.interpolation, self.interp_id = interpolation, _interp_to_id[interpolation]
        self.align_corners = align_corners

        table = level_table_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer("offsets", torch.tensor(table, dtype=torch.int32))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(table[-1], level_dim))
        self.grad_sink = None   # set by raw_ngp_b200.trainer: table gradients accumulate here, embeddings.grad stays None
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(self, inputs, bound=1, max_level=None):
        # [..., input_dim] in [-bound, bound] -> unit cube -> [..., num_levels * level_dim]
        lead = inputs.shape[:-1]
        unit = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        feats = grid_encode(unit, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, unit.requires_grad,
                            self.gridtype_id, self.align_corners, self.interp_id, max_level, self.grad_sink)
        return feats.view(*lead, self.output_dim)

    def _table_grad(self):
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        return self.embeddings.grad

    @torch.amp.autocast("cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """Adds weight * d TV / d table to embeddings.grad, TV sampled at `inputs` (or at B uniform points)."""
        table = self.embeddings
        if inputs is None:
            points = torch.rand(B, self.input_dim, device=table.device)
        else:
            points = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        grad = self._table_grad()
        points = points.to(table.dtype).contiguous()
        _lib.call("ngp_grid_grad_total_variation", _lib.ptr(points), _lib.ptr(table), _lib.ptr(grad), _lib.ptr(self.offsets),
                  float(weight), points.shape[0], self.input_dim, table.shape[1], self.offsets.shape[0] - 1,
                  float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id, int(self.align_corners),
                  _lib.dtype_id(table.dtype), _lib.stream())

    @torch.amp.autocast("cuda", enabled=False)
    def grad_weight_decay(self, weight=0.1):
        """Adds the L2 penalty's gradient, normalised per level, to embeddings.grad."""
        table, grad = self.embeddings, self._table_grad()
        _lib.call("ngp_grid_grad_weight_decay", _lib.ptr(table), _lib.ptr(grad), _lib.ptr(self.offsets), float(weight),
                  table.shape[0], table.shape[1], self.offsets.shape[0] - 1, _lib.dtype_id(table.dtype), _lib.stream())
