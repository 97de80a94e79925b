"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference grid encoder.

Follows gridencoder/src/gridencoder.cu: get_grid_index/fast_hash :45-79, kernel_grid :82-249,
kernel_grid_backward :252-349, kernel_input_backward :352-378, kernel_grad_tv :525-631, kernel_grad_wd :670-703,
and the wrapper gridencoder/grid.py:27-95 (layouts: outputs [B, L*C], dy_dx [B, L*D*C]).

fp32 arithmetic of the GPU is emulated where it decides the result: a*b+c contracted by nvcc into one FFMA is
evaluated in float64 and rounded once (exact for 24-bit operands); `half=True` reproduces the reference's at::Half
accumulation (product rounded to half, half+half evaluated in fp32 and rounded, gridencoder.cu:168,191).
Pinned by tests/golden/grid_*.npz, which are outputs of the reference's own kernels (tools/make_golden.py).
"""
import numpy as np

PRIMES = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint64)
f32 = np.float32


def level_resolutions(L, per_level_scale, H):
    """ceil(exp2f(level * S) * H) in fp32 (gridencoder.cu:133) with S = (float)log2(per_level_scale) (grid.py:38)."""
    S = f32(np.log2(per_level_scale))
    lv = np.arange(L, dtype=np.float32)
    return np.ceil(np.exp2(lv * S, dtype=np.float32) * f32(H)).astype(np.uint32)


def table_offsets(D, L, per_level_scale, H, log2_hashmap_size):
    """grid.py:124-134 (float64 on the host)."""
    max_params = 2 ** log2_hashmap_size
    offs, o = [], 0
    for i in range(L):
        res = int(np.ceil(H * per_level_scale ** i))
        n = min(max_params, res ** D)
        n = int(np.ceil(n / 8) * 8)
        offs.append(o)
        o += n
    offs.append(o)
    return np.array(offs, dtype=np.int32)


def grid_index(gridtype, hashmap_size, res, pos):
    """Entry row of integer positions pos [B, D] (uint32 arithmetic).  gridencoder.cu:61-79."""
    B, D = pos.shape
    pos = pos.astype(np.uint64)
    stride = 1
    index = np.zeros(B, dtype=np.uint64)
    for d in range(D):
        if stride <= hashmap_size:
            index = (index + pos[:, d] * np.uint64(stride)) & np.uint64(0xFFFFFFFF)
            stride = (stride * int(res)) & 0xFFFFFFFF
    if gridtype == 0 and stride > hashmap_size:
        index = np.zeros(B, dtype=np.uint64)
        for d in range(D):
            index ^= (pos[:, d] * PRIMES[d]) & np.uint64(0xFFFFFFFF)
    return (index % np.uint64(hashmap_size)).astype(np.int64)


def _locate(x, res, align_corners, interp):
    """gridencoder.cu:140-160 -> base corner [B,D] uint32, frac [B,D] f32, dfrac [B,D] f32."""
    x64 = x.astype(np.float64)
    if align_corners:
        pos = (x * f32(res - 1)).astype(np.float32)
        base = np.minimum(np.floor(pos).astype(np.int64), res - 2)
    else:
        pos = (x64 * float(res) - 0.5).astype(np.float32)  # one FFMA on the GPU
        pos = np.minimum(np.maximum(pos, f32(0)), f32(res - 1))
        base = np.floor(pos).astype(np.int64)
    pos = (pos - base.astype(np.float32)).astype(np.float32)
    if interp == 1:
        p64 = pos.astype(np.float64)
        dfrac = ((f32(6) * pos).astype(np.float32) * (f32(1) - pos)).astype(np.float32)
        inner = (3.0 - 2.0 * p64).astype(np.float32)          # FFMA
        pos = ((pos * pos).astype(np.float32) * inner).astype(np.float32)
    else:
        dfrac = np.ones_like(pos)
    return base, pos, dfrac


def _h(v):
    """round fp32 -> fp16 -> fp32"""
    return v.astype(np.float16).astype(np.float32)


def _acc(acc, w, g, half):
    if half:
        prod = _h((w * g).astype(np.float32))
        return _h(acc + prod)
    return (acc.astype(np.float64) + w.astype(np.float64) * g.astype(np.float64)).astype(np.float32)  # FFMA


def forward(inputs, table, offsets, per_level_scale, H, gridtype=0, align_corners=False, interp=0, max_level=None,
            calc_dy_dx=False, half=False, resolutions=None):
    """inputs [B,D] f32 in [0,1]; table [sO,C] (f32 values; pass the fp16-rounded table for half=True).
    Returns outputs [B, L*C] f32 and dy_dx [B, L*D*C] f32 or None."""
    inputs = np.ascontiguousarray(inputs, dtype=np.float32)
    table = np.asarray(table, dtype=np.float32)
    B, D = inputs.shape
    L = len(offsets) - 1
    C = table.shape[1]
    max_level = L if max_level is None else min(max_level, L)
    if resolutions is None:
        resolutions = level_resolutions(L, per_level_scale, H)
    out = np.zeros((B, L, C), dtype=np.float32)
    dy_dx = np.zeros((B, L, D, C), dtype=np.float32) if calc_dy_dx else None
    inb = np.all((inputs >= 0) & (inputs <= 1), axis=1)
    for l in range(max_level):
        res = int(resolutions[l])
        hs = int(offsets[l + 1] - offsets[l])
        lvl = table[offsets[l]:offsets[l + 1]]
        base, frac, dfrac = _locate(inputs, res, align_corners, interp)
        vals = []
        for k in range(1 << D):
            p = np.stack([np.minimum(base[:, d] + 1, res - 1) if (k >> d) & 1 else base[:, d] for d in range(D)], axis=1)
            vals.append(lvl[grid_index(gridtype, hs, res, np.maximum(p, 0))])
        acc = np.zeros((B, C), dtype=np.float32)
        for k in range(1 << D):
            w = np.ones(B, dtype=np.float32)
            for d in range(D):
                w = (w * (frac[:, d] if (k >> d) & 1 else (f32(1) - frac[:, d]))).astype(np.float32)
            acc = _acc(acc, w[:, None], vals[k], half)
        out[:, l] = np.where(inb[:, None], acc, 0)
        if calc_dy_dx:
            scale = f32(res - 1 if align_corners else res)
            for g in range(D):
                dacc = np.zeros((B, C), dtype=np.float32)
                others = [d for d in range(D) if d != g]
                for j in range(1 << (D - 1)):
                    w = np.full(B, scale, dtype=np.float32)
                    lo = 0
                    for nd, d in enumerate(others):
                        if (j >> nd) & 1:
                            w = (w * frac[:, d]).astype(np.float32)
                            lo |= 1 << d
                        else:
                            w = (w * (f32(1) - frac[:, d])).astype(np.float32)
                    hi = lo | (1 << g)
                    diff = (vals[hi] - vals[lo]).astype(np.float32)
                    if half:
                        diff = _h(diff)
                    wd = (w[:, None] * diff).astype(np.float32)
                    dacc = _acc(dacc, wd, dfrac[:, g:g + 1], half)
                dy_dx[:, l, g] = np.where(inb[:, None], dacc, 0)
    return out.reshape(B, L * C), (dy_dx.reshape(B, L * D * C) if calc_dy_dx else None)


def backward(grad, inputs, table_shape, offsets, per_level_scale, H, gridtype=0, align_corners=False, interp=0,
             max_level=None, half=False, resolutions=None):
    """Table gradient [sO, C] (float64 accumulation = order-free ideal of the reference's atomics).
    half=True rounds each contribution to fp16 first, as the reference does before its half2 atomicAdd."""
    inputs = np.ascontiguousarray(inputs, dtype=np.float32)
    B, D = inputs.shape
    L = len(offsets) - 1
    C = table_shape[1]
    max_level = L if max_level is None else min(max_level, L)
    if resolutions is None:
        resolutions = level_resolutions(L, per_level_scale, H)
    grad = np.asarray(grad, dtype=np.float32).reshape(B, L, C)
    gt = np.zeros(table_shape, dtype=np.float64)
    inb = np.all((inputs >= 0) & (inputs <= 1), axis=1)
    for l in range(max_level):
        res = int(resolutions[l])
        hs = int(offsets[l + 1] - offsets[l])
        base, frac, _ = _locate(inputs, res, align_corners, interp)
        for k in range(1 << D):
            p = np.stack([np.minimum(base[:, d] + 1, res - 1) if (k >> d) & 1 else base[:, d] for d in range(D)], axis=1)
            rows = grid_index(gridtype, hs, res, np.maximum(p, 0)) + int(offsets[l])
            w = np.ones(B, dtype=np.float32)
            for d in range(D):
                w = (w * (frac[:, d] if (k >> d) & 1 else (f32(1) - frac[:, d]))).astype(np.float32)
            contrib = (w[:, None] * grad[:, l]).astype(np.float32)
            if half:
                contrib = _h(contrib)
            np.add.at(gt, rows[inb], contrib[inb].astype(np.float64))
    return gt


def input_backward(grad, dy_dx, B, D, C, L):
    """grad_inputs[b,d] = sum_{l,c} grad[b,l,c] * dy_dx[b,l,d,c]   (gridencoder.cu:352-378), float64 sum."""
    g = np.asarray(grad, dtype=np.float64).reshape(B, L, 1, C)
    j = np.asarray(dy_dx, dtype=np.float64).reshape(B, L, D, C)
    return (g * j).sum(axis=(1, 3))
