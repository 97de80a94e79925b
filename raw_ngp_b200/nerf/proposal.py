"""The proposal-network sampling path of the reference renderer (`cuda_ray=False`; nerf/renderer.py:22-136, 405-513): hierarchical
sampling along each ray with two small proposal fields (mip-NeRF 360 style), volume rendering of the last level, and the
inter-level / distortion losses.  SURVEY 8(f) row 4.

Everything here is dense torch arithmetic on [N, T] tensors around the B200 operators (GridEncoder / SHEncoder / fused MLP do
the per-sample work); there is no kernel of its own.  Semantics follow the reference function by function:

  spacing / inverse spacing      renderer.py:196-197   s(t) = t/2 (t < 1), 1 - 1/(2t);   s^-1(u) = 2u (u < 1/2), 1/(2 - 2u)
  contract                       renderer.py:79-87     L-inf contraction: |x|_inf >= 1 -> the dominant axis becomes (2 - 1/m) sign,
                                                       the others x / m  (NOT the all-axes scaling of the CUDA marcher)
  resample_bins (sample_pdf)     renderer.py:103-136   stratified inverse-CDF sampling of T + 1 bin edges from (bins, weights + 0.01)
  weights_from_sigmas            renderer.py:470-485   alpha_i = 1 - exp(-delta_i sigma_i), T_i = exp(-sum_{j<i} delta_j sigma_j)
  interlevel_loss (proposal_loss) renderer.py:50-76    sum_levels mean max(0, w - bound(w; proposal histogram))^2 / (w + 1e-8)
  distortion_loss                renderer.py:22-33     the reference calls torch_efficient_distloss.eff_distloss (third party, not
                                                       installed here, unpinned): restated from its published definition,
                                                       sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 delta_i, in O(T) with prefix sums
"""
import torch


def spacing(t):
    return torch.where(t < 1, t / 2, 1 - 1 / (2 * t))


def spacing_inv(u):
    return torch.where(u < 0.5, 2 * u, 1 / (2 - 2 * u))


@torch.amp.autocast("cuda", enabled=False)
def contract(x):
    lead, C = x.shape[:-1], x.shape[-1]
    flat = x.reshape(-1, C)
    mag, axis = flat.abs().max(dim=1, keepdim=True)
    scale = (1 / mag).expand(-1, C).clone()
    scale.scatter_(1, axis, (2 - 1 / mag) / mag)
    return torch.where(mag < 1, flat, flat * scale).view(*lead, C)


def resample_bins(bins, weights, n_edges, perturb=False):
    """bins [N, T0 + 1] edges in [0, 1], weights [N, T0] -> n_edges new edges per ray, drawn from the piecewise-constant density."""
    N, T0 = weights.shape
    w = weights + 0.01
    cdf = torch.cumsum(w / w.sum(-1, keepdim=True), dim=-1).clamp(max=1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.5 / n_edges, 1 - 0.5 / n_edges, steps=n_edges).to(weights.device).expand(N, n_edges)
    if perturb:
        u = u + (torch.rand_like(u) - 0.5) / n_edges
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp(0, T0)
    hi = hi.clamp(0, T0)
    c0, c1 = cdf.gather(-1, lo), cdf.gather(-1, hi)
    b0, b1 = bins.gather(-1, lo), bins.gather(-1, hi)
    frac = torch.nan_to_num((u - c0) / (c1 - c0)).clamp(0, 1)
    return b0 + frac * (b1 - b0)


def weights_from_sigmas(real_bins, sigmas, opaque_last=False):
    deltas = real_bins[..., 1:] - real_bins[..., :-1]
    ds = deltas * sigmas
    if opaque_last:         # background == 'last_sample' (renderer.py:474-476)
        ds = torch.cat([ds[..., :-1], torch.full_like(ds[..., -1:], torch.inf)], dim=-1)
    alphas = 1 - torch.exp(-ds)
    acc = torch.cumsum(ds[..., :-1], dim=-1)
    trans = torch.exp(-torch.cat([torch.zeros_like(acc[..., :1]), acc], dim=-1))
    return (alphas * trans).nan_to_num_(0)


@torch.amp.autocast("cuda", enabled=False)
def interlevel_loss(all_bins, all_weights):
    ref_bins, ref_w = all_bins[-1].detach(), all_weights[-1].detach()
    loss = 0
    for bins, w in zip(all_bins[:-1], all_weights[:-1]):
        # upper bound of the fine weights implied by the proposal histogram: the proposal mass of every bin that overlaps [a, b)
        cw = torch.cat([torch.zeros_like(w[..., :1]), torch.cumsum(w, dim=-1)], dim=-1)
        lo = (torch.searchsorted(bins[..., :-1].contiguous(), ref_bins[..., :-1].contiguous(), right=True) - 1).clamp(0, w.shape[-1] - 1)
        hi = torch.searchsorted(bins[..., 1:].contiguous(), ref_bins[..., 1:].contiguous(), right=True).clamp(0, w.shape[-1] - 1)
        bound = torch.take_along_dim(cw[..., 1:], hi, dim=-1) - torch.take_along_dim(cw[..., :-1], lo, dim=-1)
        loss = loss + ((ref_w - bound).clamp(min=0) ** 2 / (ref_w + 1e-8)).mean()
    return loss


@torch.amp.autocast("cuda", enabled=False)
def distortion_loss(bins, weights):
    d = bins[..., 1:] - bins[..., :-1]
    m = bins[..., :-1] + d / 2
    wm = weights * m
    w_before = torch.cumsum(weights, dim=-1) - weights
    wm_before = torch.cumsum(wm, dim=-1) - wm
    inter = 2 * (wm * w_before - weights * wm_before)          # sum_ij w_i w_j |m_i - m_j|, midpoints are sorted
    intra = weights ** 2 * d / 3
    return (inter + intra).sum(dim=-1).mean()
