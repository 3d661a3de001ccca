"""Densification / pruning and the optimizer-state surgery that goes with it, on the flat parameter buffers.

Same operations, names and argument meaning as ``GaussianModel`` in /root/reference/src/models/gaussian.py:
``add_densification_stats`` (:335-338), ``densify_and_prune`` (:309-333) = ``densify_and_clone`` (:286-307) +
``densify_and_split`` (:249-284) + ``prune_points`` (:188-201), ``reset_opacity`` (:148-151), with the Adam moments
carried along exactly like ``cat_tensors_to_optimizer`` (:203-224, zeros for new Gaussians), ``_prune_optimizer``
(:163-186, rows dropped) and ``replace_tensor_to_optimizer`` (:153-161, zeros).  The reference keeps six tensors and six
optimizer states; here everything lives in ``FlatGaussians`` / ``FlatAdam`` buffers (one contiguous segment per parameter, the
layout the render path, the gradient exchange and the fused Adam use), so a change of N rebuilds the flat buffers once.
Plain torch on whatever device the buffers are on (it runs every ~100 steps); random samples of ``densify_and_split``
come from ``torch.normal`` like the reference's, so identically seeded ranks stay bit-identical.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .dist import PARAM_ORDER, FlatGaussians, param_shapes
from .optim import GROUP_OF, FlatAdam


def inverse_sigmoid(x):
    """src/utils/gaussian_utils.py:199-200"""
    return torch.log(x / (1 - x))


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """src/utils/gaussian_utils.py:278-301 (quaternion normalised inside, (r, x, y, z) order)."""
    q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device, dtype=r.dtype)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


SH_C0 = 0.28209479177387814


def initialize_parameters(points, points_colors, sh_degree: int = 3, isotropic: bool = False, dist2: Optional[torch.Tensor] = None,
                          device=None) -> Dict[str, torch.Tensor]:
    """``GaussianModel.initialize_parameters`` (/root/reference/src/models/gaussian.py:99-127), the one caller of ``distCUDA2``
    (:110): xyz = points; f_dc = RGB2SH(colours) (sh_utils.py:123-124), f_rest = 0; log-scales = log(sqrt(clamp_min(mean squared
    distance to the 3 nearest neighbours, 1e-7))) in every axis (one column when isotropic); identity quaternions (1, 0, 0, 0);
    opacity logit = inverse_sigmoid(0.1).  Returns the six parameters under the names ``FlatGaussians`` uses.
    ``dist2``: the 3-NN statistic when the caller already has it; otherwise ``manus_b200.knn.distCUDA2`` computes it (CUDA)."""
    xyz = torch.as_tensor(points).float()
    if device is not None:
        xyz = xyz.to(device)
    n, dev = xyz.shape[0], xyz.device
    colors = torch.as_tensor(points_colors).float().to(dev)
    if dist2 is None:
        from .knn import distCUDA2

        dist2 = distCUDA2(xyz)
    dist2 = torch.clamp_min(torch.as_tensor(dist2).float().to(dev), 0.0000001)
    log_scale = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 1 if isotropic else 3)
    quat = torch.zeros((n, 4), device=dev)
    quat[:, 0] = 1
    k = (sh_degree + 1) ** 2
    return {"xyz": xyz, "opacity_logit": inverse_sigmoid(0.1 * torch.ones((n, 1), dtype=torch.float, device=dev)),
            "log_scale": log_scale, "quat": quat, "f_dc": ((colors - 0.5) / SH_C0).reshape(n, 1, 3).contiguous(),
            "f_rest": torch.zeros((n, k - 1, 3), device=dev)}


def _segments(flat: FlatGaussians, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Per-parameter views of a flat buffer laid out like ``flat.data``."""
    out, off = {}, 0
    shapes = param_shapes(flat.n, flat.sh_coeffs, flat.isotropic)
    for name in PARAM_ORDER:
        cnt = flat.params[name].numel()
        out[name] = buf[off: off + cnt].view(shapes[name])
        off += cnt
    return out


class GaussianState:
    """Parameters (flat), Adam moments (flat), skin weights and densification statistics of one Gaussian model."""

    def __init__(self, flat: FlatGaussians, opt: Optional[FlatAdam] = None, skin_wts: Optional[torch.Tensor] = None,
                 percent_dense: float = 0.01):
        self.flat, self.opt, self.skin_wts, self.percent_dense = flat, opt, skin_wts, percent_dense
        dev = flat.data.device
        self.xyz_gradient_accum = torch.zeros((flat.n, 1), device=dev)       # training_setup, gaussian.py:129-131
        self.denom = torch.zeros((flat.n, 1), device=dev)
        self.max_radii2D = torch.zeros(flat.n, device=dev)

    # ---- the reference's accessors (gaussian.py:62-82) on the flat segments
    @property
    def n(self) -> int:
        return self.flat.n

    @property
    def get_xyz(self):
        return self.flat.params["xyz"]

    @property
    def get_scaling(self):
        return torch.exp(self.flat.params["log_scale"])

    @property
    def get_opacity(self):
        return torch.sigmoid(self.flat.params["opacity_logit"])

    # ---- statistics
    def add_densification_stats(self, viewspace_point_grad: torch.Tensor, update_filter: torch.Tensor,
                                radii: Optional[torch.Tensor] = None) -> None:
        """gaussian.py:335-338; with ``radii`` also the running maximum of the screen radii that the reference's
        ``density_update`` keeps (src/utils/gaussian_utils.py:461-473).  viewspace_point_grad = means2D.grad [N,3]."""
        if radii is not None:
            self.max_radii2D[update_filter] = torch.max(self.max_radii2D[update_filter], radii[update_filter].to(self.max_radii2D.dtype))
        self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_point_grad[update_filter, :2], dim=-1, keepdim=True)
        self.denom[update_filter] += 1

    def reduce_stats(self, group=None) -> None:
        """Data-parallel runs: statistics are sums / maxima over the ranks' views.  In place with SUM: call it
        ONCE per densification interval, right before densify_and_prune (which resets the statistics); a second call before the
        reset would count every rank's share world_size times."""
        if getattr(self, "_stats_reduced", False):
            raise RuntimeError("reduce_stats() was already called since the statistics were last reset")
        self._stats_reduced = True
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.denom, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=group)

    # ---- structural changes
    def _rebuild(self, params: Dict[str, torch.Tensor], exp_avg: Optional[Dict[str, torch.Tensor]],
                 exp_avg_sq: Optional[Dict[str, torch.Tensor]]) -> None:
        n = params["xyz"].shape[0]
        new = FlatGaussians(n, self.flat.data.device, self.flat.sh_coeffs, self.flat.isotropic)
        for name in PARAM_ORDER:
            new.params[name].copy_(params[name].reshape(new.params[name].shape))
        old = self.flat
        self.flat = new
        if self.opt is not None:
            m, v = torch.zeros_like(new.data), torch.zeros_like(new.data)
            ms, vs = _segments(new, m), _segments(new, v)
            for name in PARAM_ORDER:
                ms[name].copy_(exp_avg[name].reshape(ms[name].shape))
                vs[name].copy_(exp_avg_sq[name].reshape(vs[name].shape))
            self.opt.rebind(new, m, v)
        del old

    def _moments(self):
        if self.opt is None:
            return None, None
        return _segments(self.flat, self.opt.exp_avg), _segments(self.flat, self.opt.exp_avg_sq)

    def prune_points(self, mask: torch.Tensor) -> None:
        """gaussian.py:188-201 with _prune_optimizer (:163-186): rows where ``mask`` is True disappear everywhere."""
        keep = ~mask
        ms, vs = self._moments()
        params = {k: self.flat.params[k][keep] for k in PARAM_ORDER}
        self._rebuild(params, None if ms is None else {k: ms[k][keep] for k in PARAM_ORDER},
                      None if vs is None else {k: vs[k][keep] for k in PARAM_ORDER})
        if self.skin_wts is not None:
            self.skin_wts = self.skin_wts[keep]
        self.xyz_gradient_accum = self.xyz_gradient_accum[keep]
        self.denom = self.denom[keep]
        self.max_radii2D = self.max_radii2D[keep]

    def densification_postfix(self, new: Dict[str, torch.Tensor], new_skin_weights: Optional[torch.Tensor]) -> None:
        """gaussian.py:226-247 with cat_tensors_to_optimizer (:203-224): append rows, zero moments for them, reset statistics."""
        ms, vs = self._moments()
        cat = lambda a, b: torch.cat((a, b.reshape((b.shape[0],) + tuple(a.shape[1:]))), dim=0)
        params = {k: cat(self.flat.params[k], new[k]) for k in PARAM_ORDER}
        zeros = {k: torch.zeros_like(new[k]) for k in PARAM_ORDER}
        self._rebuild(params, None if ms is None else {k: cat(ms[k], zeros[k]) for k in PARAM_ORDER},
                      None if vs is None else {k: cat(vs[k], zeros[k]) for k in PARAM_ORDER})
        if self.skin_wts is not None:
            self.skin_wts = torch.cat((self.skin_wts, new_skin_weights), dim=0)
        dev = self.flat.data.device
        self.xyz_gradient_accum = torch.zeros((self.n, 1), device=dev)
        self.denom = torch.zeros((self.n, 1), device=dev)
        self.max_radii2D = torch.zeros(self.n, device=dev)
        self._stats_reduced = False

    def densify_and_split(self, grads, grad_threshold, scene_extent, N=2) -> None:
        """gaussian.py:249-284"""
        p = self.flat.params
        n_init = self.n
        padded_grad = torch.zeros(n_init, device=grads.device)
        padded_grad[: grads.shape[0]] = grads.squeeze()
        selected = torch.where(grad_threshold <= padded_grad, True, False)
        selected = torch.logical_and(selected, torch.max(self.get_scaling, dim=1).values > self.percent_dense * scene_extent)
        stds = self.get_scaling[selected].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device=stds.device)
        samples = torch.normal(mean=means, std=stds)
        rots = build_rotation(p["quat"][selected]).repeat(N, 1, 1)
        new = {
            "xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self.get_xyz[selected].repeat(N, 1),
            "log_scale": torch.log(self.get_scaling[selected].repeat(N, 1) / (0.8 * N)),
            "quat": p["quat"][selected].repeat(N, 1),
            "f_dc": p["f_dc"][selected].repeat(N, 1, 1),
            "f_rest": p["f_rest"][selected].repeat(N, 1, 1),
            "opacity_logit": p["opacity_logit"][selected].repeat(N, 1),
        }
        new_skin = self.skin_wts[selected].repeat(N, 1) if self.skin_wts is not None else None
        self.densification_postfix(new, new_skin)
        prune_filter = torch.cat((selected, torch.zeros(N * int(selected.sum()), device=selected.device, dtype=torch.bool)))
        self.prune_points(prune_filter)

    def densify_and_clone(self, grads, grad_threshold, scene_extent) -> None:
        """gaussian.py:286-307"""
        p = self.flat.params
        selected = torch.where(torch.norm(grads, dim=-1) >= grad_threshold, True, False)
        selected = torch.logical_and(selected, torch.max(self.get_scaling, dim=1).values <= self.percent_dense * scene_extent)
        new = {k: p[k][selected] for k in PARAM_ORDER}
        new_skin = self.skin_wts[selected] if self.skin_wts is not None else None
        self.densification_postfix(new, new_skin)

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size) -> None:
        """gaussian.py:309-333 (without the optional outlier removal, which is an offline point-cloud filter)."""
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent)
        prune_mask = (self.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            big_points_vs = self.max_radii2D > max_screen_size
            big_points_ws = self.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
        ls = self.flat.params["log_scale"]
        if torch.isnan(ls.mean()):
            prune_mask = torch.logical_or(prune_mask, torch.any(torch.isnan(ls), dim=-1))
        self.prune_points(prune_mask)

    def reset_opacity(self) -> None:
        """gaussian.py:148-151 with replace_tensor_to_optimizer (:153-161): opacities capped at 0.01, their moments zeroed."""
        op = self.get_opacity
        self.flat.params["opacity_logit"].copy_(inverse_sigmoid(torch.min(op, torch.ones_like(op) * 0.01)))
        if self.opt is not None:
            ms, vs = self._moments()
            ms["opacity_logit"].zero_()
            vs["opacity_logit"].zero_()
