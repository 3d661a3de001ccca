"""View-sharded data parallelism for the render path.

The reference trains one view per optimizer step on one GPU and emulates larger batches with gradient accumulation
(``final_loss / accum_iter``, /root/reference/src/modules/hand_dynamic.py:248,259-277).  Here every rank holds the same
Gaussians, renders its own view forward + backward, and the per-Gaussian parameter gradients -- which are plain sums
over views -- are combined by ONE all-reduce of a single flat buffer per step (59 floats per Gaussian at SH degree 3:
xyz 3 | opacity 1 | scaling 3 | rotation 4 | f_dc 3 | f_rest 45, the six Adam param groups of
src/models/gaussian.py:133-140).  There is no exchange inside a frame.

``FlatGaussians`` owns the flat parameter / gradient buffers and hands out the six per-parameter views;
``shard_views`` is the round-robin view assignment; ``allreduce_gradients`` is the collective (NCCL on GPUs; gloo in the
CPU tests).
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

# segment order of the flat buffers: the 11 non-SH floats first (one contiguous all-reduce range), the SH coefficients last
PARAM_ORDER = ("xyz", "opacity_logit", "log_scale", "quat", "f_dc", "f_rest")


def param_widths(sh_coeffs: int = 16, isotropic: bool = False) -> Dict[str, int]:
    return {"xyz": 3, "f_dc": 3, "f_rest": 3 * (sh_coeffs - 1), "opacity_logit": 1, "log_scale": 1 if isotropic else 3, "quat": 4}


def param_shapes(n: int, sh_coeffs: int = 16, isotropic: bool = False):
    return {"xyz": (n, 3), "f_dc": (n, 1, 3), "f_rest": (n, sh_coeffs - 1, 3), "opacity_logit": (n, 1),
            "log_scale": (n, 1 if isotropic else 3), "quat": (n, 4)}


class FlatGaussians:
    """All parameters of a GaussianModel in one contiguous fp32 buffer, gradients in a second one of the same layout
    (segment per parameter, so every view is a dense tensor the kernels can read / write directly)."""

    def __init__(self, n: int, device, sh_coeffs: int = 16, isotropic: bool = False):
        self.n, self.sh_coeffs, self.isotropic = n, sh_coeffs, isotropic
        widths = param_widths(sh_coeffs, isotropic)
        self.floats_per_gaussian = sum(widths.values())
        self.data = torch.zeros(n * self.floats_per_gaussian, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.data)
        self.params: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        off = 0
        shapes = param_shapes(n, sh_coeffs, isotropic)
        for name in PARAM_ORDER:
            cnt = n * widths[name]
            self.params[name] = self.data[off: off + cnt].view(shapes[name])
            self.grads[name] = self.grad[off: off + cnt].view(shapes[name])
            off += cnt

    @classmethod
    def from_scene(cls, scene, device):
        fg = cls(scene.n, device, sh_coeffs=1 + scene.f_rest.shape[1], isotropic=scene.log_scale.shape[1] == 1)
        for name in PARAM_ORDER:
            fg.params[name].copy_(torch.as_tensor(getattr(scene, name)).reshape(fg.params[name].shape))
        return fg

    @classmethod
    def from_params(cls, params: Dict[str, torch.Tensor], device=None):
        """From the six tensors keyed like PARAM_ORDER, e.g. ``manus_b200.densify.initialize_parameters(points, colours)``
        (the reference's GaussianModel.initialize_parameters, src/models/gaussian.py:99-127)."""
        device = params["xyz"].device if device is None else device
        fg = cls(params["xyz"].shape[0], device, sh_coeffs=1 + params["f_rest"].shape[1], isotropic=params["log_scale"].shape[1] == 1)
        for name in PARAM_ORDER:
            fg.params[name].copy_(torch.as_tensor(params[name]).reshape(fg.params[name].shape))
        return fg

    def leaves(self) -> List[torch.Tensor]:
        """The six parameters as autograd leaves, in the order pose_gaussians takes them
        (xyz, log_scale, quat, opacity_logit, f_dc, f_rest)."""
        order = ("xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest")
        return [self.params[k].detach().requires_grad_(True) for k in order]

    def scratch_like_grad(self) -> "FlatGaussians":
        """A second gradient buffer with the same layout (for accumulating more than one view per rank and step)."""
        other = FlatGaussians.__new__(FlatGaussians)
        other.n, other.sh_coeffs, other.isotropic = self.n, self.sh_coeffs, self.isotropic
        other.floats_per_gaussian = self.floats_per_gaussian
        other.data, other.params = self.data, self.params
        other.grad = torch.zeros_like(self.grad)
        other.grads = {}
        off = 0
        for name in PARAM_ORDER:
            cnt = self.grads[name].numel()
            other.grads[name] = other.grad[off: off + cnt].view(self.grads[name].shape)
            off += cnt
        return other

    def allreduce_bytes(self) -> int:
        return self.grad.numel() * 4


def shard_views(views: Sequence[int], rank: int, world_size: int) -> List[int]:
    """Round-robin: rank r renders views {v : index(v) mod R == r}."""
    return [v for i, v in enumerate(views) if i % world_size == rank]


def allreduce_gradients(flat_grad: torch.Tensor, world_size: Optional[int] = None, average: bool = True) -> None:
    """One in-place SUM all-reduce of the flat per-Gaussian gradient buffer; ``average`` divides by the number of views
    in the step, which is exactly the reference's ``final_loss / accum_iter`` scaling."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        world_size = dist.get_world_size()
    if average and world_size and world_size > 1:
        flat_grad.div_(world_size)


def sharded_step(flat: FlatGaussians, my_views: Sequence[int],
                 render_backward: Callable[[FlatGaussians, int, Dict[str, torch.Tensor]], torch.Tensor],
                 scratch: Optional[FlatGaussians] = None, global_views: Optional[int] = None) -> torch.Tensor:
    """Gradients of one optimisation step.  ``render_backward(flat, view, sink)`` runs forward + backward of one view and
    OVERWRITES the tensors in ``sink`` (layout of ``flat.grads``) with that view's parameter gradients, returning the scalar
    loss.  The first view writes straight into the flat all-reduce buffer, further views of the same rank go through
    ``scratch`` and are added.  Then ONE all-reduce (SUM) of the flat buffer and a division by the global number of views
    (the reference's ``final_loss / accum_iter``).  ``global_views`` = number of views of this step over all ranks (every
    rank can compute it from the view list); when omitted it is obtained with a second tiny all-reduce and a host read.
    Returns the mean loss over all views of the step (device scalar)."""
    dev = flat.grad.device
    total = torch.zeros((), device=dev)
    if len(my_views) == 0:
        flat.grad.zero_()
    for j, v in enumerate(my_views):
        if j == 0:
            loss = render_backward(flat, v, flat.grads)
        else:
            if scratch is None:
                scratch = flat.scratch_like_grad()
            loss = render_backward(flat, v, scratch.grads)
            flat.grad.add_(scratch.grad)
        total = total + loss.detach().reshape(())
    n_views = float(len(my_views))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
        stats = torch.stack([total, torch.tensor(n_views, device=dev)])
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        total = stats[0]
        n_views = float(global_views) if global_views is not None else float(stats[1])
    if n_views > 1:
        flat.grad.div_(n_views)
    return total / max(n_views, 1.0)


CAM_FLOATS = 39   # view 16 | proj 16 | centre 3 | fovx, fovy | tan(fovx/2), tan(fovy/2)


def pack_camera(cam):
    """One camera as the flat fp32 record the renderer takes per view (host side, numpy)."""
    import numpy as np

    return np.concatenate([np.asarray(cam.world_view_transform).reshape(-1), np.asarray(cam.full_proj_transform).reshape(-1),
                           np.asarray(cam.camera_center).reshape(-1), [cam.fovx, cam.fovy, cam.tanfovx, cam.tanfovy]]).astype("float32")


class SceneRenderer:
    """Everything one rank needs to render views of a (hand | object | composite) scene forward + backward through the
    public API: resident parameters (flat), skin weights, rest bones, and per-view cameras / bone poses."""

    def __init__(self, scene, device, width: int = 1920, height: int = 1080, bg=(1.0, 1.0, 1.0), sh_degree: int = 3, plan=None):
        from . import rasterizer as rz, synth

        # how this renderer's frames size their instance buffers (exact / reserve, high-water marks, the last frame's state):
        # the device's default plan unless the caller hands over a private rasterizer.CapacityPlan
        self.plan = plan if plan is not None else rz.plan_for(device)

        self.device, self.W, self.H, self.sh_degree = device, width, height, sh_degree
        self.flat = FlatGaussians.from_scene(scene, device)
        self.n_hand = scene.n_hand
        self.skin = None if scene.skin_wts is None else torch.as_tensor(scene.skin_wts).to(device)
        self.bones_rest = None if scene.bones_rest is None else torch.as_tensor(scene.bones_rest).to(device)
        self.rest_inv = None if self.bones_rest is None else torch.linalg.inv(self.bones_rest)   # constant per scene
        self._eye = torch.eye(4, dtype=torch.float32, device=device)[None]
        self.bg = torch.tensor(bg, dtype=torch.float32, device=device)
        self._synth = synth
        self._cams = {}
        # render_fused(fuse_backward=...): one autograd node for pose + rasterizer, projection backward inside the pose backward
        # (measured: 1461 -> 1484 frames/s with one view per step, 1891 -> 1921 with four in flight); "0" selects the
        # two-node path (pose_gaussians + GaussianRasterizer), which the GPU tests run as well
        self.fuse_backward = os.environ.get("MANUS_B200_FUSE_BACKWARD", "1") == "1"

    def rebind(self, flat: FlatGaussians, skin: Optional[torch.Tensor] = None) -> None:
        """After densify / prune (``GaussianState`` rebuilds its flat buffers and N changes): point the renderer at the new
        buffers.  Step objects captured on the old buffers (GraphedStep / PipelinedStep) refuse to replay afterwards."""
        self.flat = flat
        self.n_hand = 0 if skin is None else skin.shape[0]
        self.skin = skin
        self.generation = getattr(self, "generation", 0) + 1

    def view_inputs_host(self, view: int):
        """Per-step inputs as pinned host tensors: camera (view 16 | proj 16 | centre 3 | fovx, fovy) and posed bones [20,16]."""
        if view not in self._cams:
            cam = self._synth.camera(view, self.W, self.H)
            import numpy as np
            packed = pack_camera(cam)
            bones = self._synth.posed_bones(view).reshape(-1).astype("float32")
            self._cams[view] = (cam, torch.from_numpy(packed).pin_memory(), torch.from_numpy(bones).pin_memory())
        return self._cams[view]

    def render(self, view: int, sink: Optional[Dict[str, torch.Tensor]] = None, cam_dev=None, bones_dev=None,
               device_intrinsics: bool = False, compact_sh: bool = False, accumulate: bool = False, slot: int = 0,
               fuse_backward: Optional[bool] = None, want_posed: bool = False):
        """Forward of one view through render_fused; returns the result dict (image is out['render'], HWC).
        cam_dev / bones_dev: the packed per-view inputs already on the device (``view_inputs_host`` layout).
        device_intrinsics: read tan(fov/2) from cam_dev[37:39] on the device instead of from the host camera, so that the
        enqueued frame does not depend on which view ``cam_dev`` holds (CUDA-graph replay).
        compact_sh: the backward does not write the f_rest gradient (45 of the 59 floats per Gaussian); it is rebuilt from the
        f_dc gradients of all ranks' views by ``CompactGradExchange`` (the per-view SH gradient is rank one).
        accumulate: the backward ADDS this view's parameter gradients to ``sink`` (gradient accumulation over the views of a step).
        slot: views that are in flight at the same time (``GraphedStep(views_in_flight=V)``) use different slots, each with its
        own bone-transform buffer.
        want_posed: also return posed_xyz / posed_cov / colors / cano_opacity (52 B per Gaussian written to HBM); on the fused
        path they are otherwise left in registers between the pose step and the projection."""
        from .cameras import Camera
        from .pose import bone_transforms
        from .render import render_fused

        cam, cam_host, bones_host = self.view_inputs_host(view)
        if cam_dev is None:
            cam_dev = cam_host.to(self.device, non_blocking=True)
            bones_dev = bones_host.to(self.device, non_blocking=True)
        dcam = Camera(cam.width, cam.height, cam.fovx, cam.fovy, cam_dev[0:16].view(4, 4), cam_dev[16:32].view(4, 4), cam_dev[32:35], None,
                      tanfov_dev=cam_dev[37:39] if device_intrinsics else None)
        bone_tf = None
        fuse = self.fuse_backward if fuse_backward is None else fuse_backward
        if not fuse:
            from . import rasterizer as rz
            if self.plan is not rz.plan_for(self.device):
                # the two-node path goes through the drop-in GaussianRasterizer, whose signature (upstream's) has no room for a plan
                raise ValueError("a private CapacityPlan needs the fused path (fuse_backward=True)")
        if self.n_hand > 0:
            nb = self.rest_inv.shape[0]
            if fuse and not compact_sh:
                # the kernels build T_b = posed_b rest_b^-1 (+ the identity "background" row) in their prologue: no per-frame
                # batched product, nothing to keep per slot
                bone_tf = (bones_dev.view(-1, 4, 4), self.rest_inv, nb + 1)
            else:
                # = bone_transforms(posed, rest, append_identity=True), written into a persistent [B+1,4,4] buffer whose last
                # row stays the identity (one small batched product per frame; the compact exchange ships this buffer)
                tfs = self.__dict__.setdefault("_bone_tfs", {})
                if slot not in tfs:
                    tfs[slot] = torch.eye(4, dtype=torch.float32, device=self.device).repeat(nb + 1, 1, 1)
                bone_tf = tfs[slot]
                torch.bmm(bones_dev.view(-1, 4, 4), self.rest_inv, out=bone_tf[:nb])
                self._bone_tf = bone_tf
        if compact_sh and sink is not None:
            sink = dict(sink, f_rest=None)
        self._last_campos = cam_dev[32:35]
        return render_fused(self.flat.leaves(), self.skin, bone_tf, dcam, self.bg, self.sh_degree, self.flat.isotropic,
                            self.n_hand, grad_sink=sink, accumulate=accumulate, fuse_backward=fuse, want_posed=want_posed, plan=self.plan)


class CompactGradExchange:
    """The one exchange of the data-parallel step with 3.5x fewer bytes on the wire than all-reducing the flat buffer.

    Of the 59 gradient floats per Gaussian, 48 belong to the SH coefficients, and for ONE view they are rank one:
    g_f_dc[c] = basis_0 * go[c], g_f_rest[k][c] = basis_k(dir) * go[c] (go = colour gradient, dir = canonical-space view
    direction, a function of the view's bone transforms and camera centre).  So every rank renders with ``compact_sh=True``
    (its backward skips the 180-byte f_rest gradient), then: ONE all-gather of a per-rank record (g_f_dc, 12 B per Gaussian,
    plus the view's bone transforms and camera centre); ONE all-reduce of the other 11 floats (the front of the flat
    buffer); rebuild sum_r basis(dir_r) x go_r locally (``mb_sh_grad_from_views``).  Same result as ``all_reduce(flat.grad)`` up to fp32 summation order.
    ``rebuild``: the reconstruction function (default: the CUDA kernel; the gloo test injects a CPU restatement)."""

    def __init__(self, renderer: "SceneRenderer", group=None, rebuild=None):
        self.r, self.group = renderer, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        flat = renderer.flat
        n, dev = flat.n, flat.grad.device
        self.nb = 0 if renderer.n_hand == 0 else renderer.rest_inv.shape[0] + 1
        # one record per rank: DC gradients [N,3] | bone transforms [B,4,4] | camera centre [3] (padded to 4 floats)
        self.rec = n * 3 + self.nb * 16 + 4
        self.send = torch.zeros(self.rec, dtype=torch.float32, device=dev)
        self.recv = torch.zeros(self.world, self.rec, dtype=torch.float32, device=dev)
        # the non-SH gradients are the front of the flat buffer (PARAM_ORDER): one contiguous all-reduce range
        w = param_widths(flat.sh_coeffs, flat.isotropic)
        self.head = flat.grad[: n * (w["xyz"] + w["opacity_logit"] + w["log_scale"] + w["quat"])]
        self.rebuild = rebuild
        self._flat_id, self._n = id(flat), n      # the staging buffers and `head` are views / sized for THIS flat buffer
        # a second communicator (own stream) for the all-reduce, so that it runs beside the all-gather and the rebuild
        self.group2 = dist.new_group() if self.world > 1 and group is None else group

    def __call__(self) -> None:
        r, flat = self.r, self.r.flat
        n = flat.n
        if id(flat) != self._flat_id or n != self._n:
            # densify / prune rebuilt the flat buffers (SceneRenderer.rebind): `head` would all-reduce the stale buffer
            raise RuntimeError("this exchange was built on buffers that have been replaced (densify / prune): build a new CompactGradExchange")
        pending = None
        if self.world > 1:
            pending = dist.all_reduce(self.head, group=self.group2, async_op=True)     # see below
        self.send[: n * 3].copy_(flat.grads["f_dc"].reshape(-1))
        if self.nb:
            self.send[n * 3: n * 3 + self.nb * 16].copy_(r._bone_tf.reshape(-1))
        self.send[n * 3 + self.nb * 16: n * 3 + self.nb * 16 + 3].copy_(r._last_campos.reshape(-1)[:3])
        if self.world > 1:
            # the all-reduce of the non-SH gradients (started above on its own communicator) runs beside the staging
            # copies, the all-gather and the rebuild of the SH gradients
            dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        else:
            self.recv[0].copy_(self.send)
        gfdc_all = self.recv[:, : n * 3].unflatten(1, (n, 3))
        bone_all = self.recv[:, n * 3: n * 3 + self.nb * 16].unflatten(1, (self.nb, 4, 4)) if self.nb else None
        campos_all = self.recv[:, n * 3 + self.nb * 16: n * 3 + self.nb * 16 + 3]
        rebuild = self.rebuild
        if rebuild is None:
            from .pose import sh_grad_from_views as rebuild
        rebuild(flat.params["xyz"], r.skin, r.n_hand, r.sh_degree, flat.sh_coeffs, bone_all, campos_all, gfdc_all,
                flat.grads["f_dc"], flat.grads["f_rest"])
        if pending is not None:
            pending.wait()


def _loss_and_seed(loss_fn, image, target):
    """loss_fn returns the scalar loss, or (loss, dL/dimage) when it knows its own gradient analytically (e.g. a linear probe
    loss sum(image * G): the gradient is G, no autograd kernel has to form it).  -> (loss, image tensor to call backward on, seed)."""
    res = loss_fn(image, target)
    if isinstance(res, tuple):
        loss, seed = res
        return loss, image, seed
    return res, res, None


def _backward(root, seed) -> None:
    if seed is None:
        root.backward()
    else:
        root.backward(seed)


def _remember_buffers(step, renderer) -> None:
    step._flat_id, step._n, step._replays = id(renderer.flat), renderer.flat.n, 0


def _check_fresh(step, check_every: int = 512) -> None:
    """A captured step holds the ADDRESSES of the parameter / gradient buffers it was captured on: after a densify / prune rebuilt
    them (``SceneRenderer.rebind``) a replay would silently train the stale buffers -- refuse instead.  Every ``check_every``
    replays the step also reads the overflow counters of its frames (one synchronisation): in reserve capacity mode an overflow
    truncates the farthest instances, and nothing else would notice."""
    r = step.r
    if id(r.flat) != step._flat_id or r.flat.n != step._n:
        raise RuntimeError("this step was captured on buffers that have been replaced (densify / prune): build a new step object")
    step._replays += 1
    if check_every and step._replays % check_every == 0:
        step.check()


def gaussian_chunks(n: int, chunks: int, align: int = 128) -> List[tuple]:
    """[lo, hi) ranges that split n Gaussians into ``chunks`` pieces whose starts are multiples of ``align`` (the pose kernels' tile)."""
    per = -(-n // max(chunks, 1))
    per = -(-per // align) * align
    return [(lo, min(lo + per, n)) for lo in range(0, n, per)]


class PipelinedStep:
    """The data-parallel step with the gradient exchange hidden behind the pose backward.

    Per step and rank: V views (gradient accumulation, as ``GraphedStep``) run pose + projection forward, binning, tile forward,
    loss, and the TILE backward as parallel branches of one CUDA graph -- everything except the last kernel of each view, the
    pose backward, which is the only writer of the parameter gradients.  That kernel is then launched range by range over the
    Gaussians (C ranges; per range: view 0 overwrites, views 1..V-1 add), and as soon as a range is complete on this rank its
    six pieces of the flat gradient buffer (one per parameter) are all-reduced as ONE coalesced NCCL operation on the
    communicator's stream while the next range computes: the collective (0.38 ms for 118 MB at 8 ranks) overlaps the
    HBM-bound pose backward (0.09 ms per view) instead of following it.  Same sums as one all-reduce of the whole buffer
    (per element: the same operands; the order over ranks is NCCL's in both cases).

    group=None and no initialised process group: single rank, no collective (used by the tests to check the chunked backward).
    fused_views: the deferred pose backwards of a range run as ONE kernel pass over all their views
    (mb_pose_backward_from_raster_views) instead of one launch per view.
    stats = (xyz_gradient_accum [N,1], denom [N,1], max_radii2D [N]): the densification statistics of every view are updated by
    the pose backward kernel itself (``GaussianState.add_densification_stats`` semantics; they are per-rank sums / maxima until
    ``GaussianState.reduce_stats`` combines them, once per densification interval)."""

    def __init__(self, renderer: SceneRenderer, loss_fn, target_like: torch.Tensor, view: int = 0, views_in_flight: int = 1,
                 chunks: int = 4, warmup: int = 3, group=None, stats=None, exchange=None, exchange_ctas: int = 0,
                 deferred_views: Optional[int] = None, fused_views: bool = True, reduce: bool = True):
        from . import rasterizer as rz

        # deferred_views = d: only the LAST d views of the step leave their pose backward for the range-by-range tail; the others
        # run it inside their branch (hidden under the tile kernels of the other views, as in GraphedStep) and add into the buffer,
        # which is then cleared at the head of the step.  The tail is as long as the exchange either way; fewer deferred views
        # mean less serialised pose backward in front of it.  None = all views.
        self.deferred_views = int(views_in_flight) if deferred_views is None else max(1, min(int(deferred_views), int(views_in_flight)))
        self.stats = stats
        # exchange: a manus_b200.exchange.MulticastExchange built on renderer.flat BEFORE this step (the gradient buffer then lives
        # in NVSwitch multicast memory and the ranges are summed by the repository's own multimem kernel on a side stream);
        # None: one coalesced NCCL all-reduce per range
        self.exchange, self.exchange_ctas = exchange, exchange_ctas
        if renderer.plan.mode != "reserve":
            raise RuntimeError("PipelinedStep needs set_capacity_mode('reserve') and reserve_capacity(...) (no host read-back in a graph)")
        self.r, self.V, self.group = renderer, int(views_in_flight), group
        # reduce=False: replay() leaves the exchange to the caller (e.g. one all-reduce of the whole buffer after the step)
        self.world = dist.get_world_size(group) if (reduce and dist.is_available() and dist.is_initialized()) else 1
        dev = renderer.device
        V = self.V
        _, cam_host, bones_host = renderer.view_inputs_host(view)
        self.cams = [cam_host.to(dev) for _ in range(V)]
        self.bones_all = [bones_host.to(dev) for _ in range(V)]
        self.targets = [target_like.to(dev).clone() for _ in range(V)]
        self.ranges = gaussian_chunks(renderer.flat.n, chunks)
        flat = renderer.flat
        self.pieces = [[flat.grads[name][lo:hi] for name in PARAM_ORDER] for lo, hi in self.ranges]
        if exchange is not None:
            self.piece_ranges = [exchange.pieces(lo, hi) for lo, hi in self.ranges]
            self.comm = torch.cuda.Stream(device=dev)
            self._done = [torch.cuda.Event() for _ in self.ranges]
        self._sides = [torch.cuda.Stream(device=dev) for _ in range(V - 1)]
        self.states = [None] * V
        self.viewspace = [None] * V
        deferred: list = []

        def front():
            """Every view up to and including its tile backward; the pose backwards are collected in ``deferred``."""
            deferred.clear()
            cur = torch.cuda.current_stream(dev)
            losses = [None] * V
            early = V - self.deferred_views          # views [0, early) finish inside their branch and ADD to the buffer
            if early > 0:
                flat.grad.zero_()
            for side in self._sides:
                side.wait_stream(cur)
            outs = [None] * V
            for i in range(V):
                with torch.cuda.stream(cur if i == 0 else self._sides[i - 1]):
                    sink = dict(flat.grads, _stats=stats)
                    if i >= early:
                        sink["_defer"] = []
                    out = renderer.render(view, sink=sink, cam_dev=self.cams[i], bones_dev=self.bones_all[i], device_intrinsics=True, slot=i,
                                          accumulate=i < early)
                    self.states[i] = renderer.plan.last_state
                    self.viewspace[i] = out["viewspace_points"]
                    losses[i] = _loss_and_seed(loss_fn, out["render"], self.targets[i])
                    outs[i] = sink
            for i in range(V):
                with torch.cuda.stream(cur if i == 0 else self._sides[i - 1]):
                    _backward(losses[i][1], losses[i][2])
                    losses[i] = losses[i][0].detach()
                    deferred.extend(outs[i].get("_defer", []))
            for i in range(1, V):
                cur.wait_stream(self._sides[i - 1])
            return (losses[0] if V == 1 else torch.stack(losses).sum()), list(deferred)

        def back(defs, c):
            lo, hi = self.ranges[c]
            if fused_views and len(defs) <= 8:
                # ONE pass over the range for all deferred views (parameters staged once, gradients written once)
                from .render import pose_backward_views
                pose_backward_views(defs, lo, hi, accumulate=self.deferred_views < V)
            else:
                for i, d in enumerate(defs):
                    d.run(lo, hi, accumulate=i > 0 or self.deferred_views < V)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                _, defs = front()
                for c in range(len(self.ranges)):
                    back(defs, c)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        _remember_buffers(self, renderer)
        self.graph_front = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_front):
            self.loss, self._defs = front()
        # the per-range graphs share the front graph's memory pool: they read its saved state (records, radii, accumulator rows)
        self.graph_back = []
        for c in range(len(self.ranges)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.graph_front.pool()):
                back(self._defs, c)
            self.graph_back.append(g)
        if stats is not None:       # the warm-up steps above went through the same kernels: start the statistics from zero
            for s_ in stats:
                s_.zero_()

    def set_inputs(self, cam_dev, bones_dev, target_dev=None, slot: int = 0) -> None:
        self.cams[slot].copy_(cam_dev, non_blocking=True)
        self.bones_all[slot].copy_(bones_dev, non_blocking=True)
        if target_dev is not None:
            self.targets[slot].copy_(target_dev, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """Enqueue one step; the gradients in ``renderer.flat.grad`` are the sum over all ranks' views when the current stream
        reaches the end of what this call enqueued."""
        _check_fresh(self)
        self.graph_front.replay()
        works = []
        if self.exchange is not None and self.world > 1:
            cur = torch.cuda.current_stream(self.r.device)
            for c, g in enumerate(self.graph_back):
                g.replay()
                self._done[c].record(cur)
                with torch.cuda.stream(self.comm):      # the range's sum over the ranks runs beside the next range's pose backward
                    self.comm.wait_event(self._done[c])
                    self.exchange.all_reduce(self.piece_ranges[c], self.exchange_ctas)
            cur.wait_stream(self.comm)
            return self.loss
        for c, g in enumerate(self.graph_back):
            g.replay()
            if self.world > 1:
                # one coalesced all-reduce of the range's six pieces; issued behind the range's pose backwards on the current
                # stream, it runs on the communicator's stream beside the next range's kernels
                with dist._coalescing_manager(group=self.group, device=self.r.device, async_ops=True) as cm:
                    for p in self.pieces[c]:
                        dist.all_reduce(p, group=self.group)
                works.append(cm)
        for w in works:
            w.wait()
        return self.loss

    def check(self) -> int:
        from . import rasterizer as rz

        return sum(rz.check_overflow(st) for st in self.states)


class GraphedStep:
    """One training step -- for each of its views: pose forward, rasterizer forward, loss, rasterizer backward, pose backward
    into the flat gradient buffer -- captured ONCE in a CUDA graph and replayed per step with a single launch.  The per-view
    inputs (packed camera, posed bones, loss target) live in static device tensors that the caller overwrites before
    ``replay()``; nothing in the captured step depends on host values of the view (intrinsics are read from ``cam`` on the
    device) and nothing is read back (reserve capacity mode), so the same graph serves every view of the scene.

    views_in_flight = V > 1: the step holds V independent views (the reference's gradient accumulation over ``accum_iter``
    views, hand_dynamic.py:248,259-277; the views of one rank in the view-sharded step).  Each view is captured on its own
    stream, so the graph has V parallel branches: one frame cannot fill a B200 (the blend kernels wait on their deepest tile,
    the sort passes on look-back latency), and the HBM-bound pose kernels of one view run under the latency-bound tile
    kernels of another (measured at 500k / 1080p: 0.64 ms per frame alone, 0.52 with two, 0.48 with four views in flight).
    The gradients are summed into ``renderer.flat.grad`` by the pose backward kernels themselves (TMA reduce-add, fp32 adds
    resolved in L2): the buffer is cleared at the head of the graph and every view adds to it, in whatever order the branches
    finish (like the atomics of the blend backward, the summation order is not fixed).  ``ordered=True`` instead chains the
    pose backward launches in view order (view 0 overwrites, view i adds after view i-1: event edges in the graph) for a
    reproducible sum, at the price of serialising the last kernel of every branch.

    loss_fn(image[H,W,3], target) -> scalar tensor, or (scalar, dL/dimage [H,W,3]) when the loss knows its gradient.  After ``replay()``: ``loss`` (device scalar: the sum over the step's
    views; ``losses`` holds them one by one) and ``renderer.flat.grad`` (sum over the views) hold the step's results.
    ``check()`` (synchronises) raises if a replayed frame overflowed the reserved capacity.  ``radii_all[i]`` are view i's
    radii; the views' screen-space gradients (``viewspace_points.grad``, the densification statistic of
    src/models/gaussian.py:335-338, which the reference accumulates every step while ``global_step < densify_until_step``,
    src/utils/gaussian_utils.py:466-473) are in ``viewspace[i].grad`` after a replay.
    """

    def __init__(self, renderer: SceneRenderer, loss_fn, target_like: torch.Tensor, view: int = 0, warmup: int = 3,
                 compact_sh: bool = False, views_in_flight: int = 1, ordered: bool = False, profile: bool = False, stats=None):
        from . import rasterizer as rz

        self.stats = stats      # (xyz_gradient_accum, denom, max_radii2D) updated by the pose backward of every view, or None

        if renderer.plan.mode != "reserve":
            raise RuntimeError("GraphedStep needs set_capacity_mode('reserve') and reserve_capacity(...) (no host read-back in a graph)")
        V = int(views_in_flight)
        if V < 1 or (compact_sh and V > 1):
            raise ValueError("views_in_flight must be >= 1 (and 1 with compact_sh: the compact exchange carries one view per rank)")
        self.r, self.V = renderer, V
        dev = renderer.device
        _, cam_host, bones_host = renderer.view_inputs_host(view)
        self.cams = [cam_host.to(dev) for _ in range(V)]
        self.bones_all = [bones_host.to(dev) for _ in range(V)]
        self.targets = [target_like.to(dev).clone() for _ in range(V)]
        self.cam, self.bones, self.target = self.cams[0], self.bones_all[0], self.targets[0]
        self.view = view
        self.loss_fn = loss_fn
        self.states = [None] * V
        self.viewspace = [None] * V      # per view: the screen-space leaf; .grad is rewritten by every replay

        def frame(i, done):
            sink = renderer.flat.grads if stats is None else dict(renderer.flat.grads, _stats=stats)
            if V > 1 and ordered:
                # the pose backward kernels (the last kernel of each branch, the only writers of the flat gradient buffer) are
                # chained in view order through events; everything before them runs concurrently
                sink = dict(sink, _wait=done[i - 1] if i > 0 else None, _record=done[i])
            out = renderer.render(view, sink=sink, cam_dev=self.cams[i], bones_dev=self.bones_all[i], device_intrinsics=True,
                                  compact_sh=compact_sh, accumulate=V > 1 and (i > 0 or not ordered), slot=i)
            self.states[i] = renderer.plan.last_state
            self.viewspace[i] = out["viewspace_points"]
            loss = _loss_and_seed(loss_fn, out["render"], self.targets[i])
            return loss, out["radii"]

        def step():
            # every view on its own stream (view 0 on the current one)
            cur = torch.cuda.current_stream(dev)
            losses, radii = [None] * V, [None] * V
            done = [torch.cuda.Event() for _ in range(V)]
            if V > 1 and not ordered:
                renderer.flat.grad.zero_()          # every view adds to the buffer
            for side in self._sides:                # fork first: no branch waits for work of another branch
                side.wait_stream(cur)
            for i in range(V):
                with torch.cuda.stream(cur if i == 0 else self._sides[i - 1]):
                    losses[i], radii[i] = frame(i, done)
            for i in range(V):              # host order = view order: event i-1 is recorded before view i waits for it
                with torch.cuda.stream(cur if i == 0 else self._sides[i - 1]):
                    _backward(losses[i][1], losses[i][2])
                    losses[i] = losses[i][0].detach()
            for i in range(1, V):
                cur.wait_stream(self._sides[i - 1])
            total = losses[0] if V == 1 else torch.stack(losses).sum()
            return total, losses, radii

        self._sides = [torch.cuda.Stream(device=dev) for _ in range(V - 1)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        _remember_buffers(self, renderer)
        self.graph = torch.cuda.CUDAGraph()
        if profile:     # event-record nodes around every library kernel of the graph: _lib.profile_timeline() after a replay
            from . import _lib

            _lib.profile_report()
            _lib.profile_enable(True)
        with torch.cuda.graph(self.graph):
            self.loss, self.losses, radii = step()
        if profile:
            _lib.profile_enable(False)
        self.radii = radii[0]
        self.radii_all = radii
        self.state = self.states[0]
        if stats is not None:       # the warm-up steps above went through the same kernels: start the statistics from zero
            for s_ in stats:
                s_.zero_()

    def set_inputs(self, cam_dev: torch.Tensor, bones_dev: torch.Tensor, target_dev: Optional[torch.Tensor] = None,
                   slot: int = 0) -> None:
        """Device-to-device copies into the static inputs of view ``slot`` of the step (enqueued on the current stream)."""
        self.cams[slot].copy_(cam_dev, non_blocking=True)
        self.bones_all[slot].copy_(bones_dev, non_blocking=True)
        if target_dev is not None:
            self.targets[slot].copy_(target_dev, non_blocking=True)

    def replay(self) -> torch.Tensor:
        _check_fresh(self)
        self.graph.replay()
        return self.loss

    def check(self) -> int:
        """num_rendered of the last replay (summed over the step's views); raises on a capacity overflow."""
        from . import rasterizer as rz

        return sum(rz.check_overflow(st) for st in self.states)
