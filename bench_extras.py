"""Secondary measurements of bench.py (rank 0, one GPU, untimed by the driver's headline): BASELINE.json configs 1-3 and 5,
the drop-in path exactly as MANUS calls it, the PyTorch-GPU pre-raster baseline of BASELINE.md section 3, and a warmed CPU baseline.

Everything here goes through the public API of manus_b200 (SceneRenderer / GraphedStep / pose_gaussians / the shims); the
oracle (oracle/pose_ref.py = the reference's P1-P4 restated in PyTorch, pinned to goldens made by the reference's own code)
appears only as the BASELINE being timed, on the CPU (config 1, cpu_baseline) or on the GPU (pytorch_gpu_prerast), never inside
a number reported for this repository's kernels.
"""
from __future__ import annotations

import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def gpu_median_ms(fn, warm: int = 3, reps: int = 20) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.median(out)


def cpu_median_ms(fn, warm: int = 3, reps: int = 20) -> float:
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(out)


def merged_bones(skin21: np.ndarray, keep: int = 16) -> np.ndarray:
    """[N, 20+1] weights -> [N, keep+1]: the weight of every dropped bone goes to bone (b mod keep) -- a synthetic stand-in for
    MANO's native 16-joint weights (mano_rest.pkl is 778x16, remapped 16->20 by mano_to_ours, train_utils.py:68-70): the kernels
    take the bone count at run time, this exercises B = 16+1."""
    nb = skin21.shape[1] - 1
    out = np.zeros((skin21.shape[0], keep + 1), np.float32)
    for b in range(nb):
        out[:, b % keep] += skin21[:, b]
    out[:, keep] = skin21[:, nb]
    return out


def pose_only(dev, n: int = 50_000, view: int = 100):
    """BASELINE config 1: LBS + covariance + SH->RGB + activations of n hand Gaussians for one pose -- PyTorch-CPU restatement of
    the reference (all host threads) against the pose kernels, forward and forward+backward, at B = 16+1 and 20+1."""
    from manus_b200 import synth
    from manus_b200.pose import pose_gaussians
    from oracle import pose_ref

    scene = synth.make_hand(n, seed=0)
    cam = synth.camera(view % 51)
    names = ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]
    posed = torch.tensor(synth.posed_bones(view % 250))
    tf21 = pose_ref.bone_transforms(posed, torch.tensor(scene.bones_rest), True)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    res = {"gaussians": n, "pose": f"novel_pose.pkl pose_matrixs[{view % 250}]", "cpu_threads": torch.get_num_threads(), "os_cpu_count": threads,
           "timing": "3 warm-ups, median of 20; CPU: time.perf_counter, GPU: CUDA events"}
    for label, skin, tf in (("B20+1", scene.skin_wts, tf21), ("B16+1", merged_bones(scene.skin_wts), torch.cat([tf21[:16], tf21[20:]], 0))):
        campos = torch.tensor(cam.camera_center)
        # ---- CPU (the reported baseline)
        leaves = [torch.tensor(getattr(scene, k)).requires_grad_(True) for k in names]
        sk = torch.tensor(skin)

        def cpu_fwd():
            with torch.no_grad():
                return pose_ref.pose_gaussians_ref(*leaves, sk, tf, campos)

        def cpu_fwd_bwd():
            outs = pose_ref.pose_gaussians_ref(*leaves, sk, tf, campos)
            torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])
            for l in leaves:
                l.grad = None

        c_f, c_fb = cpu_median_ms(cpu_fwd, 3, 20), cpu_median_ms(cpu_fwd_bwd, 3, 20)
        # ---- GPU kernels
        gl = [torch.tensor(getattr(scene, k), device=dev).requires_grad_(True) for k in names]
        gsk, gtf, gcam = sk.to(dev), tf.to(dev), campos.to(dev)

        def gpu_fwd():
            with torch.no_grad():
                return pose_gaussians(*gl, gsk, gtf, gcam, 3, False, n)

        ones = None

        def gpu_fwd_bwd():
            nonlocal ones
            outs = pose_gaussians(*gl, gsk, gtf, gcam, 3, False, n)
            if ones is None:
                ones = [torch.ones_like(o) for o in outs]
            torch.autograd.backward(outs, ones)
            for l in gl:
                l.grad = None

        g_f, g_fb = gpu_median_ms(gpu_fwd, 3, 20), gpu_median_ms(gpu_fwd_bwd, 3, 20)
        res[label] = {"cpu_forward_ms": round(c_f, 3), "cpu_forward_backward_ms": round(c_fb, 3), "gpu_forward_ms": round(g_f, 4),
                      "gpu_forward_backward_ms": round(g_fb, 4), "speedup_forward": round(c_f / g_f, 1),
                      "speedup_forward_backward": round(c_fb / g_fb, 1)}
    return res


def probe_loss(image, target):
    """bench.py's probe loss sum(image * G) as one dot product with its analytic gradient (G) handed to the backward; image and
    target are [H,W,3] views of [3,H,W] storage."""
    return torch.dot(image.permute(2, 0, 1).reshape(-1), target.permute(2, 0, 1).reshape(-1)), target


def _prepare(r, views, probe_every: int = 5):
    """Device copies of the per-view inputs and the largest instance count over a sample of the views (exact mode)."""
    from manus_b200 import rasterizer as rz

    dev = r.device
    staged, dmax, ds = {}, 0, []
    rz.set_capacity_mode("exact")
    for k, v in enumerate(views):
        _, c, b = r.view_inputs_host(v)
        staged[v] = (c.to(dev), b.to(dev))
        if k % probe_every == 0:
            with torch.no_grad():
                r.render(v, cam_dev=staged[v][0], bones_dev=staged[v][1])
            d = rz.check_overflow()
            ds.append(d)
            dmax = max(dmax, d)
    return staged, dmax, float(np.mean(ds))


def config_run(dev, kind: str, n: int, W: int, H: int, nviews: int, vif: int, steps: int, peak_gbs: float, seed: int = 0):
    """Frames/s of the graph-replayed step (vif views per step, inputs resident) on another BASELINE configuration."""
    from manus_b200 import rasterizer as rz, synth
    from manus_b200.dist import GraphedStep, SceneRenderer

    scene = {"hand": synth.make_hand, "object": synth.make_object, "composite": synth.make_composite}[kind](n, seed=seed)
    r = SceneRenderer(scene, dev, W, H)
    views = list(range(nviews))
    staged, dmax, dmean = _prepare(r, views, probe_every=max(1, nviews // 10))
    rz.set_capacity_mode("reserve", margin=1.3)
    rz.reserve_capacity(dev.index, scene.n, H, W, dmax)
    G = torch.rand(3, H, W, device=dev).permute(1, 2, 0)
    if vif > 1:      # as the headline: all views' pose backward in one pass
        from manus_b200.dist import PipelinedStep
        step = PipelinedStep(r, probe_loss, G, view=views[0], views_in_flight=vif, chunks=1, reduce=False)
    else:
        step = GraphedStep(r, probe_loss, G, view=views[0], views_in_flight=vif)

    def run(k0, k):
        for it in range(k0, k0 + k):
            for j in range(vif):
                v = views[(it * vif + j) % nviews]
                step.set_inputs(staged[v][0], staged[v][1], None, slot=j)
            step.replay()

    run(0, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(3, steps)
    e1.record()
    torch.cuda.synchronize()
    step.check()
    ms = e0.elapsed_time(e1) / steps
    fps = vif * 1e3 / ms
    nh = scene.n_hand
    fbytes = 1204 * nh + 1036 * (scene.n - nh) + 244 * dmean + 40 * W * H
    out = {"scene": kind, "gaussians": n, "width": W, "height": H, "views": nviews, "views_per_step": vif, "steps": steps,
           "frames_per_s": round(fps, 1), "ms_per_frame": round(ms / vif, 4), "num_rendered_mean": round(dmean),
           "algorithmic_bytes_per_frame": round(fbytes), "achieved_gbps": round(fbytes * fps / 1e9, 1),
           "frac_of_hbm_peak": round(fbytes * fps / 1e9 / peak_gbs, 4)}
    del step, r
    torch.cuda.empty_cache()
    return out


def dropin_lines(dev, scene, W: int, H: int, views, frames: int = 12):
    """The path an UNCHANGED MANUS takes after `import diff_gaussian_rasterization` resolves to shims/ (exact capacity mode: one
    8-byte host read per frame like upstream; eager PyTorch, no CUDA graph):
      pytorch_gpu_prerast_plus_our_raster: P1-P4 as the reference's own PyTorch ops on the GPU (oracle/pose_ref.py) + render_gaussians
                                           through the shim (gaussian_utils.py:363-418)  -- BASELINE.md section 3's 1-GPU baseline;
      patched_pose_plus_shim             : the one-line patch of INTEGRATION.md (pose_gaussians kernel) + the same shim call."""
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    from manus_b200 import rasterizer as rz, synth
    from manus_b200.pose import pose_gaussians
    from manus_b200.render import render_gaussians
    from oracle import pose_ref

    rz.set_capacity_mode("exact")
    names = ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]
    leaves = {k: torch.tensor(getattr(scene, k), device=dev).requires_grad_(True) for k in names}
    nh = scene.n_hand
    skin = None if scene.skin_wts is None else torch.tensor(scene.skin_wts, device=dev)
    rest = None if scene.bones_rest is None else torch.tensor(scene.bones_rest, device=dev)
    G = torch.rand(H, W, 3, device=dev)
    bg = torch.ones(3, device=dev)
    cams = {v: synth.camera(v, W, H) for v in views[:frames]}
    bones = {v: torch.tensor(synth.posed_bones(v), device=dev) for v in views[:frames]}

    def frame_torch(v):
        cam = cams[v]
        cc = torch.tensor(cam.camera_center, device=dev)
        parts = []
        if nh:
            tfs = pose_ref.bone_transforms(bones[v], rest, True)
            parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][:nh] for k in names], skin, tfs, cc))
        if nh < scene.n:
            parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][nh:] for k in names], None, None, cc))
        px, pc, col, op = [torch.cat([p[i] for p in parts], 0) for i in range(4)]
        out = render_gaussians(px, pc, leaves["xyz"], None, op, cam, bg, colors_precomp=col, sh_degree=3, device=dev)
        (out["render"] * G).sum().backward()
        for l in leaves.values():
            l.grad = None

    def frame_patched(v):
        cam = cams[v]
        cc = torch.tensor(cam.camera_center, device=dev)
        tfs = None if not nh else pose_ref.bone_transforms(bones[v], rest, True)     # 21 4x4 products (hand_dynamic.py:93-102)
        px, pc, col, op = pose_gaussians(*[leaves[k] for k in names], skin, tfs, cc, 3, False, nh)
        out = render_gaussians(px, pc, leaves["xyz"], None, op, cam, bg, colors_precomp=col, sh_degree=3, device=dev)
        (out["render"] * G).sum().backward()
        for l in leaves.values():
            l.grad = None

    res = {}
    for name, fn in (("pytorch_gpu_prerast_plus_our_raster", frame_torch), ("patched_pose_plus_shim", frame_patched)):
        vs = list(cams)
        for v in vs[:2]:
            fn(v)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for v in vs:
            fn(v)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / len(vs)
        res[name] = {"frames_per_s": round(1e3 / ms, 1), "ms_per_frame": round(ms, 3), "frames": len(vs),
                     "mode": "eager PyTorch + shims/diff_gaussian_rasterization, exact capacity (one host read per frame), wall clock with a final synchronize"}
    torch.cuda.empty_cache()
    return res


def import_upstream():
    """The reference's own CUDA rasterizer, if somebody provisioned it under baseline/_ref (setup_env.sh:4-13 clones and pip-installs
    graphdeco-inria/diff-gaussian-rasterization; it is not part of /root/reference and there is no network here).  Returns the
    module or None -- never the shim."""
    import importlib

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(ref_dir):
        return None
    saved_path, saved_mod = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] == "diff_gaussian_rasterization"}
    for k in saved_mod:
        del sys.modules[k]
    sys.path = [ref_dir] + [p for p in sys.path if not p.rstrip("/").endswith("shims")]
    try:
        mod = importlib.import_module("diff_gaussian_rasterization")
        if not os.path.abspath(getattr(mod, "__file__", "")).startswith(os.path.abspath(ref_dir)):
            return None
        return mod
    except Exception:
        return None
    finally:
        sys.path = saved_path
        # the returned module object stays usable; the import system goes back to what it held before (e.g. the shim)
        for k in [k for k in sys.modules if k.split(".")[0] == "diff_gaussian_rasterization"]:
            del sys.modules[k]
        sys.modules.update(saved_mod)


def render_like_manus(mod, posed_means, posed_cov, opacity, colors, cam, bg, dev):
    """src/utils/gaussian_utils.py:363-418 with the rasterizer classes of ``mod`` (upstream's module or the shim)."""
    screenspace = torch.zeros_like(posed_means, dtype=posed_means.dtype, requires_grad=True, device=dev) + 0
    screenspace.retain_grad()
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    rs = mod.GaussianRasterizationSettings(image_height=int(cam.height), image_width=int(cam.width), tanfovx=float(cam.tanfovx),
                                           tanfovy=float(cam.tanfovy), bg=bg, scale_modifier=1.0, viewmatrix=t(cam.world_view_transform),
                                           projmatrix=t(cam.full_proj_transform), sh_degree=3, campos=t(cam.camera_center),
                                           prefiltered=False, debug=False)
    image, radii = mod.GaussianRasterizer(raster_settings=rs)(means3D=posed_means, means2D=screenspace, shs=None, colors_precomp=colors,
                                                              opacities=opacity, scales=None, rotations=None, cov3D_precomp=posed_cov)
    return torch.permute(image, (1, 2, 0)), radii, screenspace


def upstream_rasterizer_line(dev, scene, W: int, H: int, views, frames: int = 12):
    """north_star: "next to the reference's own CUDA rasterizer on 1 GPU".  Times PyTorch-GPU P1-P4 + upstream's rasterizer when
    baseline/_ref holds it; returns None otherwise (bench.py then keeps its "unavailable" entry)."""
    mod = import_upstream()
    if mod is None:
        return None
    from manus_b200 import synth
    from oracle import pose_ref

    names = ["xyz", "log_scale", "quat", "opacity_logit", "f_dc", "f_rest"]
    leaves = {k: torch.tensor(getattr(scene, k), device=dev).requires_grad_(True) for k in names}
    nh = scene.n_hand
    skin = None if scene.skin_wts is None else torch.tensor(scene.skin_wts, device=dev)
    rest = None if scene.bones_rest is None else torch.tensor(scene.bones_rest, device=dev)
    G, bg = torch.rand(H, W, 3, device=dev), torch.ones(3, device=dev)

    def frame(v):
        cam = synth.camera(v, W, H)
        cc = torch.tensor(cam.camera_center, device=dev)
        parts = []
        if nh:
            parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][:nh] for k in names], skin,
                                                     pose_ref.bone_transforms(torch.tensor(synth.posed_bones(v), device=dev), rest, True), cc))
        if nh < scene.n:
            parts.append(pose_ref.pose_gaussians_ref(*[leaves[k][nh:] for k in names], None, None, cc))
        px, pc, col, op = [torch.cat([p[i] for p in parts], 0) for i in range(4)]
        img, _, _ = render_like_manus(mod, px, pc, op, col, cam, bg, dev)
        (img * G).sum().backward()
        for l in leaves.values():
            l.grad = None

    vs = list(views[:frames])
    for v in vs[:2]:
        frame(v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for v in vs:
        frame(v)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / len(vs)
    return {"value": 1e3 / ms, "unit": "frames/s", "ms_per_frame": ms, "frames": len(vs), "module": getattr(mod, "__file__", "?"),
            "note": "PyTorch-GPU P1-P4 (the reference's own ops) + upstream diff_gaussian_rasterization from baseline/_ref, eager, wall clock"}
