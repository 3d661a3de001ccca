"""CUDA-graph replay of the whole step (manus_b200.dist.GraphedStep) and device-resident camera intrinsics: the replayed
frame must equal the frame enqueued kernel by kernel, for a view other than the one the graph was captured on."""
import numpy as np
import pytest
import torch

from helpers import grad_close, settings_from, small_scene_inputs, zoom_camera

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _renderer(n=6000, W=256, H=144, seed=3):
    from manus_b200 import synth
    from manus_b200.dist import SceneRenderer, pack_camera

    scene = synth.make_composite(n, seed=seed)
    r = SceneRenderer(scene, torch.device("cuda", 0), W, H)
    for view in (2, 7, 11):       # small images: zoom in so that the hand fills them (different intrinsics per view)
        cam = zoom_camera(view, W, H, 1.2 + 0.05 * view)
        r._cams[view] = (cam, torch.from_numpy(pack_camera(cam)), torch.from_numpy(synth.posed_bones(view).reshape(-1).astype("float32")))
    return scene, r


def test_device_intrinsics_match_host_intrinsics(built_lib):
    from manus_b200.rasterizer import GaussianRasterizer

    cam, t = small_scene_inputs(5, N=4000, W=200, H=120, device=DEV)
    rs = settings_from(cam, (1, 1, 1), DEV)
    tan = torch.tensor([cam.tanfovx, cam.tanfovy], dtype=torch.float32, device=DEV)
    rs_dev = rs._replace(tanfovx=tan[0:1], tanfovy=tan[1:2])
    outs = []
    for s in (rs, rs_dev):
        leaves = {k: v.clone().requires_grad_(True) for k, v in t.items()}
        m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
        img, radii = GaussianRasterizer(s)(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacity"],
                                           colors_precomp=leaves["colors"], cov3D_precomp=leaves["cov3D"])
        (img * torch.linspace(0, 1, img.numel(), device=DEV).reshape(img.shape)).sum().backward()
        outs.append((img.detach(), radii, {k: v.grad for k, v in leaves.items()}))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for k in outs[0][2]:
        ok, e, s = grad_close(outs[1][2][k].cpu().numpy(), outs[0][2][k].cpu().numpy())
        assert ok, (k, e, s)


def test_graph_replay_equals_eager_step(built_lib):
    from manus_b200 import rasterizer as rz
    from manus_b200.dist import GraphedStep

    scene, r = _renderer()
    H, W = r.H, r.W
    dev = r.device
    G = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(7)).to(dev)
    loss_fn = lambda image, target: (image * target).sum()
    rz.set_capacity_mode("exact")
    eager = {}
    try:
        for view in (7, 11):
            _, c, b = r.view_inputs_host(view)
            out = r.render(view, sink=r.flat.grads, cam_dev=c.to(dev), bones_dev=b.to(dev))
            loss = loss_fn(out["render"], G)
            loss.backward()
            eager[view] = (out["render"].detach().clone(), float(loss), r.flat.grad.clone(), int(rz.check_overflow()))
        rz.set_capacity_mode("reserve", margin=1.2)
        rz.reserve_capacity(dev.index, scene.n, H, W, max(e[3] for e in eager.values()))
        step = GraphedStep(r, loss_fn, G, view=2)          # captured on a third view
        for view in (7, 11, 7):
            _, c, b = r.view_inputs_host(view)
            step.set_inputs(c.to(dev), b.to(dev), G)
            loss = step.replay()
            torch.cuda.synchronize()
            assert step.check() == eager[view][3]
            assert abs(float(loss) - eager[view][1]) <= 1e-6 * abs(eager[view][1])
            ok, e, s = grad_close(r.flat.grad.cpu().numpy(), eager[view][2].cpu().numpy())
            assert ok, (view, e, s)
    finally:
        rz.set_capacity_mode("exact")
