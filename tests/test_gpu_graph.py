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


def test_renderer_with_a_private_capacity_plan(built_lib):
    """SceneRenderer(plan=CapacityPlan(...)): the renderer's frames are sized by ITS plan (here: reserve mode while the device's
    default plan stays exact), the device plan sees none of them, and a reserve that is too small is reported by check_overflow."""
    from manus_b200 import _lib, rasterizer as rz, synth
    from manus_b200.dist import SceneRenderer

    scene, r0 = _renderer()
    dev = r0.device
    rz.set_capacity_mode("exact")
    _, c, b = r0.view_inputs_host(7)
    c, b = c.to(dev), b.to(dev)
    img0 = r0.render(7, cam_dev=c, bones_dev=b)["render"].detach().clone()
    need = rz.check_overflow()
    shared_last = rz.plan_for(dev).last_state
    assert need > 0 and r0.plan is rz.plan_for(dev)
    mine = rz.CapacityPlan("reserve", margin=1.2)
    mine.reserve(scene.n, r0.H, r0.W, need)
    r1 = SceneRenderer(scene, dev, r0.W, r0.H, plan=mine)
    r1._cams = r0._cams
    img1 = r1.render(7, cam_dev=c, bones_dev=b)["render"].detach()
    assert mine.last_state is not None and mine.last_state.host_count is None          # sized without a host read
    assert rz.plan_for(dev).last_state is shared_last and rz.plan_for(dev).mode == "exact"
    assert rz.check_overflow(mine.last_state) == need and torch.equal(img0, img1)
    with pytest.raises(ValueError):
        r1.render(7, cam_dev=c, bones_dev=b, fuse_backward=False)
    tight = rz.CapacityPlan("reserve", margin=1.0)
    tight.reserve(scene.n, r0.H, r0.W, max(need // 2 - 1024, 1))
    r2 = SceneRenderer(scene, dev, r0.W, r0.H, plan=tight)
    r2._cams = r0._cams
    r2.render(7, cam_dev=c, bones_dev=b)
    with pytest.raises(_lib.ManusB200Error):
        rz.check_overflow(tight.last_state)


def test_accumulating_pose_backward_adds_to_the_sink(built_lib):
    """mb_pose_backward_accumulate (TMA reduce-add): sink_after = sink_before + gradient, also across the ragged last tile and
    the skinned / static boundary (runs that are not 16-byte sized take the atomicAdd path)."""
    scene, r = _renderer(n=6007)
    dev = r.device
    G = torch.rand(r.H, r.W, 3, generator=torch.Generator().manual_seed(3)).to(dev)
    _, c, b = r.view_inputs_host(7)
    c, b = c.to(dev), b.to(dev)
    (r.render(7, sink=r.flat.grads, cam_dev=c, bones_dev=b)["render"] * G).sum().backward()
    g = r.flat.grad.clone()
    base = torch.randn_like(g) * g.abs().mean()
    r.flat.grad.copy_(base)
    (r.render(7, sink=r.flat.grads, cam_dev=c, bones_dev=b, accumulate=True)["render"] * G).sum().backward()
    ok, e, s = grad_close((r.flat.grad - base).cpu().numpy(), g.cpu().numpy(), 2e-6)
    assert ok, (e, s)
    assert float(g.abs().max()) > 0


@pytest.mark.parametrize("V,ordered", [(2, False), (3, False), (3, True)])
def test_views_in_flight_sum_the_single_view_gradients(built_lib, V, ordered):
    """GraphedStep(views_in_flight=V): V views captured on V streams of one graph; flat.grad = sum of the views' gradients
    (reference: gradient accumulation over accum_iter views, hand_dynamic.py:248,259-277), loss = sum of the views' losses."""
    from manus_b200 import rasterizer as rz
    from manus_b200.dist import GraphedStep

    scene, r = _renderer()
    H, W, dev = r.H, r.W, r.device
    targets = [torch.rand(H, W, 3, generator=torch.Generator().manual_seed(20 + i)).to(dev) for i in range(3)]
    loss_fn = lambda image, target: (image * target).sum()
    views = (7, 11, 2)
    rz.set_capacity_mode("exact")
    try:
        eager, dmax = [], 0
        for i, view in enumerate(views):
            _, c, b = r.view_inputs_host(view)
            out = r.render(view, sink=r.flat.grads, cam_dev=c.to(dev), bones_dev=b.to(dev))
            loss = loss_fn(out["render"], targets[i])
            loss.backward()
            eager.append((float(loss.detach()), r.flat.grad.clone(), int(rz.check_overflow())))
            dmax = max(dmax, eager[-1][2])
        rz.set_capacity_mode("reserve", margin=1.2)
        rz.reserve_capacity(dev.index, scene.n, H, W, dmax)
        step = GraphedStep(r, loss_fn, targets[0], view=2, views_in_flight=V, ordered=ordered)
        for rep in range(3):                                    # replays are repeatable (the buffer is cleared / overwritten)
            order = [(rep + i) % 3 for i in range(V)]
            for slot, k in enumerate(order):
                _, c, b = r.view_inputs_host(views[k])
                step.set_inputs(c.to(dev), b.to(dev), targets[k], slot=slot)
            loss = step.replay()
            torch.cuda.synchronize()
            assert step.check() == sum(eager[k][2] for k in order)
            want_loss = sum(eager[k][0] for k in order)
            assert abs(float(loss) - want_loss) <= 2e-6 * abs(want_loss)
            want = sum(eager[k][1] for k in order)
            ok, e, s = grad_close(r.flat.grad.cpu().numpy(), want.cpu().numpy())
            assert ok, (V, rep, e, s)
            for slot, k in enumerate(order):
                assert abs(float(step.losses[slot]) - eager[k][0]) <= 1e-6 * abs(eager[k][0])
    finally:
        rz.set_capacity_mode("exact")


@pytest.mark.parametrize("V,chunks,deferred,fused", [(1, 1, None, True), (2, 3, None, True), (3, 5, None, True), (3, 4, 1, True), (3, 2, 2, True),
                                                   (3, 3, None, False), (2, 2, 1, False)])
def test_pipelined_step_equals_the_sum_of_single_view_gradients(built_lib, V, chunks, deferred, fused):
    """PipelinedStep: tile backward of every view inside one graph, pose backward afterwards range by range over the Gaussians
    (view 0 overwrites a range, the others add) -- the single-rank half of the data-parallel step that all-reduces a finished
    range while the next one computes.  Same gradients as accumulating the views one by one; ranges that do not divide N."""
    from manus_b200 import rasterizer as rz
    from manus_b200.dist import PipelinedStep, gaussian_chunks

    scene, r = _renderer(n=6007)
    H, W, dev = r.H, r.W, r.device
    targets = [torch.rand(H, W, 3, generator=torch.Generator().manual_seed(30 + i)).to(dev) for i in range(3)]
    loss_fn = lambda image, target: (image * target).sum()
    views = (7, 11, 2)
    assert gaussian_chunks(6007, 5) == [(0, 1280), (1280, 2560), (2560, 3840), (3840, 5120), (5120, 6007)]
    rz.set_capacity_mode("exact")
    try:
        eager, dmax = [], 0
        for i, view in enumerate(views):
            _, c, b = r.view_inputs_host(view)
            out = r.render(view, sink=r.flat.grads, cam_dev=c.to(dev), bones_dev=b.to(dev))
            loss = loss_fn(out["render"], targets[i])
            loss.backward()
            eager.append((float(loss.detach()), r.flat.grad.clone(), int(rz.check_overflow()), out["viewspace_points"].grad.clone()))
            dmax = max(dmax, eager[-1][2])
        rz.set_capacity_mode("reserve", margin=1.2)
        rz.reserve_capacity(dev.index, scene.n, H, W, dmax)
        from manus_b200.densify import GaussianState
        gs = GaussianState(r.flat)
        stats = (gs.xyz_gradient_accum, gs.denom, gs.max_radii2D)
        step = PipelinedStep(r, loss_fn, targets[0], view=2, views_in_flight=V, chunks=chunks, stats=stats, deferred_views=deferred, fused_views=fused)
        ref_stats = GaussianState(r.flat)
        assert len(step.ranges) == chunks
        for rep in range(2):
            order = [(rep + i) % 3 for i in range(V)]
            for slot, k in enumerate(order):
                _, c, b = r.view_inputs_host(views[k])
                step.set_inputs(c.to(dev), b.to(dev), targets[k], slot=slot)
            r.flat.grad.fill_(float("nan"))                      # every element must be overwritten by the step
            loss = step.replay()
            torch.cuda.synchronize()
            assert step.check() == sum(eager[k][2] for k in order)
            want_loss = sum(eager[k][0] for k in order)
            assert abs(float(loss) - want_loss) <= 2e-6 * abs(want_loss)
            want = sum(eager[k][1] for k in order)
            ok, e, s = grad_close(r.flat.grad.cpu().numpy(), want.cpu().numpy())
            assert ok, (V, chunks, rep, e, s)
            for slot, k in enumerate(order):                     # the densification statistic of every view (U1)
                ok, e, s = grad_close(step.viewspace[slot].grad.cpu().numpy(), eager[k][3].cpu().numpy())
                assert ok, ("viewspace", slot, e, s)
                # ... and the statistics the pose backward kernel accumulates from it (gaussian.py:335-338, gaussian_utils.py:461-473)
                radii = step.states[slot].radii
                ref_stats.add_densification_stats(step.viewspace[slot].grad, radii > 0, radii)
            for got, want in zip(stats, (ref_stats.xyz_gradient_accum, ref_stats.denom, ref_stats.max_radii2D)):
                ok, e, s = grad_close(got.cpu().numpy(), want.cpu().numpy(), 2e-6)
                assert ok, ("stats", e, s)
            assert float(gs.denom.max()) == (rep + 1) * V
    finally:
        rz.set_capacity_mode("exact")


@pytest.mark.parametrize("n", [6007, 4096])
def test_fused_projection_and_pose_backward_matches_the_two_kernel_path(built_lib, n):
    """render_fused(fuse_backward=True): ONE autograd node whose backward feeds the blend backward's accumulator rows straight
    into the pose backward kernel (mb_raster_backward_blend + mb_pose_backward_from_raster) -- same image, same screen-space
    gradient, same parameter gradients as mb_raster_backward + mb_pose_backward; overwrite, accumulate and autograd-leaf modes."""
    from manus_b200.render import render_fused

    scene, r = _renderer(n=n)
    dev = r.device
    G = torch.rand(r.H, r.W, 3, generator=torch.Generator().manual_seed(11)).to(dev)
    _, c, b = r.view_inputs_host(11)
    c, b = c.to(dev), b.to(dev)
    res = {}
    for fused in (False, True):
        out = r.render(11, sink=r.flat.grads, cam_dev=c, bones_dev=b, fuse_backward=fused)
        (out["render"] * G).sum().backward()
        res[fused] = (out["render"].detach().clone(), r.flat.grad.clone(), out["viewspace_points"].grad.clone(), out["radii"].clone())
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][3], res[False][3])
    ok, e, s = grad_close(res[True][2].cpu().numpy(), res[False][2].cpu().numpy(), 4e-6)    # float atomics: order noise only
    assert ok, (e, s)
    assert float(res[False][2][:, :2].abs().max()) > 0 and float(res[True][2][:, 2].abs().max()) == 0
    ok, e, s = grad_close(res[True][1].cpu().numpy(), res[False][1].cpu().numpy(), 5e-6)
    assert ok, (e, s)
    # accumulate on top of a random buffer
    base = torch.randn_like(r.flat.grad) * res[False][1].abs().mean()
    r.flat.grad.copy_(base)
    out = r.render(11, sink=r.flat.grads, cam_dev=c, bones_dev=b, fuse_backward=True, accumulate=True)
    (out["render"] * G).sum().backward()
    ok, e, s = grad_close((r.flat.grad - base).cpu().numpy(), res[False][1].cpu().numpy(), 1e-5)
    assert ok, (e, s)
    # without a sink the gradients go to the autograd leaves (and to skin_wts when it requires grad)
    grads = {}
    for fused in (False, True):
        leaves = r.flat.leaves()
        skin = r.skin.clone().requires_grad_(True)
        from manus_b200.cameras import Camera
        cam = r._cams[11][0]
        dcam = Camera(cam.width, cam.height, cam.fovx, cam.fovy, c[0:16].view(4, 4), c[16:32].view(4, 4), c[32:35], None)
        out = render_fused(leaves, skin, r._bone_tf, dcam, r.bg, 3, False, r.n_hand, fuse_backward=fused)
        (out["render"] * G).sum().backward()
        grads[fused] = [l.grad.clone() for l in leaves] + [skin.grad.clone()]
    for got, want in zip(grads[True], grads[False]):
        ok, e, s = grad_close(got.cpu().numpy(), want.cpu().numpy(), 5e-6)
        assert ok, (e, s)
