"""Rank-one SH gradient: the backward without the f_rest gradient + mb_sh_grad_from_views must reproduce what the full
backward of every view writes, summed over the views (the compact exchange of the data-parallel step)."""
import numpy as np
import pytest
import torch

from helpers import grad_close, zoom_camera

pytestmark = pytest.mark.gpu


def _renderer(n=5000, W=200, H=120):
    from manus_b200 import synth
    from manus_b200.dist import SceneRenderer, pack_camera

    scene = synth.make_composite(n, seed=4)
    r = SceneRenderer(scene, torch.device("cuda", 0), W, H)
    for view in (3, 9):
        cam = zoom_camera(view, W, H, 1.3)
        r._cams[view] = (cam, torch.from_numpy(pack_camera(cam)), torch.from_numpy(synth.posed_bones(view).reshape(-1).astype("float32")))
    return scene, r


def test_compact_backward_plus_rebuild_equals_full_backward(built_lib):
    from manus_b200.dist import CompactGradExchange
    from manus_b200.pose import sh_grad_from_views

    scene, r = _renderer()
    G = torch.rand(r.H, r.W, 3, generator=torch.Generator().manual_seed(3)).cuda()
    full, gfdc, bones, campos = {}, [], [], []
    for view in (3, 9):
        out = r.render(view, sink=r.flat.grads)
        (out["render"] * G).sum().backward()
        full[view] = r.flat.grad.clone()
        r.flat.grad.fill_(float("nan"))                                   # whatever is not written must not be read
        out = r.render(view, sink=r.flat.grads, compact_sh=True)
        (out["render"] * G).sum().backward()
        n = scene.n
        got = r.flat.grad.clone()
        assert torch.isnan(got[14 * n:]).all()                            # the f_rest gradient (last segment) is not materialised
        keep = torch.ones_like(got, dtype=torch.bool); keep[14 * n:] = False
        ok, e, s = grad_close(got[keep].cpu().numpy(), full[view][keep].cpu().numpy())
        assert ok, (view, e, s)
        gfdc.append(r.flat.grads["f_dc"].reshape(-1, 3).clone()); bones.append(r._bone_tf.clone()); campos.append(r._last_campos.clone())
        # single-view exchange (world size 1) restores the full gradient buffer
        CompactGradExchange(r)()
        ok, e, s = grad_close(r.flat.grad.cpu().numpy(), full[view].cpu().numpy())
        assert ok, ("exchange", view, e, s)
    # two views at once, as two ranks would contribute them
    expect = full[3] + full[9]
    out_dc, out_rest = torch.empty_like(r.flat.grads["f_dc"]), torch.empty_like(r.flat.grads["f_rest"])
    sh_grad_from_views(r.flat.params["xyz"], r.skin, r.n_hand, 3, 16, torch.stack(bones), torch.stack(campos), torch.stack(gfdc), out_dc, out_rest)
    n = scene.n
    ok, e, s = grad_close(out_dc.reshape(-1).cpu().numpy(), expect[11 * n: 14 * n].cpu().numpy())
    assert ok, ("f_dc", e, s)
    ok, e, s = grad_close(out_rest.reshape(-1).cpu().numpy(), expect[14 * n:].cpu().numpy())
    assert ok, ("f_rest", e, s)
