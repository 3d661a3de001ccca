"""Oracle-free property tests of the rasterizer (BASELINE-size scenes; size-independent invariants).  This file sorts after every
oracle-parity file on purpose: a property bound can never hide a parity test under `pytest -x`.
"""
import numpy as np
import pytest
import torch

from helpers import posed_scene
from manus_b200 import synth
from test_gpu_raster import DEV, gpu_forward, small_scene

pytestmark = pytest.mark.gpu


def atomic_noise_bound(run_a, run_b, floor=4e-6):
    """The backward sums its per-pixel partials with fp32 RED.ADD atomics whose order is not fixed, so two launches of the SAME
    call differ by summation-order noise.  Returns per-tensor bounds: 8x the measured same-call noise, with a floor of
    `floor` x max|g| (a same-call pair can happen to agree better than the next pair will)."""
    out = []
    for a, b in zip(run_a, run_b):
        if a is None:
            out.append(None)
            continue
        scale = max(1.0, float(a.abs().max()))
        out.append((max(8.0 * float((a - b).abs().max()), floor * scale), scale))
    return out


def test_forward_is_deterministic_and_backward_noise_is_small(built_lib):
    sc, cam, ps = small_scene(10, N=6000, W=256, H=144)
    from manus_b200.rasterizer import rasterize_backward

    G = torch.rand(3, cam.height, cam.width, device=DEV)
    c1, r1, s1 = gpu_forward(cam, (1, 1, 1), ps, debug=False)
    c2, r2, s2 = gpu_forward(cam, (1, 1, 1), ps, debug=False)
    assert torch.equal(c1, c2) and torch.equal(r1, r2)
    g1, g2 = rasterize_backward(s1, G), rasterize_backward(s2, G)
    for a, b in zip(g1, g2):
        if a is not None:
            scale = max(1.0, float(a.abs().max()))
            # float atomics: order noise only -- the parity tolerance (1e-5 of the tensor's magnitude) bounds it with margin;
            # measured ~1-2e-6 (fp32 eps x sqrt(#terms))
            assert float((a - b).abs().max()) <= 1e-5 * scale
            rel_l2 = float((a - b).double().norm() / a.double().norm().clamp_min(1e-30))
            assert rel_l2 <= 1e-5, rel_l2


@pytest.mark.parametrize("n,W,H", [(300_000, 1920, 1080)])
def test_full_size_properties(built_lib, n, W, H):
    """BASELINE-size checks that need no oracle: blending is linear in colour, the colour gradient is exactly the blend weight
    (so <grad, delta> equals the image change), radii > 0 <=> tiles touched, and num_rendered equals the sum of tile-rect areas."""
    from manus_b200.rasterizer import debug_views, rasterize_backward

    sc = synth.make_composite(n, seed=0)
    cam = synth.camera(0, W, H)
    ps = posed_scene(sc, 5, cam)
    color, radii, st = gpu_forward(cam, (1, 1, 1), ps, debug=False)
    dv = debug_views(st)
    D = st.resolve()
    assert D > n and int((radii > 0).sum()) > 0.9 * n
    ranges = dv["ranges"].cpu().numpy().astype(np.int64)
    lens = ranges[:, 1] - ranges[:, 0]
    assert lens.min() >= 0 and lens.sum() == D
    # per-tile depth order: instance depths are non-decreasing inside every tile range
    # (depth = view-space z of the instance's Gaussian)
    m = torch.tensor(ps["means3D"], device=DEV)
    v = torch.tensor(cam.world_view_transform, device=DEV)
    # same operation order as the kernel (individually rounded, no FMA), so 1-ulp neighbours order identically
    depth = ((m[:, 0] * v[0, 2] + m[:, 1] * v[1, 2]) + m[:, 2] * v[2, 2]) + v[3, 2]
    pl = dv["point_list"].long()
    dd = depth[pl]
    tile_of = torch.repeat_interleave(torch.arange(lens.size, device=DEV), torch.tensor(lens, device=DEV))
    same = tile_of[1:] == tile_of[:-1]
    assert bool(((dd[1:] >= dd[:-1]) | ~same).all())
    # linearity in colour
    rng = np.random.default_rng(0)
    delta = rng.uniform(-0.2, 0.2, ps["colors"].shape).astype(np.float32)
    ps2 = dict(ps, colors=ps["colors"] + delta)
    color2, _, _ = gpu_forward(cam, (1, 1, 1), ps2, debug=False)
    G = torch.rand(3, H, W, device=DEV)
    grads = rasterize_backward(st, G)
    lhs = float((grads[1].double() * torch.tensor(delta, device=DEV).double()).sum())
    rhs = float(((color2.double() - color.double()) * G.double()).sum())
    assert abs(lhs - rhs) <= 2e-4 * max(1.0, abs(rhs)), (lhs, rhs)
    # HWC-strided gradient input gives the same result as the contiguous one
    Ghwc = G.permute(1, 2, 0).contiguous()
    grads2 = rasterize_backward(st, Ghwc.permute(2, 0, 1))
    # bound = measured run-to-run noise of the same (contiguous) call: the strided read feeds identical values, so the two
    # results differ only by the order of the fp32 atomics
    bounds = atomic_noise_bound(grads, rasterize_backward(st, G))
    for a, b, bd in zip(grads, grads2, bounds):
        if a is not None:
            err = float((a - b).abs().max())
            assert err <= bd[0], (err, bd)
            assert err <= 1e-5 * bd[1]
            rel_l2 = float((a - b).double().norm() / a.double().norm().clamp_min(1e-30))
            assert rel_l2 <= 1e-5, rel_l2
