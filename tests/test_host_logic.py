"""Host-side logic added in round 2 that needs no GPU: range splitting of the pipelined step, the loss / gradient-seed protocol,
the piece bookkeeping of the multicast exchange, the synthetic 16+1 bone weighting of BASELINE config 1."""
import numpy as np
import pytest
import torch


def test_gaussian_chunks_cover_the_range_and_start_on_tile_boundaries():
    from manus_b200.dist import gaussian_chunks

    for n, c in ((500_000, 4), (6007, 5), (128, 3), (1, 1), (300_001, 7)):
        r = gaussian_chunks(n, c)
        assert r[0][0] == 0 and r[-1][1] == n and len(r) <= c
        assert all(lo % 128 == 0 and lo < hi for lo, hi in r)
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_loss_functions_may_return_their_gradient():
    from manus_b200.dist import _backward, _loss_and_seed

    img = torch.rand(4, 5, 3, requires_grad=True)
    tgt = torch.rand(4, 5, 3)
    loss, root, seed = _loss_and_seed(lambda i, t: (i * t).sum(), img, tgt)
    assert seed is None and root is loss
    _backward(root, seed)
    g_autograd = img.grad.clone()
    img.grad = None
    x = img * 1.0                                        # the image is a non-leaf in the real step
    loss2, root2, seed2 = _loss_and_seed(lambda i, t: (torch.dot(i.reshape(-1), t.reshape(-1)), t), x, tgt)
    assert root2 is x and seed2 is tgt
    _backward(root2, seed2)
    assert torch.allclose(loss2, loss) and torch.equal(img.grad, g_autograd)


def test_exchange_pieces_address_the_flat_segments():
    """MulticastExchange.pieces(lo, hi): (offset, count) of the six per-parameter slices of Gaussians [lo, hi) in the flat buffer."""
    from manus_b200.dist import PARAM_ORDER, FlatGaussians
    from manus_b200.exchange import MulticastExchange

    flat = FlatGaussians(1000, "cpu")
    ex = MulticastExchange.__new__(MulticastExchange)      # the bookkeeping without a process group / symmetric memory
    ex.flat, ex.offsets, off = flat, {}, 0
    for name in PARAM_ORDER:
        ex.offsets[name] = off
        off += flat.grads[name].numel()
    flat.grad.copy_(torch.arange(flat.grad.numel(), dtype=torch.float32))
    pieces = ex.pieces(256, 640)
    assert sum(c for _, c in pieces) == (640 - 256) * flat.floats_per_gaussian
    for (o, c), name in zip(pieces, PARAM_ORDER):
        want = flat.grads[name][256:640].reshape(-1)
        assert torch.equal(flat.grad[o:o + c], want), name
        assert o % 4 == 0 and c % 4 == 0                   # 16-byte pieces: what the kernels' float4 path takes


def test_merged_bone_weights_keep_rows_normalised():
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench_extras import merged_bones

    w = np.random.default_rng(0).random((50, 21)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    m = merged_bones(w, 16)
    assert m.shape == (50, 17)
    np.testing.assert_allclose(m.sum(1), 1.0, rtol=1e-6)
    np.testing.assert_allclose(m[:, 16], w[:, 20])


def test_stale_step_guard_logic():
    from manus_b200.dist import _check_fresh, _remember_buffers

    class R:
        pass

    class F:
        n = 10

    step, r = R(), R()
    r.flat = F()
    step.r = r
    step.check = lambda: None
    _remember_buffers(step, r)
    _check_fresh(step)
    r.flat = F()                                            # densify / prune replaced the buffers
    with pytest.raises(RuntimeError, match="replaced"):
        _check_fresh(step)


def test_capacity_plans_are_per_device_objects():
    """rasterizer.CapacityPlan: one default plan per device index, module-level helpers address them, a private plan shares nothing."""
    from manus_b200 import rasterizer as rz

    saved = (dict(rz._plans), list(rz._plan_defaults))
    try:
        rz._plans.clear()
        rz.set_capacity_mode("exact")
        p0, p1 = rz.plan_for(0), rz.plan_for(1)
        assert p0 is rz.plan_for(0) and p0 is not p1 and p0.mode == p1.mode == "exact"
        rz.set_capacity_mode("reserve", margin=1.2)                    # every device, and plans created later
        assert (p0.mode, p0.margin, p1.mode) == ("reserve", 1.2, "reserve") and rz.plan_for(2).mode == "reserve"
        rz.reserve_capacity(0, 1000, 48, 64, 5000)
        rz.reserve_capacity(0, 1000, 48, 64, 4000)                     # a high-water mark: never lowered
        assert p0.high_water == {(1000, 48, 64): 5000} and p1.high_water == {}
        rz.set_capacity_mode("exact", device=1)                        # one device only
        assert p1.mode == "exact" and p0.mode == "reserve" and p0.high_water
        rz.set_capacity_mode("reserve", margin=1.3, device=0)          # switching modes forgets the marks
        assert p0.high_water == {}
        mine = rz.CapacityPlan("reserve", 1.5)
        mine.reserve(10, 4, 4, 7)
        assert mine.high_water == {(10, 4, 4): 7} and p0.high_water == {} and mine.last_state is None
        assert rz.check_overflow(device=0) == 0                        # no frame yet
    finally:
        rz._plans.clear()
        rz._plans.update(saved[0])
        rz._plan_defaults[:] = saved[1]
