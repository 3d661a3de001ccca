"""Host logic of the view-sharded data-parallel step on CPU with gloo, world_size 2:
R ranks x (views / R) views + ONE all-reduce of the flat gradient buffer == one process accumulating the same views
(the reference's ``final_loss / accum_iter``, /root/reference/src/modules/hand_dynamic.py:248,259-277).
No kernel runs here: ``render_backward`` is a deterministic stand-in that fills the sink like the pose backward does."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manus_b200.dist import PARAM_ORDER, FlatGaussians, allreduce_gradients, shard_views, sharded_step

N, VIEWS = 257, [3, 8, 11, 20, 21, 34, 40]


def fake_render_backward(flat, view, sink):
    """Overwrites every sink tensor with a view-dependent pattern (what one view's backward would write) and returns a loss."""
    for j, name in enumerate(PARAM_ORDER):
        g = sink[name]
        base = torch.arange(g.numel(), dtype=torch.float32).reshape(g.shape)
        g.copy_(torch.sin(base * 0.01 + view) * (j + 1) + flat.params[name] * 0.5)
    return torch.tensor(float(view) * 0.25)


def make_flat():
    flat = FlatGaussians(N, "cpu")
    gen = torch.Generator().manual_seed(0)
    flat.data.copy_(torch.randn(flat.data.shape, generator=gen))
    return flat


def single_process_reference(views):
    flat = make_flat()
    loss = sharded_step(flat, views, fake_render_backward)
    return flat.grad.clone(), float(loss)


def _worker(rank, world, port, views, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = make_flat()
        mine = shard_views(views, rank, world)
        loss = sharded_step(flat, mine, fake_render_backward, global_views=len(views))
        np.save(os.path.join(out_dir, f"grad_{rank}.npy"), flat.grad.numpy())
        np.save(os.path.join(out_dir, f"loss_{rank}.npy"), np.array([float(loss)]))
        # the plain collective helper: SUM then average over ranks
        t = torch.full((5,), float(rank + 1))
        allreduce_gradients(t)
        np.save(os.path.join(out_dir, f"avg_{rank}.npy"), t.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views_round_robin():
    assert shard_views(VIEWS, 0, 2) == [3, 11, 21, 40]
    assert shard_views(VIEWS, 1, 2) == [8, 20, 34]
    assert sorted(sum((shard_views(list(range(50)), r, 8) for r in range(8)), [])) == list(range(50))
    assert [len(shard_views(list(range(50)), r, 8)) for r in range(8)] == [7, 7, 6, 6, 6, 6, 6, 6]


def test_flat_buffer_layout():
    flat = FlatGaussians(10, "cpu")
    assert flat.floats_per_gaussian == 59 and flat.allreduce_bytes() == 10 * 59 * 4
    # views alias the flat buffers, segment per parameter in PARAM_ORDER
    off = 0
    for name in PARAM_ORDER:
        g = flat.grads[name]
        assert g.data_ptr() == flat.grad.data_ptr() + 4 * off and g.is_contiguous()
        off += g.numel()
    assert off == flat.grad.numel()
    iso = FlatGaussians(10, "cpu", sh_coeffs=4, isotropic=True)
    assert iso.floats_per_gaussian == 3 + 3 + 9 + 1 + 1 + 4


@pytest.mark.parametrize("views", [VIEWS, VIEWS[:2], VIEWS[:1]])
def test_two_ranks_equal_gradient_accumulation(tmp_path, views):
    ref_grad, ref_loss = single_process_reference(views)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, views, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"grad_{r}.npy")
        np.testing.assert_allclose(got, ref_grad.numpy(), rtol=2e-6, atol=2e-6)       # fp32 sum order differs only
        assert abs(float(np.load(tmp_path / f"loss_{r}.npy")[0]) - ref_loss) < 1e-6
        np.testing.assert_allclose(np.load(tmp_path / f"avg_{r}.npy"), np.full(5, 1.5))


# ---- ZeRO-1 style optimiser step (manus_b200.optim.sharded_adam_step): host logic on CPU with a stand-in update rule
class _StubOpt:
    """Same interface as FlatAdam, elementwise update p -= lr * g * grad_scale on the requested slice (CPU)."""

    def __init__(self, flat):
        self.flat, self.numel = flat, flat.data.numel()

    def step(self, shard=None, grad=None, grad_scale=1.0, advance=True):
        b, e = (0, self.numel) if shard is None else shard
        g = self.flat.grad if grad is None else grad
        self.flat.data[b:e] -= 0.1 * g[b:e] * grad_scale


def _opt_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from manus_b200.optim import sharded_adam_step

        flat = make_flat()
        flat.grad.copy_(torch.arange(flat.grad.numel(), dtype=torch.float32) * (rank + 1) * 1e-3)
        sharded_adam_step(_StubOpt(flat), n_views=world)
        np.save(os.path.join(out_dir, f"param_{rank}.npy"), flat.data.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_optimizer_step_equals_allreduce_then_full_step(tmp_path):
    world = 2
    mp.spawn(_opt_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    flat = make_flat()
    g = sum(torch.arange(flat.grad.numel(), dtype=torch.float32) * (r + 1) * 1e-3 for r in range(world))
    flat.grad.copy_(g)
    _StubOpt(flat).step(grad_scale=1.0 / world)
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"param_{r}.npy"), flat.data.numpy(), rtol=1e-6, atol=1e-7)


# ---- compact gradient exchange (manus_b200.dist.CompactGradExchange): host logic on CPU with a CPU restatement of the rebuild
C0_SH = 0.28209479177387814


def _rebuild_ref(xyz, skin, n_skinned, sh_degree, sh_coeffs, bone_all, campos_all, gfdc_all, out_f_dc, out_f_rest):
    """sum_r basis(dir_r) x go_r with dir_r from the pinned pose oracle's definition (gaussian_utils.py:431-449)."""
    from oracle import pose_ref

    n = xyz.shape[0]
    acc = torch.zeros(n, sh_coeffs, 3, dtype=torch.float64)
    eye = torch.eye(sh_coeffs, dtype=torch.float64)
    for r in range(gfdc_all.shape[0]):
        campos = campos_all[r].double()
        cam_inv = campos.expand(n, 3).clone()
        if n_skinned:
            tf = torch.einsum("nb,bij->nij", skin.double(), bone_all[r].double())
            hom = torch.cat([campos, torch.ones(1, dtype=torch.float64)])
            cam_inv[:n_skinned] = (torch.linalg.inv(tf) @ hom)[:, :3]
        d = xyz.double() - cam_inv
        d = d / d.norm(dim=1, keepdim=True)
        basis = torch.stack([pose_ref.eval_sh(sh_degree, eye[k].expand(n, 1, sh_coeffs), d)[:, 0] for k in range(sh_coeffs)], 1)   # [n,K]
        acc += basis[:, :, None] * (gfdc_all[r].double() / C0_SH)[:, None, :]
    out_f_dc.copy_(acc[:, :1].float())
    out_f_rest.copy_(acc[:, 1:].float())


class _FakeRenderer:
    def __init__(self, flat, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.flat, self.n_hand, self.sh_degree = flat, flat.n - 40, 3
        w = torch.rand(self.n_hand, 21, generator=g) * (torch.rand(self.n_hand, 21, generator=g) < 0.2)
        w[:, -1] += 0.1
        gs = torch.Generator().manual_seed(7)                     # same skin weights / rest pose on every rank
        w = torch.rand(self.n_hand, 21, generator=gs) * (torch.rand(self.n_hand, 21, generator=gs) < 0.2)
        w[:, -1] += 0.1
        self.skin = w / w.sum(1, keepdim=True)
        self.rest_inv = torch.eye(4).repeat(20, 1, 1)
        tf = torch.eye(4).repeat(21, 1, 1)
        tf[:20, :3, :3] += 0.05 * torch.randn(20, 3, 3, generator=g)      # this rank's pose
        tf[:20, :3, 3] = 0.02 * torch.randn(20, 3, generator=g)
        self._bone_tf = tf
        self._last_campos = torch.tensor([0.3, -0.2, 1.4]) + 0.1 * torch.randn(3, generator=g)


def _compact_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from manus_b200.dist import CompactGradExchange

        flat = make_flat()
        flat.params["xyz"].mul_(0.05)
        g = torch.Generator().manual_seed(rank)
        flat.grad.copy_(torch.randn(flat.grad.shape, generator=g))        # f_rest part: stale values a compact backward leaves behind
        r = _FakeRenderer(flat, rank)
        torch.save(dict(grad=flat.grad.clone(), bone=r._bone_tf, campos=r._last_campos), os.path.join(out_dir, f"in_{rank}.pt"))
        CompactGradExchange(r, rebuild=_rebuild_ref)()
        np.save(os.path.join(out_dir, f"compact_{rank}.npy"), flat.grad.numpy())
    finally:
        dist.destroy_process_group()


def test_compact_exchange_equals_rebuilding_from_all_views(tmp_path):
    world = 2
    mp.spawn(_compact_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ins = [torch.load(tmp_path / f"in_{r}.pt") for r in range(world)]
    flat = make_flat()
    flat.params["xyz"].mul_(0.05)
    n = flat.n
    exp = sum(i["grad"] for i in ins)                               # head and tail: plain sums
    flat.grad.copy_(exp)
    fake = _FakeRenderer(flat, 0)
    gfdc_all = torch.stack([i["grad"][11 * n: 14 * n].reshape(n, 3) for i in ins])
    _rebuild_ref(flat.params["xyz"], fake.skin, fake.n_hand, 3, 16, torch.stack([i["bone"] for i in ins]),
                 torch.stack([i["campos"] for i in ins]), gfdc_all, flat.grads["f_dc"], flat.grads["f_rest"])
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"compact_{r}.npy"), flat.grad.numpy(), rtol=1e-6, atol=1e-6)


def test_compact_exchange_refuses_replaced_buffers():
    """After densify / prune the renderer points at NEW flat buffers (SceneRenderer.rebind); an exchange built on the old ones
    holds views of the stale gradient buffer and must refuse to run."""
    from manus_b200.dist import CompactGradExchange

    flat = make_flat()
    r = _FakeRenderer(flat, 0)
    ex = CompactGradExchange(r, rebuild=_rebuild_ref)
    ex()                                                   # single rank, no process group: runs
    r.flat = make_flat()                                   # what rebind() does
    with pytest.raises(RuntimeError, match="replaced"):
        ex()
