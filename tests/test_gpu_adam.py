"""Fused Adam (manus_b200/csrc/adam.cu) against torch.optim.Adam configured like GaussianModel.training_setup
(/root/reference/src/models/gaussian.py:129-141): six param groups, eps 1e-15, per-group learning rates."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LRS = {"xyz": 0.00016 * 0.2, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}


def _setup(n, seed=0):
    from manus_b200.dist import PARAM_ORDER, FlatGaussians

    flat = FlatGaussians(n, torch.device("cuda", 0))
    gen = torch.Generator().manual_seed(seed)
    flat.data.copy_(torch.randn(flat.data.shape, generator=gen))
    ref_params = {k: torch.nn.Parameter(flat.params[k].detach().cpu().clone()) for k in PARAM_ORDER}
    return flat, ref_params, gen


def _torch_adam(ref_params):
    from manus_b200.optim import GROUP_OF

    groups = [{"params": [ref_params[key]], "lr": LRS[name], "name": name} for name, key in GROUP_OF.items()]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15)


@pytest.mark.parametrize("n", [1, 7, 1001])
def test_fused_adam_matches_torch_adam(built_lib, n):
    from manus_b200.dist import PARAM_ORDER
    from manus_b200.optim import FlatAdam

    flat, ref_params, gen = _setup(n)
    opt, ref_opt = FlatAdam(flat, LRS), _torch_adam(ref_params)
    for step in range(5):
        g = torch.randn(flat.grad.shape, generator=gen) * (10.0 ** torch.randint(-6, 2, flat.grad.shape, generator=gen))
        flat.grad.copy_(g)
        off = 0
        for k in PARAM_ORDER:
            cnt = ref_params[k].numel()
            ref_params[k].grad = g[off:off + cnt].reshape(ref_params[k].shape).clone()
            off += cnt
        if step == 3:
            opt.set_lr("xyz", 1e-5)
            ref_opt.param_groups[0]["lr"] = 1e-5
        opt.step()
        ref_opt.step()
        got = flat.data.cpu()
        off = 0
        for k in PARAM_ORDER:
            cnt = ref_params[k].numel()
            ref = ref_params[k].detach().reshape(-1)
            err = (got[off:off + cnt] - ref).abs().max().item()
            assert err <= 2e-7 * max(1.0, ref.abs().max().item()), (k, step, err)
            off += cnt


def test_sharded_steps_equal_full_step(built_lib):
    """Updating the buffer in R disjoint shards (what every rank does for its own slice) == one full step."""
    from manus_b200.optim import FlatAdam, shard_range

    flat_a, _, gen = _setup(333, seed=1)
    flat_b, _, _ = _setup(333, seed=1)
    oa, ob = FlatAdam(flat_a, LRS), FlatAdam(flat_b, LRS)
    for step in range(3):
        g = torch.randn(flat_a.grad.shape, generator=gen)
        flat_a.grad.copy_(g)
        flat_b.grad.copy_(g)
        oa.step(grad_scale=0.25)
        for r in range(3):
            ob.step(shard=shard_range(ob.numel, r, 3), grad_scale=0.25, advance=(r == 0))
        assert torch.equal(flat_a.data, flat_b.data) and torch.equal(oa.exp_avg_sq, ob.exp_avg_sq)
