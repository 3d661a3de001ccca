"""manus_b200.densify (densification / pruning / optimizer-state surgery on the flat buffers) against golden vectors made
by the reference's own GaussianModel (tests/golden/make_golden_densify.py): bit-exact, including the torch.normal samples."""
import os

import numpy as np
import torch

from helpers import GOLDEN
from manus_b200.densify import GaussianState
from manus_b200.dist import FlatGaussians
from manus_b200.optim import GROUP_OF, FlatAdam

G = np.load(os.path.join(GOLDEN, "densify_golden.npz"))
LRS = {"xyz": 0.00016 * 0.2, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}


def load(prefix):
    n = G[f"{prefix}_xyz"].shape[0]
    flat = FlatGaussians(n, "cpu")
    opt = FlatAdam(flat, LRS)
    from manus_b200.densify import _segments
    ms, vs = _segments(flat, opt.exp_avg), _segments(flat, opt.exp_avg_sq)
    for ref_name, key in GROUP_OF.items():
        flat.params[key].copy_(torch.tensor(G[f"{prefix}_{ref_name}"]).reshape(flat.params[key].shape))
        ms[key].copy_(torch.tensor(G[f"{prefix}_{ref_name}_exp_avg"]).reshape(ms[key].shape))
        vs[key].copy_(torch.tensor(G[f"{prefix}_{ref_name}_exp_avg_sq"]).reshape(vs[key].shape))
    return GaussianState(flat, opt, torch.tensor(G[f"{prefix}_skin"]), percent_dense=0.01)


def check(state, prefix):
    from manus_b200.densify import _segments
    ms, vs = _segments(state.flat, state.opt.exp_avg), _segments(state.flat, state.opt.exp_avg_sq)
    for ref_name, key in GROUP_OF.items():
        ref = G[f"{prefix}_{ref_name}"]
        np.testing.assert_array_equal(state.flat.params[key].numpy().reshape(ref.shape), ref, err_msg=f"{prefix} {ref_name}")
        np.testing.assert_array_equal(ms[key].numpy().reshape(ref.shape), G[f"{prefix}_{ref_name}_exp_avg"], err_msg=f"{prefix} {ref_name} exp_avg")
        np.testing.assert_array_equal(vs[key].numpy().reshape(ref.shape), G[f"{prefix}_{ref_name}_exp_avg_sq"], err_msg=f"{prefix} {ref_name} exp_avg_sq")
    np.testing.assert_array_equal(state.skin_wts.numpy(), G[f"{prefix}_skin"])
    np.testing.assert_array_equal(state.xyz_gradient_accum.numpy(), G[f"{prefix}_accum"])
    np.testing.assert_array_equal(state.denom.numpy(), G[f"{prefix}_denom"])
    np.testing.assert_array_equal(state.max_radii2D.numpy(), G[f"{prefix}_max_radii2D"])


def test_densify_prune_reset_match_reference_bit_for_bit():
    st = load("before")
    for v in range(3):
        st.add_densification_stats(torch.tensor(G[f"view{v}_grad"]), torch.tensor(G[f"view{v}_filter"]), torch.tensor(G[f"view{v}_radii"]))
    check(st, "stats")
    max_grad, min_opacity, extent, max_screen, seed = G["args"]
    torch.manual_seed(int(seed))
    st.densify_and_prune(float(max_grad), float(min_opacity), float(extent), float(max_screen))
    assert st.n == G["after_xyz"].shape[0] != G["before_xyz"].shape[0]
    check(st, "after")
    st.reset_opacity()
    check(st, "reset")


def test_prune_keeps_layout_and_step_count():
    st = load("before")
    st.opt.step_count = 7
    mask = torch.zeros(st.n, dtype=torch.bool)
    mask[::3] = True
    before = st.flat.params["quat"][~mask].clone()
    st.prune_points(mask)
    assert st.n == int((~mask).sum()) and st.opt.numel == st.flat.data.numel() == st.opt.exp_avg.numel()
    assert st.opt.step_count == 7 and torch.equal(st.flat.params["quat"], before)


def test_initialize_parameters_matches_the_reference():
    """tests/golden/init_golden.npz: GaussianModel.initialize_parameters (gaussian.py:99-127) run here with the exact 3-NN
    statistic standing in for the absent distCUDA2; bit-exact (same torch operations on the same inputs)."""
    import os

    import numpy as np

    from manus_b200.densify import initialize_parameters

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_golden.npz"))
    assert float(g["dist2"].min()) == 0.0                                  # coincident points exercise the 1e-7 clamp
    for iso in (False, True):
        got = initialize_parameters(g["points"], g["colors"], sh_degree=3, isotropic=iso, dist2=torch.tensor(g["dist2"]))
        pre = "iso" if iso else "aniso"
        for k, v in got.items():
            want = g[f"{pre}_{k}"]
            assert tuple(v.shape) == want.shape, (k, v.shape, want.shape)
            np.testing.assert_array_equal(v.numpy(), want, err_msg=k)
    flat_shapes = {k: tuple(v.shape) for k, v in got.items()}
    assert flat_shapes["log_scale"] == (700, 1) and flat_shapes["f_rest"] == (700, 15, 3)
    flat = FlatGaussians.from_params(got)                                    # the flat buffers the render path works on
    assert flat.isotropic and flat.n == 700 and flat.data.numel() == 700 * (3 + 1 + 1 + 4 + 3 + 45)
    for k, v in got.items():
        assert torch.equal(flat.params[k], v)
