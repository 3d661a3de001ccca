"""Radix sort / scan building blocks through the C ABI (bit-exact: integer work)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_sort(lib, keys, vals, end_bit, n_on_device=False, max_n=None):
    from manus_b200._lib import check, ptr

    n = keys.numel()
    max_n = max_n or n
    dev = keys.device
    ko, vo = torch.full((max(max_n, 1),), 0xdead, dtype=torch.int32, device=dev), torch.full((max(max_n, 1),), 0xbeef, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.mb_sort_workspace_bytes(max_n), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    if n_on_device:
        kin = torch.zeros(max_n, dtype=torch.int32, device=dev); kin[:n] = keys
        vin = torch.zeros(max_n, dtype=torch.int32, device=dev); vin[:n] = vals
        ndev = torch.tensor([n], dtype=torch.int32, device=dev)
        check(lib.mb_radix_sort_pairs(ptr(kin), ptr(vin), ptr(ko), ptr(vo), -1, ptr(ndev), max_n, end_bit, ptr(ws), ws.numel(), stream), "sort")
    else:
        check(lib.mb_radix_sort_pairs(ptr(keys), ptr(vals), ptr(ko), ptr(vo), n, None, n, end_bit, ptr(ws), ws.numel(), stream), "sort")
    torch.cuda.synchronize()
    return ko[:n].cpu().numpy().astype(np.uint32), vo[:n].cpu().numpy().astype(np.uint32)


@pytest.mark.parametrize("n", [1, 31, 4095, 4096, 4097, 70_001, 1_000_003])
@pytest.mark.parametrize("end_bit", [13, 32])
def test_sort_matches_numpy_stable(built_lib, n, end_bit):
    rng = np.random.default_rng(n + end_bit)
    hi = (1 << end_bit) - 1
    k = rng.integers(0, hi + 1, n, dtype=np.uint64).astype(np.uint32)
    if end_bit == 13:
        k = (k % 977).astype(np.uint32)         # many duplicates, like tile ids
    v = np.arange(n, dtype=np.uint32)
    ko, vo = run_sort(built_lib, torch.tensor(k.view(np.int32), device="cuda"), torch.tensor(v.view(np.int32), device="cuda"), end_bit)
    order = np.argsort(k, kind="stable")
    np.testing.assert_array_equal(ko, k[order])
    np.testing.assert_array_equal(vo, v[order])          # stability: equal keys keep input order


def test_sort_with_device_side_count(built_lib):
    rng = np.random.default_rng(5)
    n, cap = 123_457, 200_000
    k = rng.integers(0, 8160, n).astype(np.uint32)
    v = rng.integers(0, 1 << 30, n).astype(np.uint32)
    ko, vo = run_sort(built_lib, torch.tensor(k.view(np.int32), device="cuda"), torch.tensor(v.view(np.int32), device="cuda"), 13,
                      n_on_device=True, max_n=cap)
    order = np.argsort(k, kind="stable")
    np.testing.assert_array_equal(ko, k[order])
    np.testing.assert_array_equal(vo, v[order])


def test_sort_float_depth_keys(built_lib):
    """Positive float bit patterns sort like the floats themselves (depth keys)."""
    rng = np.random.default_rng(9)
    d = rng.uniform(0.2, 3.0, 300_000).astype(np.float32)
    k = d.view(np.uint32)
    ko, vo = run_sort(built_lib, torch.tensor(k.view(np.int32), device="cuda"), torch.arange(d.size, dtype=torch.int32, device="cuda"), 32)
    assert np.all(np.diff(ko.view(np.float32)) >= 0)
    np.testing.assert_array_equal(vo, np.argsort(d, kind="stable").astype(np.uint32))


@pytest.mark.parametrize("n", [4097, 70_001])
def test_sort_with_constant_digits(built_lib, n):
    """Passes whose digit is the same for every key (the exponent byte of the depths of a hand-sized scene) take the
    identity fast path; a pass that is constant except for ONE key must not."""
    rng = np.random.default_rng(n)
    v = np.arange(n, dtype=np.uint32)
    for keys in ((np.uint32(0x3F) << 24) | rng.integers(0, 1 << 16, n).astype(np.uint32),         # digits 2 and 3 constant
                 rng.uniform(1.0, 1.9, n).astype(np.float32).view(np.uint32),                     # digit 3 constant
                 np.full(n, 0x12345678, np.uint32)):                                               # everything constant
        for odd in (False, True):
            k = keys.copy()
            if odd:
                k[n // 2] = 0xFFFFFFFF                   # a culled Gaussian's key
            ko, vo = run_sort(built_lib, torch.tensor(k.view(np.int32), device="cuda"), torch.tensor(v.view(np.int32), device="cuda"), 32)
            order = np.argsort(k, kind="stable")
            np.testing.assert_array_equal(ko, k[order])
            np.testing.assert_array_equal(vo, v[order])
