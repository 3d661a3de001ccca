"""Generate golden vectors for the image losses FROM THE REFERENCE'S OWN PYTHON (needs /root/reference):
    python tests/golden/make_golden_loss.py
Writes tests/golden/loss_golden.npz.

Reference code that runs (no arithmetic restated here): src/utils/loss_utils.py  l1_loss, ssim, _ssim, create_window
called exactly like src/modules/base.py:323-365 does: pred [H,W,3], gt [1,H,W,3] (channel = img.size(-3) = H, so the
11x11 window slides over the (W, 3) plane of every image row), weights 0.8 / 0.2 (config/*.yaml), and torch autograd
through them for d loss / d pred.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402

R.install()
import torch  # noqa: E402

import src.utils.loss_utils as ref_lu  # noqa: E402

torch.set_num_threads(4)


def case(H, W, seed, smooth):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(1, H, W, 3, generator=g)
    if smooth:   # image-like content: low-pass noise + a bright blob on white background
        gt = torch.nn.functional.avg_pool2d(gt.permute(0, 3, 1, 2), 5, 1, 2).permute(0, 2, 3, 1).contiguous()
        gt[:, : H // 3] = 1.0
    pred = (gt[0] + 0.15 * torch.randn(H, W, 3, generator=g)).clamp(0, 1.2).requires_grad_(True)
    l1 = ref_lu.l1_loss(pred, gt, mean=False).mean()             # base.py:329-331
    ss = ref_lu.ssim(pred, gt)                                   # base.py:347
    loss = 0.8 * l1 + 0.2 * (1.0 - ss)                           # config/OBJ_GAUSSIAN.yaml:22-23
    (grad,) = torch.autograd.grad(loss, pred)
    return dict(pred=pred.detach().numpy(), gt=gt[0].numpy(), l1=np.float32(l1.item()), ssim=np.float32(ss.item()),
                loss=np.float32(loss.item()), grad=grad.numpy())


def main():
    out = {}
    for name, (H, W, seed, smooth) in {"a": (37, 53, 0, False), "b": (64, 300, 1, True), "c": (5, 7, 2, False), "d": (130, 16, 3, True)}.items():
        for k, v in case(H, W, seed, smooth).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if k.endswith(("loss", "ssim", "l1"))})


if __name__ == "__main__":
    main()
