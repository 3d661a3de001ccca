"""Independent differentiable restatement of the tile rasterizer (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Written separately from oracle/raster_ref.c so that the two can be checked against each other:
the forward follows SURVEY.md Appendix A.1-A.3 with dense [pixels x gaussians] torch tensors, the backward is
torch autograd.  ``upstream_quirks=True`` reproduces the places where upstream's hand-written backward is not
the autograd of its forward (Appendix A.4, items 1-2) by straight-through / detach tricks, so its gradients are
comparable with raster_ref.c on every scene; item 3 (1/(denom^2+1e-7)) is NOT reproduced (relative effect
<= 1.3e-5), item 4-5 hold by construction.

PARITY UNPINNED, like raster_ref.c.  Only for tiny scenes (N <= a few hundred, image <= 64x64).
"""
from __future__ import annotations

import math

import torch

TILE = 16


def rasterize(means3D, opacities, colors, cov6, viewmatrix, projmatrix, tanfovx, tanfovy, W, H, bg,
              upstream_quirks=True, dtype=torch.float64):
    """All tensor inputs may require grad.  Returns (image[3,H,W], radii[N], aux dict with xy (retain_grad'ed))."""
    t64 = lambda a: a.to(dtype)
    m, op, col, c6 = t64(means3D), t64(opacities).reshape(-1), t64(colors), t64(cov6)
    V = t64(viewmatrix).reshape(4, 4)   # row-vector convention: p_view = [p,1] @ V
    P = t64(projmatrix).reshape(4, 4)
    bg = t64(bg).reshape(3)
    N = m.shape[0]
    ph = torch.cat([m, torch.ones(N, 1, dtype=dtype)], 1)
    t = (ph @ V)[:, :3]
    hom = ph @ P
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    focx, focy = W / (2 * tanfovx), H / (2 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = t[:, 2]
    rx, ry = t[:, 0] / tz, t[:, 1] / tz
    if upstream_quirks:
        inx, iny = (rx.abs() <= limx), (ry.abs() <= limy)
        tx = torch.where(inx, t[:, 0], (rx.clamp(-limx, limx) * tz).detach())
        ty = torch.where(iny, t[:, 1], (ry.clamp(-limy, limy) * tz).detach())
    else:
        tx, ty = rx.clamp(-limx, limx) * tz, ry.clamp(-limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([focx / tz, zero, -(focx * tx) / (tz * tz)], -1),
                     torch.stack([zero, focy / tz, -(focy * ty) / (tz * tz)], -1)], 1)          # [N,2,3]
    Rv = V[:3, :3].T                                                                           # world->camera rotation
    Mm = J @ Rv                                                                                # [N,2,3]
    S = torch.stack([c6[:, 0], c6[:, 1], c6[:, 2], c6[:, 1], c6[:, 3], c6[:, 4], c6[:, 2], c6[:, 4], c6[:, 5]], -1).reshape(N, 3, 3)
    cov2 = Mm @ S @ Mm.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    conx, cony, conz = c / det, -b / det, a / det
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    xy = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], -1)
    xy.retain_grad() if xy.requires_grad else None
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    xyd = xy.detach()
    trunc = lambda v: torch.trunc(v).to(torch.int64)
    rminx = trunc((xyd[:, 0] - radius) / TILE).clamp(0, gx); rmaxx = trunc((xyd[:, 0] + radius + TILE - 1) / TILE).clamp(0, gx)
    rminy = trunc((xyd[:, 1] - radius) / TILE).clamp(0, gy); rmaxy = trunc((xyd[:, 1] + radius + TILE - 1) / TILE).clamp(0, gy)
    visible = (tz.detach() > 0.2) & (det.detach() != 0) & ((rmaxx - rminx) * (rmaxy - rminy) > 0)
    radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)

    # depth order compared as float32 bits like upstream (positive floats: same order as values), stable by index
    order = torch.argsort(tz.detach().to(torch.float32), stable=True)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    px, py = xs.reshape(-1).to(dtype), ys.reshape(-1).to(dtype)
    ptx, pty = (xs.reshape(-1) // TILE), (ys.reshape(-1) // TILE)
    T = torch.ones(H * W, dtype=dtype)
    Cacc = torch.zeros(H * W, 3, dtype=dtype)
    done = torch.zeros(H * W, dtype=torch.bool)
    for g in order.tolist():
        if not bool(visible[g]):
            continue
        intile = (ptx >= rminx[g]) & (ptx < rmaxx[g]) & (pty >= rminy[g]) & (pty < rmaxy[g])
        dx, dy = xy[g, 0] - px, xy[g, 1] - py
        power = -0.5 * (conx[g] * dx * dx + conz[g] * dy * dy) - cony[g] * dx * dy
        G = torch.exp(power)
        raw = op[g] * G
        if upstream_quirks:   # value clamps at 0.99, gradient passes as if unclamped
            alpha = raw + (torch.clamp(raw, max=0.99) - raw).detach()
        else:
            alpha = torch.clamp(raw, max=0.99)
        ok = intile & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0) & ~done
        Tn = T * (1 - alpha)
        stop = ok & (Tn.detach() < 1e-4)
        done = done | stop
        use = ok & ~stop
        w = torch.where(use, alpha * T, torch.zeros_like(T))
        Cacc = Cacc + w[:, None] * col[g][None, :]
        T = torch.where(use, Tn, T)
    img = (Cacc + T[:, None] * bg[None, :]).T.reshape(3, H, W)
    return img, radii, dict(xy=xy, conic=torch.stack([conx, cony, conz], -1), depth=tz, final_T=T.reshape(H, W))


def gradients(loss_weights, means3D, opacities, colors, cov6, viewmatrix, projmatrix, tanfovx, tanfovy, W, H, bg,
              upstream_quirks=True):
    """Convenience: loss = sum(image * loss_weights); returns image and the gradients upstream would return."""
    leaves = [x.detach().to(torch.float64).requires_grad_(True) for x in (means3D, opacities, colors, cov6)]
    img, radii, aux = rasterize(leaves[0], leaves[1], leaves[2], leaves[3], viewmatrix, projmatrix, tanfovx, tanfovy,
                                W, H, bg, upstream_quirks)
    loss = (img * loss_weights.to(torch.float64)).sum()
    loss.backward()
    gxy = aux["xy"].grad if aux["xy"].grad is not None else torch.zeros_like(aux["xy"])
    means2D = torch.zeros(means3D.shape[0], 3, dtype=torch.float64)
    means2D[:, 0] = gxy[:, 0] * 0.5 * W
    means2D[:, 1] = gxy[:, 1] * 0.5 * H
    z = lambda g, ref: torch.zeros_like(ref) if g is None else g
    return img.detach(), radii, dict(means3D=z(leaves[0].grad, leaves[0]), opacity=z(leaves[1].grad, leaves[1]).reshape(-1, 1),
                                     colors=z(leaves[2].grad, leaves[2]), cov3D=z(leaves[3].grad, leaves[3]), means2D=means2D)
