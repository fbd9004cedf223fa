"""Seeded synthetic inputs of the shapes BASELINE.json names.  NumPy/SciPy only.

No dataset or checkpoint is available offline, so tests and bench.py use these: "natural-like" images
(smooth colour field + noise; lattice sizes close to real photographs), "iid" images (worst case for
the bilateral lattice), "histo" images (pinkish-white background with darker blobs) and unaries
drawn as -log softmax(3 * N(0,1)).
"""
import numpy as np


def natural_image(H, W, seed=0):
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    gh, gw = H // 32 + 2, W // 32 + 2
    grid = rng.uniform(0, 255, (gh, gw, 3))
    up = ndimage.zoom(grid, (H / gh, W / gw, 1), order=3, mode="nearest", grid_mode=True)
    up = up[:H, :W]
    if up.shape[0] < H or up.shape[1] < W:
        up = np.pad(up, ((0, H - up.shape[0]), (0, W - up.shape[1]), (0, 0)), mode="edge")
    img = up + rng.normal(0, 8, (H, W, 3))
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


def iid_image(H, W, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)


def histo_image(H, W, seed=0, n_blobs=40):
    rng = np.random.default_rng(seed)
    img = np.full((H, W, 3), 244.0) + rng.normal(0, 2.0, (H, W, 3))
    img[..., 1] -= 6
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(n_blobs):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        r = rng.uniform(0.02, 0.12) * min(H, W)
        col = rng.uniform(60, 200, 3) * np.array([1.0, 0.6, 1.0])
        m = np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
        img = img * (1 - m[..., None]) + col * m[..., None]
    img += rng.normal(0, 4.0, (H, W, 3))
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


def random_unary(L, N, seed=0, sharp=3.0):
    """(L, N) float32 energies = -log softmax(sharp * N(0,1))."""
    rng = np.random.default_rng(seed + 7919)
    z = rng.standard_normal((L, N)) * sharp
    z -= z.max(axis=0, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=0, keepdims=True))
    return np.ascontiguousarray((-(z - lse)).astype(np.float32))


def blob_probs(C, H, W, seed=0, n_active=None):
    """(C, H, W) float64 class probabilities with smooth spatial structure; inactive classes are 0
    everywhere (exercises the class sub-selection of 03c_hsn/utilities.py:425)."""
    from scipy import ndimage

    rng = np.random.default_rng(seed + 104729)
    active = np.arange(C) if n_active is None else np.sort(rng.choice(C, n_active, replace=False))
    z = np.zeros((C, H, W))
    for c in active:
        z[c] = ndimage.gaussian_filter(rng.standard_normal((H, W)), sigma=max(2.0, min(H, W) / 16.0)) * 25.0
    e = np.zeros_like(z)
    e[active] = np.exp(z[active] - z[active].max(axis=0, keepdims=True))
    e /= e.sum(axis=0, keepdims=True)
    return e


def voc_like_sizes(n, seed=0):
    """(W, H) pairs drawn from common VOC2012 image sizes."""
    rng = np.random.default_rng(seed + 15485863)
    choices = [(500, 375), (375, 500), (500, 333), (500, 500), (334, 500), (500, 334), (500, 374), (480, 360)]
    idx = rng.integers(0, len(choices), n)
    return [choices[i] for i in idx]


def voc_like_size(i, seed=0):
    """(W, H) of item i of a synthetic VOC-like list (independent of the list length)."""
    choices = [(500, 375), (375, 500), (500, 333), (500, 500), (334, 500), (500, 334), (500, 374), (480, 360)]
    rng = np.random.default_rng([seed, i, 15485863])
    return choices[int(rng.integers(0, len(choices)))]


def gt_map(H, W, C, seed=0, ignore=255, border=4):
    """int32 GT label map with blobs and an `ignore` border (VOC-style 255)."""
    from scipy import ndimage

    rng = np.random.default_rng(seed + 32452843)
    # blobs are drawn on a coarse grid (1/8 resolution) and upsampled by pixel replication
    hc, wc = (H + 7) // 8, (W + 7) // 8
    z = np.stack([ndimage.gaussian_filter(rng.standard_normal((hc, wc)), sigma=max(1.0, min(hc, wc) / 10.0))
                  for _ in range(C)])
    gt = np.kron(z.argmax(axis=0), np.ones((8, 8), np.int64))[:H, :W].astype(np.int32)
    gt[:border] = ignore
    gt[-border:] = ignore
    gt[:, :border] = ignore
    gt[:, -border:] = ignore
    return gt


# ------------------------------------------------------------------------------------------------
# GPU-side generation for the 1449-image sweep (BASELINE config 5): the NumPy generators above cost
# ~0.1 s per VOC-sized item, the sweep itself ~1 ms.  Every item is a pure function of (i, seed) --
# independent of rank, world size and batch composition -- so any sharding sees the same inputs.
# ------------------------------------------------------------------------------------------------
def torch_sweep_item(i, n_labels, device, seed=0):
    """Item i of the synthetic VOC-val-shaped list on `device`:
    (image uint8 (H, W, 3), unary float32 (L, H*W), gt int32 (H, W) with a 255 'ignore' border)."""
    import torch
    import torch.nn.functional as F

    w, h = voc_like_size(i, seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(1_000_003 * (seed + 1) + i)
    # natural-like image: smooth colour field (bicubic upsample of a coarse uniform grid) + N(0, 8) noise
    gh, gw = h // 32 + 2, w // 32 + 2
    grid = torch.rand((1, 3, gh, gw), generator=gen, device=device) * 255.0
    img = F.interpolate(grid, size=(h, w), mode="bicubic", align_corners=False)[0]
    img = img + torch.randn((3, h, w), generator=gen, device=device) * 8.0
    img = img.clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()
    # ground truth: argmax of per-class smooth fields on a 1/8 grid, replicated; 255 border
    hc, wc = (h + 7) // 8, (w + 7) // 8
    z = torch.randn((1, n_labels, hc // 4 + 2, wc // 4 + 2), generator=gen, device=device)
    z = F.interpolate(z, size=(hc, wc), mode="bilinear", align_corners=False)[0]
    gt = z.argmax(0).repeat_interleave(8, 0).repeat_interleave(8, 1)[:h, :w].to(torch.int32).contiguous()
    lab = gt.reshape(-1).long()
    border = 4
    gt[:border] = 255
    gt[-border:] = 255
    gt[:, :border] = 255
    gt[:, -border:] = 255
    # unary: noisy evidence for the GT label, -log softmax
    e = torch.randn((n_labels, h * w), generator=gen, device=device) * 1.5
    e[lab, torch.arange(h * w, device=device)] += 2.0
    unary = (-torch.log_softmax(e, dim=0)).contiguous()
    return img, unary, gt
