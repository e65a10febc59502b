"""Deterministic synthetic "monument" workloads (BASELINE.json configs 3-5).

A 4-way-symmetric label grid of edge N built from integer primitives defined at N=256 and scaled by
N/256: plinth box, main-body box, dome sphere, four corner minarets (front pair `front_minarets`,
back pair `back_minarets`), four chhatri cylinders, four small minarets, door and window slabs.
Axes are (z, y, x) with y pointing up, as stage 2 of the reference expects
(utils/voxel_utils.py:17-19: x = a2, y = a1, z = a0).  No RNG in the geometry; all tests are integer
comparisons, so CPU and GPU tensors give identical grids.
"""
from __future__ import annotations

import numpy as np
import torch

from .utils.config import PART_COLORS

PART_NAMES = [k for k in PART_COLORS if k != "background"]            # label = index + 1
LABEL = {name: i + 1 for i, name in enumerate(PART_NAMES)}
CANDIDATE_SEED = 20240607


def label_lut() -> np.ndarray:
    lut = np.zeros((256, 3), np.uint8)
    for name, lab in LABEL.items():
        lut[lab] = PART_COLORS[name]
    return lut


def monument_labels(N: int, device="cpu", chunk: int = 64) -> torch.Tensor:
    """(N,N,N) uint8 label grid, axes (z,y,x)."""
    if N % 32:
        raise ValueError("N must be a multiple of 32")
    dev = torch.device(device)
    q = N // 32                                   # one unit = 8 voxels at N=256

    def u(v):                                     # 256-scale coordinate -> voxels
        return (v * q) // 8

    out = torch.zeros((N, N, N), dtype=torch.uint8, device=dev)
    y = torch.arange(N, device=dev, dtype=torch.int64).view(1, N, 1)
    x = torch.arange(N, device=dev, dtype=torch.int64).view(1, 1, N)
    for z0 in range(0, N, chunk):
        z = torch.arange(z0, min(N, z0 + chunk), device=dev, dtype=torch.int64).view(-1, 1, 1)
        lab = torch.zeros((z.shape[0], N, N), dtype=torch.uint8, device=dev)

        def box(x0, x1, y0, y1, zz0, zz1):
            return (x >= u(x0)) & (x < u(x1)) & (y >= u(y0)) & (y < u(y1)) & (z >= u(zz0)) & (z < u(zz1))

        def cyl(cx, cz, r, y0, y1):
            return ((x - u(cx)) ** 2 + (z - u(cz)) ** 2 <= u(r) ** 2) & (y >= u(y0)) & (y < u(y1))

        lab[box(28, 228, 0, 24, 28, 228)] = LABEL["plinth"]
        lab[box(68, 188, 24, 120, 68, 188)] = LABEL["full_building"]
        dome = ((x - u(128)) ** 2 + (y - u(116)) ** 2 + (z - u(128)) ** 2 <= u(44) ** 2) & (y >= u(120))
        lab[dome] = LABEL["dome"]
        for cx, cz in ((84, 84), (172, 84), (84, 172), (172, 172)):
            lab[cyl(cx, cz, 10, 120, 150)] = LABEL["chhatris"]
        for cx, cz in ((72, 72), (184, 72), (72, 184), (184, 184)):
            lab[cyl(cx, cz, 4, 120, 168)] = LABEL["small_minarets"]
        for cx, cz, name in ((44, 44, "front_minarets"), (212, 44, "front_minarets"),
                             (44, 212, "back_minarets"), (212, 212, "back_minarets")):
            lab[cyl(cx, cz, 9, 24, 200)] = LABEL[name]
        lab[box(116, 140, 24, 72, 64, 68)] = LABEL["main_door"]
        for x0 in (80, 160):
            lab[box(x0, x0 + 16, 56, 88, 64, 68)] = LABEL["windows"]
        out[z0:z0 + z.shape[0]] = lab
    return out


def monument_rgb(N: int, device="cpu") -> torch.Tensor:
    """(N,N,N,3) uint8 RGB grid in the PART_COLORS palette."""
    lut = torch.from_numpy(label_lut()).to(device)
    return lut[monument_labels(N, device).long()]


def base_camera(N: int, H: int, W: int, view: str = "front") -> np.ndarray:
    """A plausible camera [cam_pos, target, f, cx, cy] framing the monument in an HxW image."""
    s = N / 256.0
    if view == "front":
        cam = (128 * s + 7.3, 70 * s + 3.1, -330 * s)
        tgt = (128 * s, 80 * s, 128 * s)
        f = 0.62 * 458.0 / 200.0 * W
    elif view == "aerial":
        cam = (300 * s, 330 * s, -250 * s)
        tgt = (128 * s, 70 * s, 128 * s)
        f = 0.5 * 560.0 / 200.0 * W
    else:
        raise ValueError(view)
    return np.array([*cam, *tgt, f, W / 2.0 + 1.7, H * 0.62], dtype=np.float64)


def candidates(base_row: np.ndarray, K: int, seed: int = CANDIDATE_SEED) -> np.ndarray:
    """Base camera at index 0 plus K-1 perturbations with the reference's random-search step sizes
    (utils/camera_estimation.py:611-625), drawn from default_rng(seed)."""
    from .utils.camera_estimation import random_candidates
    return random_candidates(base_row, K, np.random.default_rng(seed))
