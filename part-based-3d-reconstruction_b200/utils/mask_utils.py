"""Drop-in for the live part of the reference's utils/mask_utils.py.

Mask loading and resizing stay on the host with OpenCV exactly as in the reference (tiny data,
and the bytes produced by cv2.resize are part of the contract -- see the note in
load_and_prepare_masks); the per-pixel colour matching runs on the GPU.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv


def _imread_rgb(path):
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def load_mask(root_path, monument_name, view_name, max_dim=None):
    """mask_utils.py:14-33: read <root>/<name>/masks/<name>_<view>_mask.png as RGB; optional
    nearest-neighbour resize so that the longer side is `max_dim`."""
    import cv2
    path = os.path.join(root_path, monument_name, "masks", f"{monument_name}_{view_name}_mask.png")
    mask = _imread_rgb(path)
    if max_dim is not None:
        h, w = mask.shape[:2]
        s = max_dim / max(h, w)
        mask = cv2.resize(mask, (int(w * s), int(h * s)), interpolation=cv2.INTER_NEAREST)
    return mask


def load_and_prepare_masks(root_path, monument_name, view_name, max_dim, part_colors_np, interior_parts,
                           visualize=False):
    """mask_utils.py:35-87 -> (semantic_mask, semantic_mask_exterior, binary_mask).

    Note: the reference passes cv2.INTER_NEAREST as the third POSITIONAL argument of cv2.resize,
    which is `dst`, so the effective interpolation is the default bilinear one and the resized masks
    contain blended off-palette colours.  That behaviour is reproduced (the carving kernels use a
    dynamic palette), because the carved grid depends on those bytes.
    """
    import cv2
    mask_dir = os.path.join(root_path, monument_name, "masks")
    semantic = _imread_rgb(os.path.join(mask_dir, f"{monument_name}_{view_name}_mask.png"))

    interior = np.zeros(semantic.shape[:2], bool)
    for p in interior_parts:
        interior |= np.all(semantic == part_colors_np[p], axis=-1)
    exterior = semantic.copy()
    exterior[interior] = part_colors_np["full_building"]

    def resize_to_max(img):
        h, w = img.shape[:2]
        s = max_dim / max(h, w)
        return cv2.resize(img, (int(w * s), int(h * s)))        # bilinear, see note above

    semantic_r = resize_to_max(semantic)
    exterior_r = resize_to_max(exterior)
    if monument_name == "Charminar":
        win = os.path.join(mask_dir, f"{monument_name}_{view_name}_mask_win.png")
        if os.path.exists(win):
            semantic_r = resize_to_max(_imread_rgb(win))

    binary = (~np.all(exterior_r == part_colors_np["background"], axis=-1)).astype(np.uint8)

    if visualize:
        import matplotlib.pyplot as plt
        fig, axs = plt.subplots(1, 3, figsize=(12, 4))
        for ax, im, title in zip(axs, (semantic_r, exterior_r, binary), ("Original Mask", "Exterior Mask", "Binary Mask")):
            ax.imshow(im, cmap="gray" if im.ndim == 2 else None)
            ax.set_title(title)
            ax.axis("off")
        plt.tight_layout()
        plt.show()
    return semantic_r, exterior_r, binary


@nv.on_device
def image_labels(image, colours, device) -> torch.Tensor:
    """(H,W,3) uint8 image -> (H,W) u8 labels: 1 + index into `colours`, 0 elsewhere (device tensor)."""
    img = nv.to_device(image, torch.uint8, device)
    if img.dim() != 3 or img.shape[-1] != 3:
        raise ValueError(f"expected an (H,W,3) image, got {tuple(img.shape)}")
    pal = nv.palette_tensor(colours, device) if len(colours) else torch.zeros((0, 3), dtype=torch.uint8, device=device)
    return eng.rgb_to_labels(img, pal)


@nv.on_device
def mask_parts_from_image(image, part_colors, selected_parts, device=None):
    """mask_utils.py:89-97: keep the pixels whose colour is one of the selected part colours, zero
    everything else."""
    dev = nv.require_cuda(device)
    colours = []
    for part in selected_parts:
        c = tuple(int(v) for v in np.asarray(part_colors[part]).reshape(3))
        if c not in colours:
            colours.append(c)
    labels = image_labels(image, colours, dev)
    out = eng.labels_to_rgb(labels, eng.make_lut(nv.palette_tensor(colours, dev)))
    if isinstance(image, torch.Tensor):
        return out
    return out.cpu().numpy().astype(np.asarray(image).dtype, copy=False)
