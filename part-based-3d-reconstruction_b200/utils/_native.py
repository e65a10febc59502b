"""ctypes binding of libp3d_b200.so (the C ABI declared in include/p3d_b200.h).

There is no CPU fallback: importing this module without the built library, or calling
into it without a CUDA device, raises.  PyTorch is used only for device memory and
streams; all signatures at the C boundary are plain pointers and sizes.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P3D_LIB") or os.path.normpath(os.path.join(_HERE, "..", "csrc", "libp3d_b200.so"))

MODE_JOINT = 0
MODE_PER_PART = 1
MAX_PARTS = 32


class P3DError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with "
        "`python part-based-3d-reconstruction_b200/build_native.py` (needs nvcc); "
        "there is no CPU fallback for this package")

lib = ctypes.CDLL(LIB_PATH)

_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

_SIGNATURES = {
    "p3d_version": ([], _i32),
    "p3d_last_error": ([], ctypes.c_char_p),
    "p3d_device_info": ([_vp, _vp, _vp, _vp], _i32),
    "p3d_rgb_to_labels": ([_vp, _i64, _vp, _i32, _vp, _vp], _i32),
    "p3d_labels_to_rgb": ([_vp, _i64, _vp, _vp, _vp], _i32),
    "p3d_points_workspace_bytes": ([_i64], _sz),
    "p3d_points_count": ([_vp, _i64, _vp, _vp, _sz, _vp], _i32),
    "p3d_points_fill": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp], _i32),
    "p3d_strided_occupancy": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "p3d_gather_scale_points": ([_vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp], _i32),
    "p3d_mesh_workspace_bytes": ([_i32, _i32, _i32], _sz),
    "p3d_mesh_count": ([_vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp], _i32),
    "p3d_mesh_emit": ([_vp, _i32, _i32, _i32, _vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp], _i32),
    "p3d_colour_presence_bytes": ([], _sz),
    "p3d_colour_presence": ([_vp, _i64, _vp, _vp], _i32),
    "p3d_colour_lookup": ([_vp, _i64, _vp, _vp, _vp], _i32),
    "p3d_setup_cameras_f64": ([_vp, _i32, _vp, _vp], _i32),
    "p3d_setup_cameras_f32": ([_vp, _i32, _vp, _vp], _i32),
    "p3d_splat_f64": ([_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp], _i32),
    "p3d_points_bbox": ([_vp, _i64, _vp, _vp], _i32),
    "p3d_fast_cameras_f64": ([_vp, _i32, _vp, _i32, _i32, _vp, _vp], _i32),
    "p3d_fast_cameras_f32": ([_vp, _i32, _vp, _i32, _i32, _vp, _vp], _i32),
    "p3d_splat_f32": ([_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp], _i32),
    "p3d_resolve_rgb": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "p3d_partwise_counts_rgb": ([_vp, _vp, _i64, _vp, _i32, _vp, _vp], _i32),
    "p3d_sweep_workspace_bytes": ([_i32, _i32, _i32, _i32, _i32], _sz),
    "p3d_sweep_f64": ([_vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp,
                       _vp], _i32),
    "p3d_sweep_f32": ([_vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp,
                       _vp], _i32),
    "p3d_sweep_ctx_create": ([], _vp),
    "p3d_sweep_ctx_destroy": ([_vp], None),
    "p3d_sweep_ctx_launches": ([_vp], _i32),
    "p3d_sweep_ctx_timing": ([_vp, _i32], _i32),
    "p3d_sweep_ctx_timing_read": ([_vp, _vp, _vp], _i32),
    "p3d_segment_length": ([], _i32),
    "p3d_segments_workspace_bytes": ([_i64], _sz),
    "p3d_segments_count": ([_vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_segments_fill": ([_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp], _i32),
    "p3d_best_pack": ([_vp, _vp, _i64, _vp, _vp], _i32),
    "p3d_best_select": ([_vp, _i32, _vp, _vp], _i32),
    "p3d_depth_workspace_bytes": ([_i32, _i32, _i32], _sz),
    "p3d_depth_buffer_f32": ([_vp, _i64, _vp, _i32, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_depth_buffer_f64": ([_vp, _i64, _vp, _i32, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_visible_f32": ([_vp, _i64, _vp, _vp, ctypes.c_float, _i32, _i32, _vp, _vp], _i32),
    "p3d_part_visible_f64": ([_vp, _i64, _vp, _vp, ctypes.c_double, _i32, _i32, _vp, _vp], _i32),
    "p3d_deform_centres": ([_vp, _i64, _i64, _vp, _vp, _vp], _i32),
    "p3d_deform_points": ([_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp], _i32),
    "p3d_pack_label_bits": ([_vp, _i64, _i32, _vp, _vp], _i32),
    "p3d_deform_sweep_f64": ([_vp, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp,
                              _vp, _vp, _vp], _i32),
    "p3d_deform_sweep_f32": ([_vp, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp,
                              _vp, _vp, _vp], _i32),
    "p3d_deform_scatter": ([_vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_resample_carve": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "p3d_resample_carve_passes": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp], _i32),
    "p3d_fold_table": ([_i32, _i32, _vp, _vp, _vp, _vp, _vp], _i32),
    "p3d_fold_gather": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp], _i32),
    "p3d_global_carve_fold": ([_i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp], _i32),
    "p3d_fold_analyse": ([_vp, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_pack_mask_bits": ([_vp, _i32, _i32, _vp, _i32, _vp], _i32),
    "p3d_global_carve_fold_bits": ([_i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp], _i32),
    "p3d_mask_carve": ([_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp], _i32),
    "p3d_colourise": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_part_carve_fold": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp], _i32),
    "p3d_part_carve_bits_workspace_bytes": ([_i32, _i32, _i32, _i32], _sz),
    "p3d_part_carve_fold_bits": ([_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_carve_fold_bits_slab": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_carve_slab_pass_a": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_carve_slab_pass_a_packed": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_carve_pack_groups": ([_vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp], _i32),
    "p3d_part_carve_slab_pass_b": ([_i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp], _i32),
    "p3d_part_carve_slab_pass_b_peers": ([_i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp, _i32, _vp], _i32),
    "p3d_crop_occupancy": ([_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_accumulate_part": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_paste_component": ([_vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp], _i32),
    "p3d_lr_carve_components": ([_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp,
                                 _vp], _i32),
    "p3d_group_image": ([_vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp], _i32),
    "p3d_colour_mask": ([_vp, _i64, _i32, _i32, _i32, _vp, _vp], _i32),
    "p3d_label6_workspace_bytes": ([_i64], _sz),
    "p3d_label6": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp], _i32),
    "p3d_label8_2d": ([_vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp], _i32),
    "p3d_label26": ([_vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp], _i32),
    "p3d_component_stats": ([_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_label_equals": ([_vp, _i64, _i32, _vp, _vp], _i32),
    "p3d_coords_extremes": ([_vp, _i64, _i32, _vp, _vp, _vp], _i32),
    "p3d_component_extremes": ([_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp], _i32),
    "p3d_recolour_components": ([_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp], _i32),
    "p3d_extrude": ([_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp], _i32),
    "p3d_reorient": ([_vp, _i32, _i32, _i32, _vp, _vp], _i32),
}


def _bind(signatures):
    for name, (argtypes, restype) in signatures.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype


_bind(_SIGNATURES)

# kernel launches issued through this binding (bench.py's `gpu_launches` claim)
launch_count = 0


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.p3d_last_error().decode("utf-8", "replace")
        raise P3DError(f"{what or 'libp3d_b200'} failed (code {rc}): {msg}")


_CUDA_OK = False          # set once a CUDA device has been seen (torch.cuda.is_available() costs ~3 us per call)


def cuda_available() -> bool:
    global _CUDA_OK
    if not _CUDA_OK:
        _CUDA_OK = bool(torch.cuda.is_available())
    return _CUDA_OK


def require_cuda(device=None) -> torch.device:
    """Return the CUDA device to run on; raise (never fall back to the CPU)."""
    if not cuda_available():
        raise P3DError("no CUDA device: this package runs its hot path only on a B200 "
                       "(sm_100a) GPU and has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise P3DError(f"device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def on_device(fn):
    """Decorator for public entry points: run `fn` with the CUDA device of its `device=` argument -- or, without one, of
    its first CUDA tensor argument -- made current.  The C ABI launches on the current device's stream
    (`stream_ptr()`), so a call that names another GPU than the current one must switch to it for its duration;
    otherwise kernels would run on one GPU against memory of another."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    names = list(sig.parameters)
    dev_pos = names.index("device") if "device" in names else -1

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if not args and not kwargs:
            return fn()
        # fast resolution (this wrapper sits on calls that take ~100 us): the `device` argument by keyword or position,
        # else the first CUDA tensor among the positional and keyword arguments
        d = kwargs.get("device")
        if d is None and 0 <= dev_pos < len(args):
            d = args[dev_pos]
        dev = None
        if d is not None:
            dev = d if isinstance(d, torch.device) else torch.device(d)
        else:
            for v in args:
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    dev = v.device
                    break
            else:
                for v in kwargs.values():
                    if isinstance(v, torch.Tensor) and v.is_cuda:
                        dev = v.device
                        break
        if dev is None or dev.type != "cuda" or dev.index is None or not cuda_available() \
                or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def ptr(t) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> ctypes.c_void_p:
    """The current device's current stream as a raw cudaStream_t (the private fast accessor when this torch has it:
    torch.cuda.current_stream() builds a Stream object, ~4 us on calls that take ~30)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(a, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """numpy array / torch tensor -> contiguous device tensor of `dtype` (no copy if already there)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.ascontiguousarray(a)
        t = torch.from_numpy(arr)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def palette_tensor(colors, device) -> torch.Tensor:
    pal = np.ascontiguousarray(np.asarray(colors, dtype=np.int64).reshape(-1, 3))
    if pal.size and (pal.min() < 0 or pal.max() > 255):
        raise ValueError("palette colours must be in 0..255")
    return torch.from_numpy(pal.astype(np.uint8)).to(device)
