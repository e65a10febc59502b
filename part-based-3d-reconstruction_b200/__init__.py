"""B200-native geometry hot path for part-based 3-D reconstruction.

Drop-in for the reference's `utils` modules on the carving + camera-scoring path:

    import importlib
    p3d = importlib.import_module("part-based-3d-reconstruction_b200")
    from p3d.utils ...                       # or put this directory on sys.path and
    from utils.voxel_carving_utils import global_carve, partwise_carve   # as the notebooks do

All numerics run in hand-written sm_100a CUDA kernels behind the C ABI in
include/p3d_b200.h (csrc/libp3d_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
