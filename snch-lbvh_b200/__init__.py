"""snch-lbvh_b200 — B200-native SNCH-LBVH hot path (build + closest-point / closest-silhouette / ray / sphere sampling).

Python host mirror of the reference's ``lbvh::scene<3>`` (scene.cuh:705-1268) on top of the C-ABI in
``include/snch_b200.h`` (``libsnch_b200.so``, hand-written CUDA for sm_100a).  There is NO CPU fallback: importing
this package without the built library, or calling it without a GPU, raises.

The directory name contains a hyphen (it is the name the project was given); ``import snch_lbvh_b200`` works through
the one-line shim module at the repository root.
"""
from .binding import (Scene3, Scene2, Comm, SnchError, lib, lib_path, ABI_SYMBOLS, ExportKind)  # noqa: F401
from . import meshes  # noqa: F401

scene3 = Scene3  # the reference spells it lbvh::scene<3>
scene2 = Scene2  # lbvh::scene<2>
