"""cuauv_vision_pipeline_b200 -- B200-native (sm_100a) implementation of the per-frame pixel hot
path of ayf7/cuauv-vision-pipeline behind the reference's own interfaces.

    csrc/            hand-written CUDA kernels + the C ABI (include/b200vision.h)
    _ffi.py          cffi ABI-mode binding (style of core/bindings/camera_message_framework.py)
    runtime.py       Context + tensor-level operations
    color.py         mirror of utils/color.py            transform.py   mirror of utils/transform.py
    feature.py       blobs (role of utils/feature.py)    color_balance.py  mirror of balance()
    preprocessor.py  mirror of modules/preprocessor.py   yolo_input.py  YOLO letterbox / normalise
    modules.py       drop-in ModuleBase subclasses       sharding.py    stream -> GPU partitioning

Importing the package loads libb200vision.so (building it with nvcc if absent); operations need a
CUDA device and raise otherwise -- there is no CPU path.
"""
from ._ffi import BVError, lib as _lib  # noqa: F401
from .runtime import Context, PinnedArray, default_context, BLOB_DTYPE  # noqa: F401

__version__ = "0.1.0"
