"""webrender_b200 — B200-native backend for WebRender's frame-draw hot path.

Holds only what the path needs: the CUDA kernels + C ABI (csrc/, built into
libwrcu.so), the ctypes binding (device.py), the instance/table layouts
(gpu_types.py) and the draw_frame call sequence (frame.py).
"""
from . import abi, gpu_types  # noqa: F401
from .frame import Batch, Clear, Frame, Target, TextureDesc, draw_frame  # noqa: F401
