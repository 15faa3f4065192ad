"""simpleraytracing_b200 -- B200-native X-ray L-buffer ray tracing (drop-in for the per-pixel
hot path of Brandagot/SimpleRayTracing).

The product is the C-ABI shared library ``lib/libsrt_b200.so`` (declared in
``include/srt.h``; CUDA sm_100a kernels + C++ host code under ``csrc/``) and the C++ host
classes under ``include/srt/``.  This Python package is only the thin ctypes binding used
by the tests and by ``bench.py``; it performs no computation and has no CPU fallback: if
the library is missing it raises.
"""
from .capi import (Library, Context, MultiContext, PinnedArray, Camera, MeshDesc, Params, Timing, SrtError, load_library,
                   KERNEL_AUTO, KERNEL_EXACT, KERNEL_FILTERED, KERNEL_AUDIT, KERNEL_CULLED,
                   FLAG_ALL_MESHES_CONTRIBUTE)

__all__ = ["Library", "Context", "MultiContext", "PinnedArray", "Camera", "MeshDesc", "Params", "Timing", "SrtError", "load_library",
           "KERNEL_AUTO", "KERNEL_EXACT", "KERNEL_FILTERED", "KERNEL_AUDIT", "KERNEL_CULLED", "FLAG_ALL_MESHES_CONTRIBUTE"]
