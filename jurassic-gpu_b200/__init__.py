"""jurassic-gpu_b200: B200-native (sm_100a) implementation of JURASSIC's EGA forward model behind the reference's
formod() interface.  The product is the C ABI in include/ + the CUDA kernels in csrc/; this Python package is the
thin host-side mirror used by tests and bench.py (ctypes bindings, synthetic workload generators).

The directory name contains a hyphen, so import it with importlib.import_module("jurassic-gpu_b200").
"""
from . import abi, core, shard, synth  # noqa: F401
from .core import Context, Control, Group, JrbError, Package, Tables, load_core  # noqa: F401

__all__ = ["abi", "core", "synth", "Context", "Control", "Group", "Package", "Tables", "JrbError", "load_core"]
