"""omm_b200 -- B200-native Opacity Micro-Map baker behind the SDK's ommCpuBake C ABI.

The product is omm_b200/lib/libomm-b200.so (C++ host + sm_100a CUDA kernels, sources under omm_b200/csrc).
This Python package is only the host-side mirror of the SDK's wrapper used by tests and bench.py.
"""
from . import capi  # noqa: F401
from .baker import Baker, BakeInput, BakeResult, OmmError, Texture  # noqa: F401
from .capi import OmmLib, load_product_library  # noqa: F401
