"""B200-native tree-likelihood hot path of Bio++ bpp-phyl (ChromEvol fork).

The product is ``lib/libbppgpu.so`` (CUDA kernels + C ABI, see include/bppgpu.h) and the C++ host shim in
``host/`` that mirrors the reference's class surface.  ``capi`` is the ctypes binding used by tests and benches.
"""
__all__ = ["capi"]
