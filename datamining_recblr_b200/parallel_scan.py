"""Drop-in for the reference's `parallel_scan` module (parallel_scan.py:117-118): same name, same
argument meaning, same assertion-style error behaviour, CUDA kernels instead of Triton."""
from .ops import parallel_scan  # noqa: F401

__all__ = ["parallel_scan"]
