"""Tensor-level ops backed by the native kernels (CUDA) or the C++ CPU path."""
from .quantization import compression_ratio, dequantize, fake_quantize, quantize, wire_bytes  # noqa: F401
