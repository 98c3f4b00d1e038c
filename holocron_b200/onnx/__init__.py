"""ONNX export of inference-form models without the ``onnx`` package: see :mod:`.export`."""
from .export import export_onnx  # noqa: F401

__all__ = ["export_onnx"]
