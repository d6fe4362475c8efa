"""neural_graph_mapping_b200 -- the B200 (sm_100a) ray-render hot path of
KTH-RPL/neural_graph_mapping behind the reference's own Python surface.

Importing the package loads ``libngm_b200.so`` (hand-written CUDA behind the C ABI in
``include/ngm_b200.h``); a missing library is an ImportError -- there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from .camera import Camera
from .models import NeuralField, NeuralFieldSet, get_default_precision, set_default_precision
from .positional_encodings import (PermutohedralEncoding, PositionalEncodingFourier, PositionalEncodingNeRF,
                                   TriplaneEncoding)
from .renderer import Prediction, RenderState, install, quadrature, render_image, render_rays

__all__ = [
    "Camera", "NeuralField", "NeuralFieldSet", "PermutohedralEncoding", "PositionalEncodingFourier",
    "PositionalEncodingNeRF", "TriplaneEncoding", "Prediction", "RenderState", "install", "quadrature",
    "render_image", "render_rays", "set_default_precision", "get_default_precision",
]
