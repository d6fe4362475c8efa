"""Positional encodings with the reference's class names and constructor arguments
(ngm/positional_encodings.py).  Here they are *parameter holders + descriptors*: the
arithmetic runs inside the CUDA field kernels (csrc/encodings.cuh), selected by ``KIND``.
"""
from __future__ import annotations

import math
from typing import Literal

import numpy as np
import torch


class PositionalEncoding(torch.nn.Module):
    """Base class (ngm/positional_encodings.py:10-16)."""

    KIND = None

    def get_out_dim(self) -> int:
        raise NotImplementedError()

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "encodings are evaluated inside the fused CUDA field kernel; call NeuralField / "
            "NeuralFieldSet (neural_graph_mapping_b200.models) instead of the bare encoding"
        )


class PositionalEncodingNeRF(PositionalEncoding):
    """sin/cos(2^i * pi * x), ngm/positional_encodings.py:219-276."""

    KIND = "nerf"

    def __init__(self, dim_in: int, num_octaves: int = 8, start_octave: int = 0) -> None:
        super().__init__()
        if dim_in != 3:
            raise NotImplementedError("Only 3D points are supported by the CUDA path.")
        self.num_octaves = num_octaves
        self.start_octave = start_octave
        self.dim_in = dim_in

    def get_out_dim(self) -> int:
        return self.dim_in * self.num_octaves * 2


class PositionalEncodingFourier(PositionalEncoding):
    """sin(W x) with optional raw coordinates, ngm/positional_encodings.py:164-216."""

    KIND = "fourier"

    def __init__(self, dim_in: int, dim_out: int, mu: float, sigma: float, raw_coords: bool) -> None:
        super().__init__()
        if dim_in != 3:
            raise NotImplementedError("Only 3D points are supported by the CUDA path.")
        self._linear = torch.nn.Linear(dim_in, dim_out - dim_in if raw_coords else dim_out, False)
        self._dim_out = dim_out
        self._raw_coords = raw_coords
        torch.nn.init.normal_(self._linear.weight, mu, sigma)

    def get_out_dim(self) -> int:
        return self._dim_out


class TriplaneEncoding(PositionalEncoding):
    """Learned triplane features, ngm/positional_encodings.py:69-161."""

    KIND = "triplane"

    def __init__(self, resolution: int = 32, num_components: int = 64, init_scale: float = 0.1,
                 mode: Literal["sum", "product", "concat"] = "sum") -> None:
        super().__init__()
        if mode not in ("sum", "product", "concat"):
            raise ValueError(f"{mode=} is not supported.")
        self.resolution = resolution
        self.num_components = num_components
        self.init_scale = init_scale
        self.mode = mode
        self.plane_coef = torch.nn.Parameter(
            self.init_scale * torch.randn((3, self.num_components, self.resolution, self.resolution))
        )

    def get_out_dim(self) -> int:
        return self.num_components * (3 if self.mode == "concat" else 1)


class PermutohedralEncoding(PositionalEncoding):
    """Permutohedral-lattice multi-resolution hash encoding with the wrapper's kwargs
    (ngm/positional_encodings.py:19-66).  The reference delegates to the third-party
    ``permutohedral_encoding`` CUDA extension whose source is not in the reference tree:
    this implements the published algorithm (see ``oracle/permuto.py``); parity with the
    third-party kernel is UNPINNED."""

    KIND = "permuto"

    def __init__(self, pos_dim: int, log2_hashmap_size: int, nr_levels: int, nr_feat_per_level: int,
                 coarsest_scale: float, finest_scale: float, appply_random_shift_per_level: bool = True,
                 concat_points: bool = False, concat_points_scaling: float = 1.0,
                 init_scale: float = 1e-5) -> None:
        super().__init__()
        if pos_dim != 3:
            raise NotImplementedError("Only 3D points are supported by the CUDA path.")
        if not 1 <= nr_feat_per_level <= 8:
            raise NotImplementedError("nr_feat_per_level must be in [1, 8]")
        self.pos_dim = pos_dim
        self.log2_hashmap_size = log2_hashmap_size
        self.nr_levels = nr_levels
        self.nr_feat_per_level = nr_feat_per_level
        self.concat_points = concat_points
        self.concat_points_scaling = concat_points_scaling
        scale_per_level = np.geomspace(coarsest_scale, finest_scale, num=nr_levels)  # :50
        sf = np.zeros((nr_levels, pos_dim), dtype=np.float64)
        for lvl, sigma in enumerate(scale_per_level):
            for i in range(pos_dim):
                sf[lvl, i] = 1.0 / (math.sqrt((i + 1) * (i + 2)) * sigma)
        self.register_buffer("scale_factor", torch.from_numpy(sf.astype(np.float32)), persistent=False)
        capacity = 2**log2_hashmap_size  # :51
        self.lattice_values = torch.nn.Parameter(
            (torch.rand(nr_levels, capacity, nr_feat_per_level) * 2 - 1) * init_scale
        )
        shift = torch.randn(nr_levels, pos_dim) * 10.0 if appply_random_shift_per_level \
            else torch.zeros(nr_levels, pos_dim)
        self.register_buffer("random_shift_per_level", shift)

    def output_dims(self) -> int:
        return self.nr_levels * self.nr_feat_per_level + (self.pos_dim if self.concat_points else 0)

    def get_out_dim(self) -> int:
        return self.output_dims()
