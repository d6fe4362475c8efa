"""Import the UNMODIFIED reference (``/root/reference/src/neural_graph_mapping``)
on CPU by stubbing the third-party packages that are absent from this image.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``not gpu`` tests that pin ``oracle.restatement`` -- both skip when
``/root/reference`` is absent (it does not exist on the GPU box).

Stubbed: pytorch3d, permutohedral_encoding, yoco, open3d,
rerun, evo, torchmetrics, trimesh, pyrender, matplotlib, wandb.  Only three
pytorch3d functions carry behaviour, restated here from their documented
semantics (pytorch3d @47d5dc88, pyproject.toml:19):

* ``quaternion_invert(q) = q * (1,-1,-1,-1)``  (real-first)
* ``quaternion_apply(q, p)`` = vector part of ``q (0,p) q*`` (Hamilton product)
* ``knn_points(p1, p2, K, return_sorted)`` -> squared distances, indices
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch

REFERENCE_SRC = os.environ.get("NGM_REFERENCE_SRC", "/root/reference/src")

_STUB_ROOTS = (
    "pytorch3d", "permutohedral_encoding", "yoco", "open3d", "rerun", "evo",
    "torchmetrics", "trimesh", "pyrender", "matplotlib", "wandb", "lpips",
)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "neural_graph_mapping"))


class _Anything:
    """Permissive placeholder: any attribute / call / .to() returns another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS and root in _stubbed_roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_stubbed_roots: set = set()


def _quaternion_raw_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_invert(quaternion: torch.Tensor) -> torch.Tensor:
    scaling = torch.tensor([1, -1, -1, -1], device=quaternion.device)
    return quaternion * scaling


def quaternion_apply(quaternion: torch.Tensor, point: torch.Tensor) -> torch.Tensor:
    if point.size(-1) != 3:
        raise ValueError(f"Points are not in 3D, {point.shape}.")
    real_parts = point.new_zeros(point.shape[:-1] + (1,))
    point_as_quaternion = torch.cat((real_parts, point), -1)
    out = _quaternion_raw_multiply(
        _quaternion_raw_multiply(quaternion, point_as_quaternion),
        quaternion_invert(quaternion),
    )
    return out[..., 1:]


def knn_points(p1, p2, K=1, return_sorted=True, **_):
    """Exact brute-force K nearest neighbours; returns (sq_dists, idx, None)."""
    d2 = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    dists, idx = torch.topk(d2, K, dim=-1, largest=False, sorted=True)
    return dists, idx, None


def _install_stubs() -> None:
    finder_installed = any(isinstance(f, _StubFinder) for f in sys.meta_path)
    for root in _STUB_ROOTS:
        if root in sys.modules:
            continue
        try:
            if importlib.util.find_spec(root) is not None and root not in ("wandb",):
                continue  # the real package exists -> use it
        except (ImportError, ValueError):
            pass
        _stubbed_roots.add(root)
    if not finder_installed:
        sys.meta_path.insert(0, _StubFinder())
    # behaviour-carrying pytorch3d pieces
    if "pytorch3d" in _stubbed_roots:
        tr = importlib.import_module("pytorch3d.transforms")
        tr.quaternion_apply = quaternion_apply
        tr.quaternion_invert = quaternion_invert
        knn = importlib.import_module("pytorch3d.ops.knn")
        knn.knn_points = knn_points
    if "permutohedral_encoding" in _stubbed_roots:
        pe = importlib.import_module("permutohedral_encoding")

        class PermutoEncoding(torch.nn.Module):  # placeholder base class; never evaluated
            def __init__(self, *a, **k):
                super().__init__()
                raise RuntimeError("permutohedral_encoding is not installed (parity unpinned)")

        pe.PermutoEncoding = PermutoEncoding


_ref = None


def load():
    """Return a namespace with the reference's own modules (imported once)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_SRC}")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    utils = importlib.import_module("neural_graph_mapping.utils")
    utils.benchmark.enabled = False  # decorator default calls torch.cuda.synchronize (utils.py:66-75)
    ns = types.SimpleNamespace(
        utils=utils,
        camera=importlib.import_module("neural_graph_mapping.camera"),
        models=importlib.import_module("neural_graph_mapping.models"),
        positional_encodings=importlib.import_module("neural_graph_mapping.positional_encodings"),
    )
    # evaluation.py moves torchmetrics objects to "cuda" at import; the stubs absorb that.
    ns.run_mapping = importlib.import_module("neural_graph_mapping.run_mapping")
    ns.evaluation = importlib.import_module("neural_graph_mapping.evaluation")
    utils.benchmark.enabled = False
    _ref = ns
    return ns


def default_config(**overrides) -> dict:
    """The shipped ``config/neural_graph_map.yaml`` + nrgbd camera as a plain dict (CPU),
    with NeRF encoding instead of the un-installable permutohedral one."""
    cfg = {
        "model_type": "neural_graph_mapping.models.NeuralFieldSet",
        "model_kwargs": {
            "dim_points": 3,
            "field_type": "neural_graph_mapping.models.NeuralField",
            "field_kwargs": {
                "encoding_type": "neural_graph_mapping.positional_encodings.PositionalEncodingNeRF",
                "encoding_kwargs": {"dim_in": 3, "num_octaves": 4},
                "num_layers": 2,
                "dim_out": 4,
                "dim_mlp_out": 32,
                "skip_mode": "no",
                "initial_geometry_bias": 0.0,
                "neus_initial_sd": 1.0,
            },
            "num_knn": 2,
            "distance_factor": 10.0,
            "field_radius": 1.0,
            "scale_mode": "unit_cube",
            "outside_value": 1.0,
        },
        "color_factor": 1.0, "geometry_factor": 20.0, "device": "cpu",
        "learning_rate": 1e-3, "field_radius": 1.0, "termination_weight": 0.0,
        "photometric_weight": 1.0, "photometric_loss": "l1", "depth_weight": 1.0,
        "depth_loss": "huber", "freespace_weight": 40.0, "tsdf_weight": 50.0,
        "near_distance": 0.0, "far_distance": 8.0, "freeze_model": False,
        "pixel_block_size": 8192, "block_size": 3000000, "preview_res_factor": 0.3,
        "render_frames": [], "render_frame_freq": 200, "extract_mesh_frame_freq": 100,
        "extract_mesh_frames": [], "extract_mesh_fields": [], "log_iteration_freq": 100,
        "num_iterations_per_frame": 5, "rerun_vis": False, "render_vis": False,
        "rerun_save": None, "rerun_connect_addr": None, "geometry_mode": "nrgbd",
        "truncation_distance": 0.1, "disable_relative_fields": False, "disable_vis": True,
        "loglevel": 30, "num_train_fields": 32, "num_rays_per_field": 512,
        "num_samples_coarse": 8, "num_samples_depth_guided": 16, "range_depth_guided": None,
        "benchmark": False, "adam_eps": 1e-15, "adam_weight_decay": 1e-5,
        "update_mode": "multi_view", "single_field_id": None, "max_depth": None,
        "dataset_type": "neural_graph_mapping.slam_datasets.nrgbd_dataset.NRGBDDataset",
        "dataset_config": {},
    }

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                merge(dst[k], v)
            else:
                dst[k] = v

    merge(cfg, overrides)
    return cfg


class injected_jitter:
    """Context manager replacing ``torch.rand`` with a replay of supplied tensors.

    ``Camera.sample_ijs_uniform`` draws ``torch.rand(*leading, S)`` (camera.py:274), in the
    depth-guided case twice -- coarse first, then guided (run_mapping.py:513,531).
    """

    def __init__(self, *jitters: torch.Tensor):
        self._queue = list(jitters)

    def __enter__(self):
        self._orig = torch.rand

        def fake_rand(*shape, **kw):
            if not self._queue:
                raise RuntimeError("more torch.rand draws than injected jitter tensors")
            j = self._queue.pop(0)
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(j.shape) == shape, (j.shape, shape)
            return j.clone()

        torch.rand = fake_rand
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig
        return False
