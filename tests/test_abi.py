"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/ngm_b200.h declares, and the ctypes mirror has the C struct sizes.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ngm_b200.h")


@pytest.fixture(scope="module")
def lib_mod():
    import __graft_entry__

    __graft_entry__.build()
    from neural_graph_mapping_b200 import _lib

    return _lib


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^\s*(?:int|size_t|uint64_t|const char\*)\s+(ngm_\w+)\s*\(", src, flags=re.M)))


def test_header_symbols_exported(lib_mod):
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib_mod.lib, name), f"{name} declared in ngm_b200.h but not exported"
    assert sorted(lib_mod.EXPORTS) == declared


def test_abi_version_and_struct_sizes(lib_mod):
    assert lib_mod.lib.ngm_abi_version() == lib_mod.NGM_ABI_VERSION
    for i, s in enumerate(lib_mod.STRUCTS):
        assert lib_mod.lib.ngm_struct_size(i) == ctypes.sizeof(s), s.__name__
    assert lib_mod.lib.ngm_struct_size(99) == 0


def test_argument_errors_without_gpu(lib_mod):
    """Validation happens before any CUDA call, so it is checkable on CPU."""
    a = lib_mod.NgmSampleArgs()
    a.num_rays, a.num_samples = 4, 0
    assert lib_mod.lib.ngm_sample_rays(ctypes.byref(a), None) == -1
    with pytest.raises(ValueError):
        lib_mod.check(-1)
    f = lib_mod.NgmFieldFwdArgs()
    f.field.num_layers = 99
    assert lib_mod.lib.ngm_field_fwd(ctypes.byref(f), None) == -1
    assert b"num_layers" in lib_mod.lib.ngm_last_error()


def test_adam_argument_errors_without_gpu(lib_mod):
    a = lib_mod.NgmAdamArgs()
    a.num_active, a.step, a.lr, a.beta1, a.beta2 = 4, 0, 1e-3, 0.9, 0.999
    assert lib_mod.lib.ngm_adam_step(ctypes.byref(a), None) == -1
    assert b"step" in lib_mod.lib.ngm_last_error()
    a.step, a.beta2 = 1, 1.0
    assert lib_mod.lib.ngm_adam_step(ctypes.byref(a), None) == -1
    assert b"betas" in lib_mod.lib.ngm_last_error()
    a.beta2, a.num_params = 0.999, lib_mod.NGM_ADAM_MAX_PARAMS + 1
    assert lib_mod.lib.ngm_adam_step(ctypes.byref(a), None) == -1
    d = (lib_mod.NgmAdamParam * 1)()
    d[0].row = 8  # pointers missing
    a.params, a.num_params = d, 1
    assert lib_mod.lib.ngm_adam_step(ctypes.byref(a), None) == -1
    assert b"missing" in lib_mod.lib.ngm_last_error()
    a.num_active = 0  # nothing to do: valid, and no launch
    assert lib_mod.lib.ngm_adam_step(ctypes.byref(a), None) == 0


def test_target_argument_errors_without_gpu(lib_mod):
    v = lib_mod.NgmTargetVisArgs()
    v.num_fields, v.num_frames, v.num_probes = 2, 3, 0
    assert lib_mod.lib.ngm_target_visibility(ctypes.byref(v), None) == -1
    v.num_probes = 20
    assert lib_mod.lib.ngm_target_visibility(ctypes.byref(v), None) == -1  # zero focal length
    assert b"camera" in lib_mod.lib.ngm_last_error()
    v.cam = lib_mod.NgmCamera(50.0, 50.0, 31.5, 23.5, 64, 48)
    assert lib_mod.lib.ngm_target_visibility(ctypes.byref(v), None) == -1  # pointers missing
    v.num_fields = 0
    assert lib_mod.lib.ngm_target_visibility(ctypes.byref(v), None) == 0   # nothing to do
    r = lib_mod.NgmTargetRaysArgs()
    r.cam = v.cam
    r.num_fields, r.rays_per_field, r.num_frames = 2, 8, 0
    assert lib_mod.lib.ngm_target_rays(ctypes.byref(r), None) == -1
    r.num_frames = 3
    assert lib_mod.lib.ngm_target_rays(ctypes.byref(r), None) == -1
    assert b"missing" in lib_mod.lib.ngm_last_error()
    r.rays_per_field = 0
    assert lib_mod.lib.ngm_target_rays(ctypes.byref(r), None) == 0
    o = lib_mod.NgmObservedArgs()
    o.cam, o.num_fields, o.num_points, o.pixel_stride = v.cam, 3, 500, 0
    assert lib_mod.lib.ngm_observed_fields(ctypes.byref(o), None) == -1  # stride
    o.pixel_stride = 4
    assert lib_mod.lib.ngm_observed_fields(ctypes.byref(o), None) == -1  # pointers missing
    assert b"missing" in lib_mod.lib.ngm_last_error()
    o.num_fields = 0
    assert lib_mod.lib.ngm_observed_fields(ctypes.byref(o), None) == 0


def test_no_cpu_fallback(lib_mod):
    import torch

    import neural_graph_mapping_b200 as ngm

    fld = ngm.NeuralField("neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                          {"dim_in": 3, "num_octaves": 4}, num_layers=2, dim_out=4, dim_mlp_out=32)
    assert fld.numel() == 1989 - 1  # reference count minus _neus_sd (800+1056+132)
    with pytest.raises(RuntimeError, match="no CPU"):
        fld(torch.rand(8, 3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neural_graph_mapping_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
                assert "/root/reference" not in txt or fn.endswith((".cu", ".cuh", ".h")), fn


def test_build_stamp_does_not_depend_on_where_the_tree_lives(tmp_path, monkeypatch):
    """The snapshot that travels to a GPU box lives under another path: the staleness hash of `_build.py` must be the
    same there (otherwise every process on the box rebuilds the library -- and, under torch.distributed.run, all
    ranks at once)."""
    import shutil

    from neural_graph_mapping_b200 import _build

    here = _build._source_hash("x")
    shutil.copytree(_build.CSRC, tmp_path / "elsewhere" / "csrc")
    shutil.copytree(_build.INCLUDE, tmp_path / "elsewhere" / "include")
    monkeypatch.setattr(_build, "CSRC", str(tmp_path / "elsewhere" / "csrc"))
    monkeypatch.setattr(_build, "INCLUDE", str(tmp_path / "elsewhere" / "include"))
    assert _build._source_hash("x") == here
    (tmp_path / "elsewhere" / "csrc" / "knn.cu").write_text("// changed\n")
    assert _build._source_hash("x") != here
