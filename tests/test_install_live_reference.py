"""install(NeuralGraphMap) against the LIVE reference class (skipped where /root/reference is absent, e.g. on the
GPU box): every patched method keeps the reference's parameter names and order, the driver still constructs, and
a CPU call fails loudly instead of falling back."""
import inspect

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

RENDER = ["_render_ijs", "_quadrature", "render_image"]
TRAIN = ["_set_vmap_fields", "_update_step"]
TARGETS = ["_sample_target_mv", "_get_observed_fields"]


def _unwrap(fn):
    """Through torch.no_grad (functools.wraps) and the reference's utils.benchmark (a bare closure, utils.py:60-86)."""
    while True:
        if hasattr(fn, "__wrapped__"):
            fn = fn.__wrapped__
        elif fn.__name__ == "wrapper" and fn.__closure__:
            fn = next(c.cell_contents for c in fn.__closure__ if callable(c.cell_contents))
        else:
            return fn


def _params(fn):
    return [p for p in inspect.signature(_unwrap(fn)).parameters.values() if p.name not in ("self", "driver")]


def test_install_replaces_methods_with_compatible_signatures():
    import neural_graph_mapping_b200 as ngm

    ref = ref_loader.load()
    base = ref.run_mapping.NeuralGraphMap

    class Patched(base):
        pass

    ngm.install(Patched)
    for name in RENDER + TRAIN + TARGETS:
        new, old = getattr(Patched, name), getattr(base, name)
        assert new is not old and new.__module__.startswith("neural_graph_mapping_b200"), name
        old_p, new_p = _params(old), _params(new)
        # same leading parameters (names, order); anything extra on our side must be optional
        assert [p.name for p in new_p[:len(old_p)]] == [p.name for p in old_p], (name, old_p, new_p)
        for o, n in zip(old_p, new_p):
            # a parameter the reference requires may be optional here, never the other way round
            assert o.default is inspect.Parameter.empty or n.default is not inspect.Parameter.empty, (name, o.name)
        assert all(p.default is not inspect.Parameter.empty for p in new_p[len(old_p):]), name

    class RenderOnly(base):
        pass

    ngm.install(RenderOnly, optimizer=False, targets=False)
    for name in TRAIN + TARGETS:
        assert getattr(RenderOnly, name) is getattr(base, name), name
    for name in RENDER:
        assert getattr(RenderOnly, name) is not getattr(base, name), name
    assert all(_unwrap(getattr(base, n)).__module__.startswith("neural_graph_mapping.") for n in RENDER + TRAIN + TARGETS)


def test_patched_driver_has_no_cpu_fallback():
    import neural_graph_mapping_b200 as ngm

    ref = ref_loader.load()

    class Patched(ref.run_mapping.NeuralGraphMap):
        pass

    ngm.install(Patched)
    m = Patched(ref_loader.default_config())
    m._optimizer = None
    m._model.add_fields(2)
    m._global_map_dict["positions"][:2] = torch.tensor([[0.0, 0.0, -2.0], [0.5, 0.0, -2.0]])
    m._global_map_dict["orientations"][:2] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    m._global_map_dict["num"] = 2
    cam = ref.camera.Camera(width=64, height=48, fx=50.0, fy=50.0, cx=31.5, cy=23.5)
    m._camera = cam  # fit() takes it from the dataset
    ijs = torch.zeros(2, 4, 2, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU"):
        m._render_ijs(ijs, torch.eye(4), cam, torch.tensor([0, 1]), True)
    with pytest.raises(RuntimeError, match="no CPU"):
        m._quadrature(torch.rand(3, 5, 3), torch.rand(3, 5), torch.rand(3, 5), torch.rand(3, 5), None)
    with pytest.raises(RuntimeError, match="no CPU"):
        m._get_observed_fields(torch.rand(48, 64, 4), torch.eye(4))
