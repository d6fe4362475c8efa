"""Host logic of targets.sample_target_mv / get_observed_fields on CPU: the two CUDA entry points are replaced by the
oracle's restatement of the same halves, so what is tested is the composition around them -- field choice, the
data-dependent compaction, the order and shapes of the random draws (ngm/run_mapping.py:1295-1400) -- against the
golden output of the unmodified reference."""
import types

import pytest
import torch

import golden_util as G
from neural_graph_mapping_b200 import targets
from oracle import restatement as R
from oracle import targets as T


def _spec(camera):
    fx, fy, cx0, cy0, _ = camera.get_pinhole_camera_parameters(0.0)
    return R.CameraSpec(width=camera.width, height=camera.height, fx=fx, fy=fy, cx=cx0, cy=cy0, pixel_center=0.0)


@pytest.fixture
def cpu_kernels(monkeypatch):
    def vis(camera, c2ws, rgbds, f2s, positions, field_ids, probes, radius):
        return T.visibility(_spec(camera), c2ws, rgbds, f2s, positions, field_ids, probes, radius)[:3]

    def rays(camera, c2ws, rgbds, f2s, positions, field_ids, cids, uv, lo, hi, radius):
        t = T.rays(_spec(camera), c2ws, rgbds, f2s, positions, field_ids, cids, uv, lo, hi, radius)
        return targets.Target(*[(v.long() if k == "ijs" else v) for k, v in zip(t._fields, t)])

    def observed(camera, depth, pix, c2w, positions, radius):
        flat = torch.nonzero(depth)
        lookup = {int(v): i for i, v in enumerate((flat[:, 0] * camera.width + flat[:, 1]).tolist())}
        subset = torch.tensor([lookup[int(p)] for p in pix.tolist()])
        m = torch.zeros(positions.shape[0], dtype=torch.bool)
        m[T.observed_fields(_spec(camera), depth, c2w, positions, radius, subset)] = True
        return m

    monkeypatch.setattr(targets, "target_visibility", vis)
    monkeypatch.setattr(targets, "target_rays", rays)
    monkeypatch.setattr(targets, "observed_fields", observed)


def _driver(meta, a):
    import neural_graph_mapping_b200 as ngm

    d = types.SimpleNamespace()
    d._device, d._camera, d._field_radius = "cpu", ngm.Camera(**meta["camera"]), meta["field_radius"]
    d._num_train_fields, d._num_rays_per_field = meta["num_train_fields"], meta["num_rays_per_field"]
    d._global_map_dict = {"positions": a["positions"], "num": meta["num_fields"]}
    d._c_c2w_tensor, d._nc_rgbd_tensor, d._frame_cid_to_ncid = a["c2ws"], a["rgbds"], a["frame_to_store"]
    return d


def test_sample_target_mv_composition_golden(cpu_kernels):
    meta, a = G.load("target_mv")
    draws = {k[len("draw:"):]: v for k, v in a.items() if k.startswith("draw:")}
    t = targets.sample_target_mv(_driver(meta, a), a["current_field_ids"], draws)
    for k in t._fields:
        ours, ref = getattr(t, k), a["out:" + k]
        assert ours.shape == ref.shape, k
        if ref.dtype.is_floating_point:
            assert torch.allclose(ours, ref, atol=1e-5, rtol=1e-5), k
        else:
            assert torch.equal(ours.to(ref.dtype), ref), k


def test_random_draws_have_the_references_order_and_shapes(cpu_kernels, monkeypatch):
    """Seeded runs draw what the reference draws only if the same torch functions are called in the same order
    with the same shapes: multinomial, multinomial, randn, multinomial, rand (recorded by the fixture generator)."""
    meta, a = G.load("target_mv")
    calls = []
    for name in ("multinomial", "randn", "rand"):
        orig = getattr(torch, name)

        def wrap(*args, _n=name, _f=orig, **kw):
            out = _f(*args, **kw)
            calls.append((_n, tuple(out.shape)))
            return out

        monkeypatch.setattr(torch, name, wrap)
    torch.manual_seed(0)
    t = targets.sample_target_mv(_driver(meta, a), a["current_field_ids"])
    assert [c[0] for c in calls] == ["multinomial", "multinomial", "randn", "multinomial", "rand"]
    assert calls[0][1] == tuple(a["draw:subset_observed"].shape) and calls[1][1] == tuple(a["draw:subset_random"].shape)
    assert calls[2][1] == (20, 3)
    F = len(t.field_ids)
    assert calls[3][1] == (F, meta["num_rays_per_field"]) and calls[4][1] == (F, meta["num_rays_per_field"], 2)
    # fields no keyframe sees never become targets, whatever was drawn
    assert 8 not in t.field_ids.tolist() and 9 not in t.field_ids.tolist()


def test_get_observed_fields_composition_golden(cpu_kernels):
    meta, a = G.load("target_mv")
    d = _driver(meta, a)
    obs = targets.get_observed_fields(d, a["observed:rgbd"], a["observed:c2w"], {"subset": a["observed:draw_subset"]})
    assert obs.dtype == torch.int64 and torch.equal(obs, a["observed:out"])
    torch.manual_seed(1)
    again = targets.get_observed_fields(d, a["observed:rgbd"], a["observed:c2w"])  # own multinomial draw of 500 pixels
    assert set(again.tolist()) <= set(range(meta["num_fields"])) and 9 not in again.tolist()
