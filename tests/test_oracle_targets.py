"""oracle/targets.py against the golden output of the unmodified reference's _sample_target_mv
(tests/golden/target_mv.npz, made by oracle/make_target_fixture.py; ngm/run_mapping.py:1261-1459)."""
import torch

import golden_util as G
from oracle import targets as T


def _inputs(meta, a):
    cam = G.camera_spec(meta["camera"])
    draws = {k[len("draw:"):]: v for k, v in a.items() if k.startswith("draw:")}
    return cam, draws


def test_target_sampling_matches_reference():
    meta, a = G.load("target_mv")
    cam, draws = _inputs(meta, a)
    t = T.sample_target_mv(cam, a["c2ws"], a["rgbds"], a["frame_to_store"], a["positions"], meta["num_fields"],
                           a["current_field_ids"], meta["num_train_fields"], meta["field_radius"], draws)
    for k in t._fields:
        ours, ref = getattr(t, k), a["out:" + k]
        assert ours.shape == ref.shape, k
        if ref.dtype.is_floating_point:
            assert torch.allclose(ours, ref, atol=1e-5, rtol=1e-5), k
        else:
            assert torch.equal(ours.to(ref.dtype), ref), k
    # the fixture exercises the visibility filter: some chosen field was dropped because no keyframe sees it
    chosen = T.choose_fields(a["current_field_ids"], meta["num_train_fields"], meta["num_fields"],
                             draws["subset_observed"], draws["subset_random"])
    assert len(chosen) > len(a["out:field_ids"])
    assert not a["out:depth_mask"].all() and a["out:depth_mask"].any()
    assert not a["out:rgb_mask"].all() and not a["out:term_mask"].all()


def test_fixture_is_not_on_a_rounding_edge():
    """The GPU kernels round differently from torch's einsum in the last bit; the fixture must not depend on it:
    the same computation in float64 yields the same pixels, masks and keyframe visibility."""
    meta, a = G.load("target_mv")
    cam, draws = _inputs(meta, a)
    d64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in draws.items()}
    t = T.sample_target_mv(cam, a["c2ws"].double(), a["rgbds"].double(), a["frame_to_store"], a["positions"].double(),
                           meta["num_fields"], a["current_field_ids"], meta["num_train_fields"], meta["field_radius"], d64)
    for k in ("ijs", "field_ids", "rgb_mask", "depth_mask", "term_mask"):
        assert torch.equal(getattr(t, k).to(a["out:" + k].dtype), a["out:" + k]), k
    assert torch.equal(t.term_probs.float(), a["out:term_probs"])


def test_observed_fields_matches_reference():
    """_get_observed_fields (ngm/run_mapping.py:1643-1670) on the fixture frame, in fp32 and (same answer) fp64."""
    meta, a = G.load("target_mv")
    cam = G.camera_spec(meta["camera"])
    n = meta["num_fields"]
    for dt in (torch.float32, torch.float64):
        obs = T.observed_fields(cam, a["observed:rgbd"][..., 3].to(dt), a["observed:c2w"].to(dt),
                                a["positions"][:n].to(dt), meta["field_radius"], a["observed:draw_subset"])
        assert torch.equal(obs, a["observed:out"]), dt
    assert 0 < len(a["observed:out"]) < n
