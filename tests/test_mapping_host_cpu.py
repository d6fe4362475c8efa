"""Host-side logic of the mapping harness on the CPU (no kernels): the synthetic RGB-D stream's geometry and the field
growth of `_extend_global_map_dict` (ngm/run_mapping.py:267-345: fields on a shifted grid of cell 2 r / sqrt(3) cover
every back-projected depth point; covered points add nothing)."""
import math

import torch


def _setup():
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import mapping

    cam = ngm.Camera(width=64, height=48, fx=55.4256, fy=55.4256, cx=31.5, cy=23.5)
    stream = mapping.SyntheticStream(cam, "cpu", num_frames=20, keyframe_every=4)
    cfg = {
        "model_kwargs": {"dim_points": 3, "field_type": "neural_graph_mapping_b200.models.NeuralField",
                         "field_kwargs": {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                          "encoding_kwargs": {"dim_in": 3, "num_octaves": 4}, "num_layers": 1, "dim_out": 4,
                                          "dim_mlp_out": 32},
                         "num_knn": 2, "distance_factor": 10.0, "field_radius": 1.0, "scale_mode": "unit_cube",
                         "outside_value": 1.0},
        "device": "cpu", "geometry_mode": "nrgbd", "field_radius": 1.0, "truncation_distance": 0.1, "max_keyframes": 8,
    }
    return mapping.MappingLoop(cfg, cam, stream), cam, stream


def test_stream_depth_is_on_the_analytic_surfaces():
    loop, cam, stream = _setup()
    assert [stream.is_keyframe(i) for i in range(6)] == [True, False, False, False, True, False]
    seen_wall = seen_ball = False
    for fid in (0, 7, 19):
        item = stream.frame(fid)
        rgbd, c2w = item["rgbd"], item["c2w"]
        assert rgbd.shape == (48, 64, 4) and (rgbd[..., 3] > 0).all() and (rgbd[..., :3] >= 0).all() and (rgbd[..., :3] <= 1).all()
        assert torch.allclose(c2w[:3, :3] @ c2w[:3, :3].T, torch.eye(3), atol=1e-5) and abs(torch.det(c2w[:3, :3]).item() - 1) < 1e-5
        ij = torch.nonzero(rgbd[..., 3])
        z = rgbd[ij[:, 0], ij[:, 1], 3]
        fx, fy, cx, cy, _ = cam.get_pinhole_camera_parameters(0.0)
        pc = torch.stack(((ij[:, 1].float() - cx) / fx * z, -(ij[:, 0].float() - cy) / fy * z, -z), -1)
        pw = pc @ c2w[:3, :3].T + c2w[:3, 3]
        on_wall = ((pw.abs() - stream.half).abs() < 2e-3).any(-1)
        on_ball = ((pw - stream.sphere_c).norm(dim=-1) - stream.sphere_r).abs() < 2e-3
        assert (on_wall | on_ball).all()
        seen_wall, seen_ball = seen_wall or bool(on_wall.any()), seen_ball or bool(on_ball.any())
    assert seen_wall and seen_ball


def test_field_growth_covers_the_depth_image_once():
    loop, cam, stream = _setup()
    item = stream.frame(0)
    shift = torch.tensor([0.3, 0.7, 0.5])
    n0 = loop._extend_global_map_dict(item["rgbd"][..., 3], 0, item["c2w"], shift=shift)
    assert n0 > 0 and loop._num_fields == n0
    g = loop._global_map_dict
    pos = g["positions"][:n0]
    # every back-projected point lies inside some field's sphere (cell diagonal = 2 r)
    ij = torch.nonzero(item["rgbd"][..., 3])
    z = item["rgbd"][ij[:, 0], ij[:, 1], 3]
    fx, fy, cx, cy, _ = cam.get_pinhole_camera_parameters(0.0)
    pc = torch.stack(((ij[:, 1].float() - cx) / fx * z, -(ij[:, 0].float() - cy) / fy * z, -z), -1)
    pw = pc @ item["c2w"][:3, :3].T + item["c2w"][:3, 3]
    # -- up to the reference's own cell-centre formula, (ijk - shift + 0.5) * cell (run_mapping.py:325, mirrored as it
    # is): it displaces every centre by shift * (cell - 1) from the centre of the cell the points were binned into
    cell = 2 * 1.0 / math.sqrt(3)
    assert torch.cdist(pw, pos).min(dim=-1)[0].max().item() <= 1.0 + (cell - 1.0) * shift.norm().item() + 1e-4
    # grid cells are distinct, orientations are identity quaternions, bookkeeping and parameter tables grew together
    assert torch.cdist(pos, pos).fill_diagonal_(9.0).min().item() > 0.99 * cell
    assert torch.equal(g["orientations"][:n0], torch.tensor([1.0, 0, 0, 0]).expand(n0, 4))
    assert (g["kf_ids"][:n0] == 0).all() and g["positions"].shape[0] >= n0
    assert all(v.shape[0] == n0 for v in loop._model.all_fields_params.values())
    assert all(s["exp_avg"].shape[0] == n0 for s in loop._optim_state.values())
    # the same frame again adds nothing; a later keyframe adds only what it newly sees and keeps the old rows
    assert loop._extend_global_map_dict(item["rgbd"][..., 3], 0, item["c2w"], shift=shift) == 0
    before = pos.clone()
    loop._optim_state["_linears.0.weight"]["exp_avg"][:] = 0.5
    item2 = stream.frame(12)
    n1 = loop._extend_global_map_dict(item2["rgbd"][..., 3], 12, item2["c2w"])
    assert torch.equal(loop._global_map_dict["positions"][:n0], before)
    assert loop._num_fields == n0 + n1
    assert (loop._global_map_dict["kf_ids"][n0:n0 + n1] == 12).all()
    ea = loop._optim_state["_linears.0.weight"]["exp_avg"]
    assert ea.shape[0] == n0 + n1 and (ea[:n0] == 0.5).all() and (ea[n0:] == 0).all()


def test_masked_sum_losses_equal_the_reference_means():
    """`MappingLoop._compute_losses` (masked sums over dense tensors) against the oracle's restatement of the
    reference's `_compute_losses` (compaction, then means; pinned by tests/golden/train_steps.npz): same values, and
    NaN for an empty mask like the reference."""
    from collections import namedtuple

    from neural_graph_mapping_b200 import mapping
    from oracle import training as T

    g = torch.Generator().manual_seed(0)
    F, R = 4, 64
    Pred = namedtuple("Pred", "rgbds color_vars depth_vars term_probs freespace_geometry tsdf_residuals")
    Tgt = namedtuple("Tgt", "rgbds depth_mask term_probs term_mask")
    pred = Pred(torch.rand(F, R, 4, generator=g), torch.rand(F, R, 3, generator=g), torch.rand(F, R, generator=g),
                torch.rand(F, R, generator=g) * 0.4 + 0.7, torch.rand(50, generator=g), torch.rand(70, generator=g))
    tgt = Tgt(torch.rand(F, R, 4, generator=g), torch.rand(F, R, generator=g) > 0.3,
              (torch.rand(F, R, generator=g) > 0.5).float(), torch.rand(F, R, generator=g) > 0.2)

    class Cfg:
        _termination_weight, _photometric_weight, _photometric_loss = 0.3, 1.0, "l1"
        _depth_weight, _depth_loss, _truncation_distance, _freespace_weight, _tsdf_weight = 1.0, "huber", 0.1, 40.0, 50.0

    mine = mapping.MappingLoop._compute_losses(Cfg(), tgt, pred)
    ref = T.compute_losses(T.LossSpec(termination_weight=0.3), pred, tgt.rgbds, tgt.depth_mask, tgt.term_probs, tgt.term_mask)
    for k_mine, k_ref in (("termination", "termination"), ("photometric_l1", "photometric"), ("depth_huber", "depth"),
                          ("freespace", "freespace"), ("tsdf", "tsdf"), ("combined", "combined")):
        assert abs(float(mine[k_mine]) - float(ref[k_ref])) <= 1e-6 * max(1.0, abs(float(ref[k_ref]))), k_mine
    empty = Tgt(tgt.rgbds, torch.zeros(F, R, dtype=torch.bool), tgt.term_probs, tgt.term_mask)
    assert math.isnan(float(mapping.MappingLoop._compute_losses(Cfg(), empty, pred)["combined"]))
    assert math.isnan(float(T.compute_losses(T.LossSpec(termination_weight=0.3), pred, empty.rgbds, empty.depth_mask,
                                             empty.term_probs, empty.term_mask)["combined"]))
