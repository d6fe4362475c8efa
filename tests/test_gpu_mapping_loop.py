"""BASELINE config 3 on the GPU: the mapping loop of neural_graph_mapping_b200/mapping.py (synthetic RGB-D stream ->
field growth -> target sampling -> render under autograd -> losses -> Adam, the reference's
_current_frame_optimization, ngm/run_mapping.py:1124-1251) against the CPU oracle running the SAME iterations with the
same random draws (oracle/targets.py + oracle/training.py, both pinned to outputs of the unmodified reference)."""
import math

import pytest
import torch

from oracle import restatement as R
from oracle import targets as OT
from oracle import training as T

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"

CAM = dict(width=160, height=120, fx=138.5640646, fy=138.5640646, cx=79.5, cy=59.5, pixel_center=0.0)
ENC_KW = {"dim_in": 3, "num_octaves": 4}
L, W, S, G = 2, 32, 8, 8


def _config(precision):
    return {
        "model_type": "neural_graph_mapping_b200.models.NeuralFieldSet",
        "model_kwargs": {
            "dim_points": 3, "field_type": "neural_graph_mapping_b200.models.NeuralField",
            "field_kwargs": {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                             "encoding_kwargs": dict(ENC_KW), "num_layers": L, "dim_out": 4, "dim_mlp_out": W,
                             "skip_mode": "no", "initial_geometry_bias": 0.0, "neus_initial_sd": 1.0},
            "num_knn": 2, "distance_factor": 10.0, "field_radius": 1.0, "scale_mode": "unit_cube", "outside_value": 1.0},
        "color_factor": 1.0, "geometry_factor": 20.0, "device": DEV, "field_radius": 1.0, "learning_rate": 3e-3,
        "adam_eps": 1e-8, "adam_weight_decay": 1e-5, "termination_weight": 0.0, "photometric_weight": 1.0,
        "photometric_loss": "l1", "depth_weight": 1.0, "depth_loss": "huber", "freespace_weight": 40.0, "tsdf_weight": 50.0,
        "near_distance": 0.0, "far_distance": 8.0, "pixel_block_size": 8192, "block_size": 3000000,
        "geometry_mode": "nrgbd", "truncation_distance": 0.1, "num_train_fields": 8, "num_rays_per_field": 64,
        "num_samples_coarse": S, "num_samples_depth_guided": G, "num_iterations_per_frame": 3, "eval_num_samples": 48,
        "single_field_id": None, "precision": precision, "max_keyframes": 48,
    }


def _loop(precision, frames=40, keyframe_every=2):
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import mapping

    cam = ngm.Camera(**CAM)
    stream = mapping.SyntheticStream(cam, DEV, num_frames=frames, keyframe_every=keyframe_every)
    torch.manual_seed(0)
    loop = mapping.MappingLoop(_config(precision), cam, stream)
    # The reference's losses average over rays with term_prob > 0.8 and are NaN when none passes
    # (ngm/run_mapping.py:1787, 1820-1835); its default field starts with an almost constant, small geometry output
    # (permutohedral features of 1e-5: output = last bias), i.e. occupancy ~1 and term_prob ~1 for every ray.  The
    # NeRF test field gets the same start: a small output head.
    with torch.no_grad():
        head = loop._model._prototype_field._linears[-1]
        head.weight.mul_(0.05)
        head.bias.mul_(0.05)
    return loop, cam, stream


def test_synthetic_stream_geometry():
    """Depth is the z-depth of the analytic surface: back-projecting it with the frame's pose lands on the box walls
    or on the sphere."""
    loop, cam, stream = _loop("fp32")
    item = stream.frame(7)
    rgbd, c2w = item["rgbd"], item["c2w"]
    assert rgbd.shape == (CAM["height"], CAM["width"], 4) and (rgbd[..., 3] > 0).all()
    assert torch.allclose(c2w[:3, :3] @ c2w[:3, :3].T, torch.eye(3, device=DEV), atol=1e-5)
    ij = torch.nonzero(rgbd[..., 3])
    z = rgbd[ij[:, 0], ij[:, 1], 3]
    fx, fy, cx, cy, _ = cam.get_pinhole_camera_parameters(0.0)
    pc = torch.stack(((ij[:, 1].float() - cx) / fx * z, -(ij[:, 0].float() - cy) / fy * z, -z), -1)
    pw = pc @ c2w[:3, :3].T + c2w[:3, 3]
    on_wall = ((pw.abs() - stream.half).abs() < 1e-3).any(-1)
    on_ball = ((pw - stream.sphere_c).norm(dim=-1) - stream.sphere_r).abs() < 1e-3
    assert (on_wall | on_ball).all() and on_ball.any() and on_wall.any()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_mapping_loop_learns_the_scene(precision):
    """Fields are created to cover the keyframes, the combined loss falls, and the rendered frame approaches the
    stream's ground truth (PSNR per ngm/evaluation.py:46-56)."""
    loop, cam, stream = _loop(precision)
    first = loop._current_frame_optimization(0)
    assert loop._num_fields > 0 and math.isfinite(float(first["combined"]))
    assert loop._optim_state["_linears.0.weight"]["exp_avg"].shape[0] == loop._num_fields
    before = loop.evaluate_frame(3)
    losses = [float(first["combined"])]
    for f in range(1, 40):
        out = loop._current_frame_optimization(f)
        losses.append(float(out["combined"]))
    after = loop.evaluate_frame(3)
    assert all(math.isfinite(v) for v in losses)
    assert sum(losses[-5:]) / 5 < 0.8 * sum(losses[:5]) / 5, losses
    assert after["psnr"] > before["psnr"] + 1.0 and after["depth_l1"] < before["depth_l1"], (before, after, losses[::5])
    assert loop._global_map_dict["training_iterations"][:loop._num_fields].sum().item() > 0
    assert loop._fps_estimate > 0


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-3), ("fp16", 6e-2)])
def test_mapping_iterations_follow_the_cpu_oracle(precision, tol):
    """The same iterations on the GPU and through the CPU oracle: identical field grid (injected grid shift),
    identical random draws (fields, probes, keyframes, pixels, sampling jitter), losses compared iteration by
    iteration, and the PSNR of a rendered pixel subset after the last one."""
    from neural_graph_mapping_b200 import targets

    loop, cam, stream = _loop(precision, keyframe_every=1)
    cfg = loop._config
    g = torch.Generator().manual_seed(123)
    cs = R.CameraSpec(**CAM)
    fspec = R.FieldSpec("nerf", dict(ENC_KW), L, 4, W, "no")
    rspec = R.RenderSpec(num_samples=S, num_samples_depth_guided=G, range_depth_guided=0.1, truncation_distance=0.1,
                         freespace_weight=40.0, tsdf_weight=50.0, geometry_mode="nrgbd", geometry_factor=20.0,
                         color_factor=1.0, field_radius=1.0, scale_mode="unit_cube")
    lspec = T.LossSpec(termination_weight=0.0, photometric_weight=1.0, photometric_loss="l1", depth_weight=1.0,
                       depth_loss="huber", freespace_weight=40.0, tsdf_weight=50.0, truncation_distance=0.1)
    cell = 2 * 1.0 / math.sqrt(3)
    o_params, o_state = None, None
    gpu_losses, cpu_losses = [], []
    for frame_id in range(3):
        # ---- per-frame state on the GPU, with the grid shift injected; mirrored into the oracle's tables ----
        item = stream.frame(frame_id)
        loop._current_frame_id, loop._current_rgbd, loop._current_c2w = frame_id, item["rgbd"], item["c2w"]
        loop._frame_c2ws[frame_id] = item["c2w"]
        loop._current_is_keyframe = True
        shift = torch.rand(3, generator=g) * cell
        n_before = loop._num_fields
        loop._extend_global_map_dict(item["rgbd"][:, :, 3], frame_id, item["c2w"], shift=shift.to(DEV))
        n = loop._num_fields
        assert n > n_before or frame_id > 0
        positions = loop._global_map_dict["positions"][:n].cpu()
        # every field is at least a grid cell away from the others and covers part of the frame
        if n > 1:
            assert torch.cdist(positions, positions).fill_diagonal_(9.0).min().item() > 0.5 * cell
        orientations = loop._global_map_dict["orientations"][:n].cpu()
        new_rows = {k: v[n_before:n].cpu().clone() for k, v in loop._model.all_fields_params.items()}
        if o_params is None:
            o_params = new_rows
            o_state = T.new_optim_state(o_params)
        elif n > n_before:
            o_params = {k: torch.cat((v, new_rows[k])) for k, v in o_params.items()}
            o_state = T.new_optim_state(o_params, o_state, n - n_before)
        depth = item["rgbd"][..., 3]
        npix = int((depth != 0).sum().item())
        subset = torch.multinomial(torch.ones(npix), 500, generator=g)
        cur = targets.get_observed_fields(loop, item["rgbd"], item["c2w"], {"subset": subset})
        o_cur = OT.observed_fields(cs, depth.cpu(), item["c2w"].cpu(), positions, 1.0, subset)
        assert sorted(cur.tolist()) == sorted(o_cur.tolist())
        loop._current_field_ids = cur
        loop._update_mv_training_data()
        c2ws_store = loop._c_c2w_tensor.cpu()
        rgbd_store = loop._nc_rgbd_tensor.cpu()
        f2s = loop._frame_cid_to_ncid.cpu()
        K = c2ws_store.shape[0]
        for it in range(2):
            # ---- the reference's random draws, in its order (run_mapping.py:1295-1408), from one CPU generator ----
            ntf = cfg["num_train_fields"]
            n_obs = min(ntf // 2, len(o_cur))
            so = torch.multinomial(torch.ones(len(o_cur)), n_obs, generator=g)
            observed = o_cur[so]
            n_rand = min(ntf - len(observed), n - len(observed))
            draws = {"subset_observed": so}
            if n_rand > 0:
                dist = torch.ones(n)
                dist[observed] = 0.0
                draws["subset_random"] = torch.multinomial(dist, n_rand, generator=g)
            draws["probe_offsets"] = torch.randn(20, 3, generator=g)
            ids = OT.choose_fields(o_cur, ntf, n, so, draws.get("subset_random"))
            off = draws["probe_offsets"] / torch.linalg.norm(draws["probe_offsets"], dim=-1, keepdim=True)
            mask, _, _, _ = OT.visibility(cs, c2ws_store, rgbd_store, f2s, positions, ids, off, 1.0)
            fm = mask.any(-1)
            Fv = int(fm.sum().item())
            assert Fv > 0
            Rr = cfg["num_rays_per_field"]
            draws["frame_cids"] = torch.multinomial(mask[fm].float(), Rr, replacement=True, generator=g)
            draws["uv"] = torch.rand(Fv, Rr, 2, generator=g)
            jit, jg = torch.rand(Fv, Rr, S, generator=g), torch.rand(Fv, Rr, G, generator=g)
            # ---- GPU iteration ----
            out = loop._optimization_iteration({k: v.to(DEV) for k, v in draws.items()}, jit.to(DEV), jg.to(DEV))
            gpu_losses.append(float(out["combined"]))
            # ---- oracle iteration ----
            tgt = OT.sample_target_mv(cs, c2ws_store, rgbd_store, f2s, positions, n, o_cur, ntf, 1.0, draws)
            assert torch.equal(tgt.field_ids, loop._target.field_ids.cpu())
            assert torch.equal(tgt.ijs.long(), loop._target.ijs.cpu())
            ol, _, _ = T.training_iteration(o_params, o_state, positions, orientations, tgt.field_ids, cs, rspec, fspec, lspec,
                                            tgt.ijs.long(), tgt.c2ws, tgt.near_distances, tgt.far_distances, tgt.gt_distances,
                                            jit, jg, tgt.rgbds, tgt.depth_mask, tgt.term_probs, tgt.term_mask,
                                            cfg["learning_rate"], cfg["adam_eps"], cfg["adam_weight_decay"])
            cpu_losses.append(float(ol["combined"]))
            assert abs(gpu_losses[-1] - cpu_losses[-1]) <= tol * abs(cpu_losses[-1]), (frame_id, it, gpu_losses, cpu_losses)
    # ---- the two maps render the same pixels: PSNR against the stream's ground truth within 0.1 dB ----
    gsel = torch.Generator().manual_seed(5)
    item = stream.frame(1)
    pix = torch.randint(0, CAM["height"] * CAM["width"], (512,), generator=gsel)
    ijs = torch.stack((pix // CAM["width"], pix % CAM["width"]), -1)
    jit = torch.rand(512, 48, generator=gsel)
    loop.eval()
    with torch.no_grad():
        pg = loop._render_ijs(ijs.to(DEV), item["c2w"], cam, jitter=jit.to(DEV))
    loop.train()
    er = R.RenderSpec(num_samples=48, near_distance=0.0, far_distance=8.0, truncation_distance=0.1, geometry_mode="nrgbd",
                      geometry_factor=20.0, color_factor=1.0, field_radius=1.0, scale_mode="unit_cube", num_knn=2,
                      distance_factor=10.0, outside_value=1.0)
    n = loop._num_fields
    with torch.no_grad():
        po = R.render_rays(ijs, item["c2w"].cpu(), cs, er, fspec, o_params, loop._global_map_dict["positions"][:n].cpu(),
                           loop._global_map_dict["orientations"][:n].cpu(), use_vmap=False, jitter=jit)
    gt = item["rgbd"].cpu()[ijs[:, 0], ijs[:, 1], :3]

    def psnr(x):
        return 10 * math.log10(1.0 / ((x.clamp(0, 1) - gt.clamp(0, 1)) ** 2).mean().item())

    assert abs(psnr(pg.rgbds[:, :3].cpu()) - psnr(po.rgbds[:, :3])) < 0.1, (psnr(pg.rgbds[:, :3].cpu()), psnr(po.rgbds[:, :3]))
