"""Train a small field on an analytic RGB-D scene with the ORACLE (CPU autograd) and store it as a
golden fixture for the PSNR check (north_star: "PSNR within 0.1 dB of the reference").

    python -m oracle.make_trained_fixture        # ~1-2 min on 8 CPU threads

Scene: a coloured sphere (radius 0.35) in front of a wall, seen by the NRGBD camera; one field of
radius 1 covers it.  The field (NeRF-4 encoding, 2-layer x 32 MLP: BASELINE configs[0] architecture) is
trained with the reference's own objective shape (ngm/losses.py + run_mapping.py:1769-1860): L1 colour,
L1 depth, free-space and truncated-SDF terms, `nrgbd` compositing.  The fixture stores the trained
parameters, a 96x72 test image of rays with injected jitter, the analytic ground truth and the
oracle's (= reference arithmetic) render, so that the GPU test can compare
|PSNR(ours, gt) - PSNR(reference, gt)| without the reference or the oracle training code present.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restatement as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CAM = dict(width=640, height=480, fx=554.2562584220408, fy=554.2562584220408, cx=319.5, cy=239.5, pixel_center=0.0)
FIELD_KW = {"encoding_type": "neural_graph_mapping.positional_encodings.PositionalEncodingNeRF",
            "encoding_kwargs": {"dim_in": 3, "num_octaves": 4}, "num_layers": 2, "dim_out": 4, "dim_mlp_out": 32,
            "skip_mode": "no", "initial_geometry_bias": 0.0, "neus_initial_sd": 1.0}
SPHERE_C = torch.tensor([0.1, -0.05, -2.0])
SPHERE_R = 0.35
WALL_Z = -2.6
TRUNC = 0.1


def scene_gt(ijs, cam: R.CameraSpec):
    """Analytic RGB + ray distance of the sphere-in-front-of-a-wall scene (camera at the origin, c2w = I)."""
    dirs = R.ijs_to_directions(ijs, cam)
    oc = -SPHERE_C
    b = (dirs * oc).sum(-1)
    disc = b * b - ((oc * oc).sum() - SPHERE_R**2)
    t_s = torch.where(disc > 0, -b - torch.sqrt(disc.clamp_min(0)), torch.full_like(b, float("inf")))
    t_w = WALL_Z / dirs[..., 2]
    hit_s = t_s < t_w
    t = torch.where(hit_s, t_s, t_w)
    p = dirs * t[..., None]
    n = (p - SPHERE_C) / SPHERE_R
    col_s = 0.5 + 0.5 * n
    col_w = torch.stack([0.5 + 0.4 * torch.sin(3 * p[..., 0]), 0.5 + 0.4 * torch.sin(3 * p[..., 1] + 1.0),
                         0.6 * torch.ones_like(t)], -1)
    rgb = torch.where(hit_s[..., None], col_s, col_w)
    return rgb, t


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    cam = R.CameraSpec(**CAM)
    fs = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 4}, 2, 4, 32, "no")
    rs = R.RenderSpec(num_samples=16, num_samples_depth_guided=16, range_depth_guided=TRUNC, truncation_distance=TRUNC,
                      freespace_weight=40.0, tsdf_weight=50.0, geometry_mode="nrgbd", geometry_factor=20.0,
                      field_radius=1.0, scale_mode="unit_cube")
    g = torch.Generator().manual_seed(1)
    params = R.stack_params([R.init_field_params(fs, g)])
    params["_neus_sd"] = torch.ones(1)
    leaves = {k: v.clone().requires_grad_(k != "_neus_sd") for k, v in params.items()}
    positions = torch.tensor([[0.0, 0.0, -2.2]])
    orientations = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    fid = torch.tensor([0])
    c2w = torch.eye(4)
    opt = torch.optim.Adam([v for v in leaves.values() if v.requires_grad], lr=5e-3)
    for it in range(900):
        ijs = torch.stack([torch.randint(120, 360, (1, 1024), generator=g), torch.randint(160, 480, (1, 1024), generator=g)], -1)
        rgb, t = scene_gt(ijs, cam)
        near, far = (t - 0.6).clamp_min(0.2), t + 0.4
        pred = R.render_rays(ijs, c2w, cam, rs, fs, leaves, positions, orientations, field_ids=fid, use_vmap=True,
                             near_distances=near, far_distances=far, gt_distances=t.clone(),
                             jitter=torch.rand(1, 1024, 16, generator=g), jitter_guided=torch.rand(1, 1024, 16, generator=g))
        dirs = R.ijs_to_directions(ijs, cam)
        gt_depth = t * (-dirs[..., 2])
        loss = (pred.rgbds[..., :3] - rgb).abs().mean() + (pred.rgbds[..., 3] - gt_depth).abs().mean() \
            + 40.0 * ((pred.freespace_geometry - TRUNC) ** 2).mean() + 50.0 * (pred.tsdf_residuals ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it % 100 == 0:
            print(it, float(loss))
    trained = {k: v.detach() for k, v in leaves.items()}

    # test image: 96 x 72 pixel grid over the trained region, eval-style sampling (no depth guidance)
    H, W, S = 72, 96, 64
    ii = torch.linspace(130, 350, H).round().long()
    jj = torch.linspace(170, 470, W).round().long()
    ijs = torch.cartesian_prod(ii, jj).view(1, H * W, 2)
    rgb, t = scene_gt(ijs, cam)
    dirs = R.ijs_to_directions(ijs, cam)
    gt_depth = t * (-dirs[..., 2])
    near, far = (t - 0.6).clamp_min(0.2), t + 0.4
    jit = torch.rand(1, H * W, S, generator=torch.Generator().manual_seed(77))  # regenerated by the test (not stored)
    rs_eval = R.RenderSpec(num_samples=S, num_samples_depth_guided=0, truncation_distance=TRUNC, geometry_mode="nrgbd",
                           geometry_factor=20.0, field_radius=1.0, scale_mode="unit_cube")
    with torch.no_grad():
        pred = R.render_rays(ijs, c2w, cam, rs_eval, fs, trained, positions, orientations, field_ids=fid, use_vmap=True,
                             near_distances=near, far_distances=far, jitter=jit)
    img = pred.rgbds[0, :, :3].reshape(H, W, 3)
    psnr = R.psnr(img, rgb[0].reshape(H, W, 3))
    dl1 = (pred.rgbds[0, :, 3] - gt_depth[0]).abs().mean().item()
    print(f"oracle render vs analytic gt: PSNR {psnr:.2f} dB, depth L1 {dl1:.4f} m")
    meta = {"case": "trained_render", "camera": CAM, "field_kwargs": FIELD_KW, "num_samples": S, "image_hw": [H, W], "jitter_seed": 77,
            "psnr_reference_db": psnr, "depth_l1_reference": dl1,
            "config": {"color_factor": 1.0, "geometry_factor": 20.0, "field_radius": 1.0, "freespace_weight": 40.0,
                       "tsdf_weight": 50.0, "near_distance": 0.0, "far_distance": 8.0, "geometry_mode": "nrgbd",
                       "truncation_distance": TRUNC, "num_samples_coarse": S, "num_samples_depth_guided": 0,
                       "range_depth_guided": None, "block_size": 3000000, "pixel_block_size": 8192,
                       "model_kwargs": {"dim_points": 3, "num_knn": 2, "distance_factor": 10.0, "field_radius": 1.0,
                                        "scale_mode": "unit_cube", "outside_value": 1.0}}}
    arrays = {"ijs": ijs.numpy(), "c2ws": c2w.numpy(), "near": near.numpy(), "far": far.numpy(),
              "field_ids": fid.numpy(), "positions": positions.numpy(), "orientations": orientations.numpy(),
              "gt_rgb": rgb.numpy(), "gt_depth": gt_depth.numpy(), "out_rgbds": pred.rgbds.numpy(),
              "meta_json": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for k, v in trained.items():
        arrays["param:" + k] = v.numpy()
    path = os.path.join(OUT, "trained_sphere.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
