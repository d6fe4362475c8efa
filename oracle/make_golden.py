"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Run in the build container only (needs ``/root/reference``):

    python -m oracle.make_golden

Each case stores every input (rays, poses, jitter, stacked field parameters, specs as
JSON) and every output of the reference call, so the CUDA path, the restatement and the
fixtures can be compared without the reference present (GPU box).  The reference has no
tests/golden vectors of its own, so these fixtures -- produced by
its own code on seeded inputs with injected sampling jitter -- are the pin.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CAM = dict(width=640, height=480, fx=554.2562584220408, fy=554.2562584220408, cx=319.5, cy=239.5,
           pixel_center=0.0)  # ngm/config/nrgbd_dataset.yaml:18-25


def _np(x):
    return None if x is None else x.detach().cpu().numpy()


def _save(name, meta, arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB")


def _rand_quat(g, n):
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


def _rand_c2w(g, shape=()):
    """Random rigid camera-to-world (OpenGL) transforms of the requested leading shape."""
    n = int(np.prod(shape)) if shape else 1
    q = _rand_quat(g, n)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(n, 3, 3)
    T = torch.eye(4).repeat(n, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = torch.randn(n, 3, generator=g) * 0.3
    return T.view(*shape, 4, 4) if shape else T[0]


def _make_map(ref, g, field_kwargs, num_fields, cfg_over=None, geometry_scale=0.12, geometry_bias=0.15):
    """NeuralGraphMap on CPU with ``num_fields`` independently seeded fields."""
    over = {"model_kwargs": {"field_kwargs": field_kwargs}}
    if cfg_over:
        for k, v in cfg_over.items():
            if k == "model_kwargs":
                over["model_kwargs"].update({kk: vv for kk, vv in v.items() if kk != "field_kwargs"})
            else:
                over[k] = v
    cfg = ref_loader.default_config(**over)
    m = ref.run_mapping.NeuralGraphMap(cfg)
    m._optimizer = None
    m._model.add_fields(num_fields)
    # independent seeded parameters per field (add_fields replicates the prototype)
    for k, v in m._model.all_fields_params.items():
        if v.dtype.is_floating_point and v.dim() > 1:
            scale = v.abs().max().clamp_min(1e-3)
            m._model.all_fields_params[k] = (torch.rand(v.shape, generator=g) * 2 - 1) * scale
    # calibrate the geometry channel so that occupancies are non-degenerate:
    # std(g) = geometry_scale over the unit cube, mean = geometry_bias
    last = f"_linears.{field_kwargs['num_layers']}"
    x = torch.rand(2048, 3, generator=g)
    with torch.no_grad():
        for f in range(num_fields):
            params = {k: v[f] for k, v in m._model.all_fields_params.items()}
            gch = torch.func.functional_call(m._model._prototype_field, (params, {}), x)[:, 3]
            s = geometry_scale / gch.std().clamp_min(1e-6)
            m._model.all_fields_params[last + ".weight"][f, 3] *= s
            b = m._model.all_fields_params[last + ".bias"]
            b[f, 3] = (b[f, 3] - gch.mean()) * s + geometry_bias
    pos = torch.randn(num_fields, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, -2.0])
    m._global_map_dict["positions"] = pos
    m._global_map_dict["orientations"] = _rand_quat(g, num_fields)
    m._global_map_dict["num"] = num_fields
    return m, cfg


def _params_np(m):
    return {"param:" + k: _np(v) for k, v in m._model.all_fields_params.items()}


NERF4 = {"encoding_type": "neural_graph_mapping.positional_encodings.PositionalEncodingNeRF",
         "encoding_kwargs": {"dim_in": 3, "num_octaves": 4},
         "num_layers": 2, "dim_out": 4, "dim_mlp_out": 32, "skip_mode": "no",
         "initial_geometry_bias": 0.0, "neus_initial_sd": 1.0}
NERF8_128 = dict(NERF4, encoding_kwargs={"dim_in": 3, "num_octaves": 8}, num_layers=4, dim_mlp_out=128)


def case_c1_vmap(ref):
    """BASELINE config 1: single field, 256 rays x 32 samples, 2-layer x 32 MLP."""
    g = torch.Generator().manual_seed(101)
    m, cfg = _make_map(ref, g, NERF4, 1, {"num_samples_coarse": 32, "num_samples_depth_guided": 0})
    m._global_map_dict["positions"] = torch.tensor([[0.0, 0.0, -2.0]])
    m._global_map_dict["orientations"] = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    F, R, S = 1, 256, 32
    ijs = torch.stack([torch.randint(0, 480, (F, R), generator=g),
                       torch.randint(0, 640, (F, R), generator=g)], -1)
    c2ws = torch.eye(4)
    near, far = torch.full((F, R), 1.0), torch.full((F, R), 3.0)
    jit = torch.rand(F, R, S, generator=g)
    fid = torch.tensor([0])
    cam = ref.camera.Camera(**CAM)
    with ref_loader.injected_jitter(jit), torch.no_grad():
        p = m._render_ijs(ijs, c2ws, cam, fid, True, near, far, None)
    _save("c1_vmap_256x32", {"case": "render_vmap", "camera": CAM, "field_kwargs": NERF4,
                             "config": {k: cfg[k] for k in RENDER_KEYS}, "num_samples": S},
          dict(ijs=_np(ijs), c2ws=_np(c2ws), near=_np(near), far=_np(far), jitter=_np(jit),
               field_ids=_np(fid), positions=_np(m._global_map_dict["positions"]),
               orientations=_np(m._global_map_dict["orientations"]),
               out_rgbds=_np(p.rgbds), out_color_vars=_np(p.color_vars),
               out_depth_vars=_np(p.depth_vars), out_term_probs=_np(p.term_probs), **_params_np(m)))


RENDER_KEYS = ["color_factor", "geometry_factor", "field_radius", "freespace_weight", "tsdf_weight",
               "near_distance", "far_distance", "geometry_mode", "truncation_distance",
               "num_samples_coarse", "num_samples_depth_guided", "range_depth_guided",
               "block_size", "pixel_block_size", "model_kwargs"]


def case_vmap_guided(ref, name, field_kwargs, F, R, Sc, Sg, mode="nrgbd", seed=202, F_all=None,
                     neg_near=False, color_factor=1.0, geometry_scale=0.12, geometry_bias=0.15):
    """Training-shape call: per-ray c2ws, near/far/gt, depth-guided merge, aux outputs."""
    g = torch.Generator().manual_seed(seed)
    F_all = F_all or F + 2
    m, cfg = _make_map(ref, g, field_kwargs, F_all,
                       {"num_samples_coarse": Sc, "num_samples_depth_guided": Sg,
                        "geometry_mode": mode, "color_factor": color_factor},
                       geometry_scale=geometry_scale, geometry_bias=geometry_bias)
    if mode == "neus":
        m._model.all_fields_params["_neus_sd"] = torch.rand(F_all, generator=g) * 0.5 + 0.2
    ijs = torch.stack([torch.randint(0, 480, (F, R), generator=g),
                       torch.randint(0, 640, (F, R), generator=g)], -1)
    c2ws = _rand_c2w(g, (F, R))
    near = torch.rand(F, R, generator=g) * 0.5 + 0.5
    if neg_near:
        near = near - 1.0
    far = near + 1.5 + torch.rand(F, R, generator=g)
    gt = near + (far - near) * (torch.rand(F, R, generator=g) * 1.4 - 0.2)  # some outside [near,far]
    gt[torch.rand(F, R, generator=g) < 0.15] = 0.0  # some unavailable
    if Sg == 0:
        gt_arg = None
    else:
        gt_arg = gt
    jit = torch.rand(F, R, Sc, generator=g)
    jit_g = torch.rand(F, R, Sg, generator=g) if Sg > 0 else None
    fid = torch.randperm(F_all, generator=g)[:F]
    cam = ref.camera.Camera(**CAM)
    jits = (jit, jit_g) if Sg > 0 else (jit,)
    with ref_loader.injected_jitter(*jits), torch.no_grad():
        p = m._render_ijs(ijs, c2ws, cam, fid, True, near.clone(), far.clone(),
                          None if gt_arg is None else gt_arg.clone())
    _save(name, {"case": "render_vmap", "camera": CAM, "field_kwargs": field_kwargs,
                 "config": {k: cfg[k] for k in RENDER_KEYS}, "num_samples": Sc,
                 "num_samples_depth_guided": Sg},
          dict(ijs=_np(ijs), c2ws=_np(c2ws), near=_np(near), far=_np(far),
               gt=_np(gt_arg) if gt_arg is not None else None,
               jitter=_np(jit), jitter_guided=_np(jit_g), field_ids=_np(fid),
               positions=_np(m._global_map_dict["positions"]),
               orientations=_np(m._global_map_dict["orientations"]),
               out_rgbds=_np(p.rgbds), out_color_vars=_np(p.color_vars),
               out_depth_vars=_np(p.depth_vars), out_term_probs=_np(p.term_probs),
               out_freespace=_np(p.freespace_geometry), out_tsdf=_np(p.tsdf_residuals),
               **_params_np(m)))


def case_knn(ref, name, field_kwargs, F_all, N, S, seed=303):
    """Eval-shape call (render_image's inner call): all fields, kNN K=2 blend, scalar near/far."""
    g = torch.Generator().manual_seed(seed)
    m, cfg = _make_map(ref, g, field_kwargs, F_all, {"eval_num_samples": S})
    # fields on a loose cluster in front of the camera so that rays cross several of them
    m._global_map_dict["positions"] = torch.randn(F_all, 3, generator=g) * 0.8 + torch.tensor([0., 0., -3.])
    m.eval()
    m._far_distance = 6.0
    ijs = torch.stack([torch.randint(0, 480, (N,), generator=g),
                       torch.randint(0, 640, (N,), generator=g)], -1)
    c2w = _rand_c2w(g)
    c2w[:3, 3] *= 0.3
    c2w[:3, :3] = torch.eye(3)
    jit = torch.rand(N, S, generator=g)
    cam = ref.camera.Camera(**CAM)
    with ref_loader.injected_jitter(jit), torch.no_grad():
        p = m._render_ijs(ijs, c2w, cam)
        # also pin the bare kNN field-set forward on the same sample points
        pts_cam, _ = cam.sample_ijs_uniform(ijs[:64], S, 0.0, 6.0) if False else (None, None)
    q = torch.randn(512, 3, generator=g) * 1.2 + torch.tensor([0., 0., -3.])
    with torch.no_grad():
        fs = m._model(q, m._global_map_dict["positions"][:F_all],
                      m._global_map_dict["orientations"][:F_all], None, False)
    _save(name, {"case": "render_knn", "camera": CAM, "field_kwargs": field_kwargs,
                 "config": {k: cfg[k] for k in RENDER_KEYS}, "num_samples": S,
                 "near_distance": 0.0, "far_distance": 6.0},
          dict(ijs=_np(ijs), c2ws=_np(c2w), jitter=_np(jit),
               positions=_np(m._global_map_dict["positions"]),
               orientations=_np(m._global_map_dict["orientations"]),
               out_rgbds=_np(p.rgbds), out_color_vars=_np(p.color_vars),
               out_depth_vars=_np(p.depth_vars), out_term_probs=_np(p.term_probs),
               knn_query=_np(q), knn_out=_np(fs), **_params_np(m)))


def case_quadrature(ref):
    g = torch.Generator().manual_seed(404)
    arrays, meta = {}, {"case": "quadrature", "modes": {}}
    lead, S = (3, 50), 24
    colors = torch.rand(*lead, S, 3, generator=g)
    geom = torch.randn(*lead, S, generator=g) * 0.3
    dist, _ = torch.sort(torch.rand(*lead, S, generator=g) * 4 + 0.2, dim=-1)
    depth = dist * (0.7 + 0.3 * torch.rand(*lead, 1, generator=g))
    isd = 1.0 / (torch.rand(lead[0], 1, 1, generator=g) + 0.3)
    arrays.update(colors=_np(colors), geom=_np(geom), dist=_np(dist), depth=_np(depth), isd=_np(isd))
    for mode, gf in [("nrgbd", 20.0), ("occupancy", 3.0), ("density", 5.0), ("neus", 10.0)]:
        m, _ = _make_map(ref, g, NERF4, 1, {"geometry_mode": mode, "geometry_factor": gf})
        with torch.no_grad():
            out = m._quadrature(colors, geom * (8.0 if mode == "density" else 1.0), dist, depth,
                                isd if mode == "neus" else None)
        meta["modes"][mode] = {"geometry_factor": gf, "geom_scale": 8.0 if mode == "density" else 1.0}
        for nm, o in zip(["colors", "depths", "color_vars", "depth_vars", "term", "weights"], out):
            arrays[f"out_{mode}_{nm}"] = _np(o)
    _save("quadrature_modes", meta, arrays)


def case_fields(ref):
    """NeuralField.forward for each in-tree encoding and skip mode (single field, local points)."""
    g = torch.Generator().manual_seed(505)
    pe = "neural_graph_mapping.positional_encodings."
    variants = {
        "nerf8_w128_l4": dict(NERF8_128),
        "nerf4_concat": dict(NERF4, skip_mode="concat", dim_mlp_out=40),
        "nerf4_add": dict(NERF4, skip_mode="add", dim_mlp_out=48),
        "nerf_start2": dict(NERF4, encoding_kwargs={"dim_in": 3, "num_octaves": 3, "start_octave": 2},
                            dim_mlp_out=None, num_layers=1),
        "fourier_raw": dict(NERF4, encoding_type=pe + "PositionalEncodingFourier",
                            encoding_kwargs={"dim_in": 3, "dim_out": 35, "mu": 0.0, "sigma": 4.0,
                                             "raw_coords": True}, dim_mlp_out=64),
        "fourier": dict(NERF4, encoding_type=pe + "PositionalEncodingFourier",
                        encoding_kwargs={"dim_in": 3, "dim_out": 32, "mu": 0.0, "sigma": 4.0,
                                         "raw_coords": False}, num_layers=3),
        "triplane_sum": dict(NERF4, encoding_type=pe + "TriplaneEncoding",
                             encoding_kwargs={"resolution": 16, "num_components": 16, "mode": "sum"}),
        "triplane_product": dict(NERF4, encoding_type=pe + "TriplaneEncoding",
                                 encoding_kwargs={"resolution": 16, "num_components": 16,
                                                  "init_scale": 0.7, "mode": "product"}),
        "triplane_concat": dict(NERF4, encoding_type=pe + "TriplaneEncoding",
                                encoding_kwargs={"resolution": 8, "num_components": 8, "mode": "concat"}),
    }
    arrays, meta = {}, {"case": "fields", "variants": {}}
    for name, fk in variants.items():
        torch.manual_seed(hash(name) % 1000)
        fld = ref.models.NeuralField(**fk)
        n = 300
        if "triplane" in name:
            x = torch.rand(n, 3, generator=g) * 2.4 - 1.2  # exercises border padding
        else:
            x = torch.rand(n, 3, generator=g)
        with torch.no_grad():
            y = fld(x)
            enc = fld._encoding(x)
        meta["variants"][name] = {"field_kwargs": fk}
        arrays[f"{name}:x"] = _np(x)
        arrays[f"{name}:y"] = _np(y)
        arrays[f"{name}:enc"] = _np(enc)
        for k, v in fld.state_dict().items():
            arrays[f"{name}:param:{k}"] = _np(v)
    _save("fields_forward", meta, arrays)


def case_sampler(ref):
    g = torch.Generator().manual_seed(606)
    cam = ref.camera.Camera(**CAM)
    ijs = torch.stack([torch.randint(0, 480, (2, 40), generator=g),
                       torch.randint(0, 640, (2, 40), generator=g)], -1)
    near = torch.rand(2, 40, generator=g)
    far = near + 2 * torch.rand(2, 40, generator=g) + 0.1
    jit = torch.rand(2, 40, 12, generator=g)
    with ref_loader.injected_jitter(jit):
        pts, dist = cam.sample_ijs_uniform(ijs, 12, near, far)
    jit2 = torch.rand(2, 40, 7, generator=g)
    with ref_loader.injected_jitter(jit2):
        pts2, dist2 = cam.sample_ijs_uniform(ijs, 7, 0.25, 5.0)
    cam2 = ref.camera.Camera(320, 240, 300.0, 310.0, 158.7, 121.2, pixel_center=0.5)
    dirs2 = cam2.ijs_to_directions(ijs // 2)
    c2w = _rand_c2w(g, (2, 40))
    world = ref.utils.transform_points(pts, c2w.unsqueeze(-3))
    _save("sampler", {"case": "sampler", "camera": CAM,
                      "camera2": dict(width=320, height=240, fx=300.0, fy=310.0, cx=158.7, cy=121.2,
                                      pixel_center=0.5)},
          dict(ijs=_np(ijs), near=_np(near), far=_np(far), jitter=_np(jit), out_points=_np(pts),
               out_dist=_np(dist), jitter2=_np(jit2), out_points2=_np(pts2), out_dist2=_np(dist2),
               out_dirs_cam2=_np(dirs2), c2ws=_np(c2w), out_world=_np(world)))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    torch.set_num_threads(8)
    case_c1_vmap(ref)
    case_vmap_guided(ref, "vmap_guided_nrgbd", NERF4, F=3, R=48, Sc=8, Sg=16)
    case_vmap_guided(ref, "vmap_behind_camera", NERF4, F=2, R=40, Sc=12, Sg=0, neg_near=True, seed=212,
                     color_factor=0.5)
    case_vmap_guided(ref, "vmap_neus", dict(NERF4), F=2, R=32, Sc=16, Sg=0, mode="neus", seed=222)
    case_vmap_guided(ref, "vmap_density", NERF4, F=2, R=32, Sc=16, Sg=8, mode="density", seed=232)
    case_vmap_guided(ref, "vmap_occupancy", NERF4, F=2, R=32, Sc=16, Sg=0, mode="occupancy", seed=242,
                     geometry_scale=0.5, geometry_bias=-1.2)
    case_vmap_guided(ref, "c2_vmap_w128_s64", NERF8_128, F=2, R=48, Sc=64, Sg=0, seed=252,
                     geometry_scale=0.08, geometry_bias=0.3)
    case_knn(ref, "knn_render", NERF4, F_all=6, N=192, S=24)
    case_knn(ref, "knn_render_w128", NERF8_128, F_all=4, N=64, S=64, seed=313)
    case_quadrature(ref)
    case_fields(ref)
    case_sampler(ref)


if __name__ == "__main__":
    main()
