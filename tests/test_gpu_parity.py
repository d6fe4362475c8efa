"""GPU parity tests: the CUDA path (through the C ABI / the Python mirror) against the golden
vectors of the unmodified reference and against the oracle restatement on seeded inputs.

Tolerances (fp32 path; stated per north_star "within a stated fp32 tolerance"):
  colour / variances / term prob : 2e-5 abs + 2e-5 rel  (sum of <=128 products of O(1) terms)
  depth                          : 1e-4 m abs
The reference's own CPU-vs-CUDA difference is of the same order (different sgemm order)."""
import pytest
import torch

import golden_util as G
from oracle import restatement as R
from tests_support import make_state, product_field_kwargs, run_vmap_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, atol=2e-5, rtol=2e-5, what=""):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    ok = err <= atol + rtol * b.abs()
    assert bool(ok.all()), f"{what}: max abs err {err.max().item():.3e}, {int((~ok).sum())} of {ok.numel()} off"


# ---------------------------------------------------------------- sampler (a1-a4)
def test_sampler_golden():
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200.camera import sample_rays

    meta, a = G.load("sampler")
    cam = ngm.Camera(**meta["camera"])
    ijs = a["ijs"].to(DEV)
    pts, dist, world, depth = sample_rays(cam, ijs, 12, a["near"].to(DEV), a["far"].to(DEV), c2ws=a["c2ws"].to(DEV),
                                          jitter=a["jitter"].to(DEV), want_world=True, want_depth=True)
    _close(pts, a["out_points"], 2e-6, 2e-6, "points")
    _close(dist, a["out_dist"], 2e-6, 2e-6, "dist")
    _close(world, a["out_world"], 3e-6, 3e-6, "world")
    _close(depth, -a["out_points"][..., 2], 2e-6, 2e-6, "depth")
    pts2, dist2 = sample_rays(cam, ijs, 7, 0.25, 5.0, jitter=a["jitter2"].to(DEV))
    _close(pts2, a["out_points2"], 2e-6, 2e-6)
    _close(dist2, a["out_dist2"], 2e-6, 2e-6)
    cam2 = ngm.Camera(**meta["camera2"])
    _close(cam2.ijs_to_directions((a["ijs"] // 2).to(DEV)), a["out_dirs_cam2"], 1e-6, 1e-6, "dirs")


@pytest.mark.parametrize("S,Sg", [(8, 16), (1, 1), (64, 0), (33, 31), (200, 57)])
def test_sampler_guided_merge_vs_oracle(S, Sg):
    """Depth-guided merge: rank computation == the reference's cat + sort + gather."""
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200.camera import sample_rays

    g = torch.Generator().manual_seed(S * 131 + Sg)
    F, Rr = 3, 257
    cam_spec = R.CameraSpec()
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5
    far = near + 0.5 + torch.rand(F, Rr, generator=g) * 2
    gt = near + (far - near) * (torch.rand(F, Rr, generator=g) * 1.4 - 0.2)
    gt[torch.rand(F, Rr, generator=g) < 0.2] = 0.0
    j1, j2 = torch.rand(F, Rr, S, generator=g), torch.rand(F, Rr, max(Sg, 1), generator=g)[..., :Sg]
    # oracle (reference lines run_mapping.py:521-545)
    pc, d = R.sample_ijs_uniform(ijs, cam_spec, S, near, far, j1)
    if Sg:
        m = (gt == 0.0) + (near > gt) + (far < gt)
        gn, gf = gt - 0.1, gt + 0.1
        gn[m], gf[m] = near[m], far[m]
        pg, dg = R.sample_ijs_uniform(ijs, cam_spec, Sg, gn, gf, j2)
        pc, d = torch.cat([pc, pg], -2), torch.cat([d, dg], -1)
        d, idx = torch.sort(d, dim=-1)
        pc = torch.gather(pc, -2, idx.unsqueeze(-1).expand(*idx.shape, 3))
    cam = ngm.Camera(cam_spec.width, cam_spec.height, cam_spec.fx, cam_spec.fy, cam_spec.cx, cam_spec.cy)
    pts, dist = sample_rays(cam, ijs.to(DEV), S, near.to(DEV), far.to(DEV), gt=gt.to(DEV) if Sg else None,
                            num_samples_guided=Sg, range_guided=0.1, jitter=j1.to(DEV),
                            jitter_guided=j2.to(DEV) if Sg else None)
    assert bool((dist[..., 1:] >= dist[..., :-1]).all()), "distances must come out sorted"
    _close(dist, d, 2e-6, 2e-6, "dist")
    _close(pts, pc, 2e-6, 2e-6, "points")


def test_sampler_philox_statistics():
    """Without an injected jitter tensor the kernel draws Philox noise: stratified, in-range, uniform."""
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200.camera import sample_rays

    cam = ngm.Camera(640, 480, 554.25, 554.25, 319.5, 239.5)
    ijs = torch.cartesian_prod(torch.arange(0, 480, 7), torch.arange(0, 640, 9)).to(DEV)
    S = 64
    _, dist = sample_rays(cam, ijs, S, 1.0, 3.0, seed=1234)
    u = (dist - 1.0) / (2.0 / S) - torch.arange(S, device=DEV)
    assert float(u.min()) >= -1e-4 and float(u.max()) < 1.0 + 1e-4
    assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1 / 12) < 5e-3
    _, dist_b = sample_rays(cam, ijs, S, 1.0, 3.0, seed=1234)
    assert torch.equal(dist, dist_b)
    _, dist_c = sample_rays(cam, ijs, S, 1.0, 3.0, seed=99)
    assert not torch.equal(dist, dist_c)


# ---------------------------------------------------------------- compositor (a12-a13)
def test_quadrature_modes_golden():
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("quadrature_modes")
    for mode, info in meta["modes"].items():
        st = type("D", (), {"_geometry_mode": mode, "_geometry_factor": info["geometry_factor"]})()
        out = ngm.quadrature(st, a["colors"].to(DEV), (a["geom"] * info["geom_scale"]).to(DEV), a["dist"].to(DEV),
                             a["depth"].to(DEV), a["isd"].to(DEV) if mode == "neus" else None)
        for nm, o in zip(["colors", "depths", "color_vars", "depth_vars", "term", "weights"], out):
            _close(o, a[f"out_{mode}_{nm}"], 5e-6, 2e-5, f"{mode}/{nm}")


@pytest.mark.parametrize("S", [1, 31, 64, 129, 640])
def test_quadrature_sizes_vs_oracle(S):
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(S)
    N = 300
    colors, geom = torch.rand(N, S, 3, generator=g), torch.randn(N, S, generator=g) * 0.2
    dist, _ = torch.sort(torch.rand(N, S, generator=g) * 4, dim=-1)
    depth = dist * 0.9
    for mode, gf in [("nrgbd", 20.0), ("occupancy", 2.0)] + ([("density", 4.0)] if S > 1 else []):
        ref = R.quadrature(colors, geom, dist, depth, None, mode, gf)
        st = type("D", (), {"_geometry_mode": mode, "_geometry_factor": gf})()
        out = ngm.quadrature(st, colors.to(DEV), geom.to(DEV), dist.to(DEV), depth.to(DEV), None)
        for nm, o, r in zip(["colors", "depths", "color_vars", "depth_vars", "term", "weights"], out, ref):
            _close(o, r, 1e-5, 5e-5, f"S={S} {mode}/{nm}")


@pytest.mark.parametrize("S", [4, 8, 12, 64, 100])
@pytest.mark.parametrize("overwrite", [False, True])
def test_composite_staged_kernel(S, overwrite, monkeypatch):
    """Packed (N,S,4) MLP output takes the thread-per-ray staged kernel: against the oracle's _quadrature
    (behind-camera overwrite applied on the host, run_mapping.py:614-622) and against the general kernel."""
    from neural_graph_mapping_b200 import renderer

    g = torch.Generator().manual_seed(100 + S)
    N = 1000 + S  # not a multiple of 32: partial last tile
    packed = torch.cat([torch.rand(N, S, 3, generator=g), torch.randn(N, S, 1, generator=g) * 0.2], -1)
    dist, _ = torch.sort(torch.rand(N, S, generator=g) * 4, dim=-1)
    depth = dist * 0.9
    if overwrite:
        depth = depth - 0.5  # some samples behind the camera
    isd = torch.rand(4, generator=g) + 0.5
    pk, dd, zz = packed.to(DEV), dist.to(DEV), depth.to(DEV)
    for mode, gf in [("nrgbd", 20.0), ("occupancy", 2.0), ("density", 4.0), ("neus", 3.0)]:
        geom = packed[..., 3].clone()
        if overwrite:
            geom[depth < 0] = -100.0 if mode in ("occupancy", "density") else 1.0
        rays_per = (N + 3) // 4
        isds = isd.repeat_interleave(rays_per)[:N, None] if mode == "neus" else None
        ref = R.quadrature(packed[..., :3], geom, dist, depth, isds, mode, gf)
        kw = dict(color_stride=4, geometry_stride=4, overwrite_behind_camera=overwrite,
                  neus_isd=isd.to(DEV) if mode == "neus" else None, rays_per_isd=rays_per)
        out = renderer.composite(pk, pk[..., 3], dd, zz, mode, gf, **kw)
        monkeypatch.setenv("NGM_COMPOSITE_STAGED", "0")
        gen = renderer.composite(pk, pk[..., 3], dd, zz, mode, gf, **kw)
        monkeypatch.delenv("NGM_COMPOSITE_STAGED")
        got = (out[0][:, :3], out[0][:, 3], out[1], out[2], out[3])
        old = (gen[0][:, :3], gen[0][:, 3], gen[1], gen[2], gen[3])
        for nm, o, r, q in zip(["colors", "depths", "color_vars", "depth_vars", "term"], got, ref, old):
            _close(o, r, 1e-5, 5e-5, f"staged S={S} {mode}/{nm}")
            _close(o, q.cpu(), 1e-5, 5e-5, f"staged vs general S={S} {mode}/{nm}")


# ---------------------------------------------------------------- field evaluation (a6, a8-a11, a16)
def test_fields_forward_golden():
    """NeuralField.forward for every in-tree encoding and skip mode vs the reference's outputs."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("fields_forward")
    for name, info in meta["variants"].items():
        fk = product_field_kwargs(info["field_kwargs"])
        fld = ngm.NeuralField(**fk).to(DEV)
        sd = {k: v.to(DEV) for k, v in G.params(a, prefix=f"{name}:param:").items()}
        missing = fld.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, (name, missing)
        with torch.no_grad():
            y = fld(a[f"{name}:x"].to(DEV))
        scale = a[f"{name}:y"].abs().max().item()
        _close(y, a[f"{name}:y"], 3e-5 * max(scale, 1.0), 3e-5, name)


def test_permuto_field_vs_oracle():
    """Permutohedral encoding (parity UNPINNED vs the third-party original): CUDA vs oracle/permuto.py."""
    import neural_graph_mapping_b200 as ngm

    kw = dict(pos_dim=3, log2_hashmap_size=12, nr_levels=16, nr_feat_per_level=2, coarsest_scale=1.0,
              finest_scale=1e-4, init_scale=0.5)
    for concat in (False, True):
        torch.manual_seed(5)
        ekw = dict(kw, concat_points=concat, concat_points_scaling=0.5)
        fld = ngm.NeuralField("neural_graph_mapping_b200.positional_encodings.PermutohedralEncoding", ekw,
                              num_layers=1, dim_out=4, dim_mlp_out=None).to(DEV)
        x = torch.rand(1000, 3)
        spec = R.FieldSpec("permuto", ekw, 1, 4, None, "no")
        params = {k: v.detach().cpu() for k, v in fld.state_dict().items()}
        ref = R.field_forward(x, spec, params)
        with torch.no_grad():
            y = fld(x.to(DEV))
        assert y.shape == (1000, 4)
        _close(y, ref, 2e-5, 2e-5, f"permuto concat={concat}")


def test_fieldset_vmap_vs_oracle():
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    st = make_state(meta, a, DEV)
    g = torch.Generator().manual_seed(11)
    ids = torch.tensor([4, 0, 2])
    q = torch.randn(3, 333, 3, generator=g) * 0.5 + a["positions"][ids][:, None]
    st._model.set_vmap_fields(ids.to(DEV))
    with torch.no_grad():
        y = st._model(q.to(DEV), a["positions"][ids].to(DEV), a["orientations"][ids].to(DEV), ids.to(DEV), True)
    fs, rs = G.field_spec(meta["field_kwargs"]), G.render_spec(meta)
    ref = R.fieldset_forward_vmap(q, a["positions"][ids], a["orientations"][ids], fs,
                                  {k: v[ids] for k, v in G.params(a).items()}, rs)
    _close(y, ref, 2e-5, 2e-5, "fieldset vmap")


# ---------------------------------------------------------------- fused render (a14)
@pytest.mark.parametrize("name", G.VMAP_CASES)
def test_render_vmap_golden(name):
    meta, a = G.load(name)
    p = run_vmap_case(meta, a, DEV, precision="fp32")
    _close(p.rgbds[..., :3], a["out_rgbds"][..., :3], 3e-5, 3e-5, "colour")
    _close(p.rgbds[..., 3], a["out_rgbds"][..., 3], 1e-4, 3e-5, "depth")
    _close(p.color_vars, a["out_color_vars"], 3e-5, 3e-5, "colour var")
    _close(p.depth_vars, a["out_depth_vars"], 1e-4, 3e-5, "depth var")
    _close(p.term_probs, a["out_term_probs"], 3e-5, 3e-5, "term")
    if "out_freespace" in a:
        assert p.freespace_geometry.shape == a["out_freespace"].shape
        _close(p.freespace_geometry, a["out_freespace"], 3e-5, 3e-5, "freespace")
        _close(p.tsdf_residuals, a["out_tsdf"], 3e-5, 3e-5, "tsdf")
    else:
        assert p.freespace_geometry is None and p.tsdf_residuals is None


def test_render_errors_match_reference():
    meta, a = G.load("c1_vmap_256x32")
    st = make_state(meta, a, DEV)
    import neural_graph_mapping_b200 as ngm

    cam = ngm.Camera(**meta["camera"])
    with pytest.raises(ValueError, match="field_ids=None only supported"):
        st._render_ijs(a["ijs"].to(DEV), a["c2ws"].to(DEV), cam, None, True)
    with pytest.raises(RuntimeError, match="no CPU"):
        st._render_ijs(a["ijs"], a["c2ws"], cam, a["field_ids"], True)


# ---------------------------------------------------------------- kNN path (a7, a15)
@pytest.mark.parametrize("name", G.KNN_CASES)
def test_render_knn_golden(name):
    """_render_ijs(use_vmap=False): all fields, K=2 softmax blend, outside fill (models.py:347-405)."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load(name)
    st = make_state(meta, a, DEV)
    st.eval()
    cam = ngm.Camera(**meta["camera"])
    with torch.no_grad():
        p = st._render_ijs(a["ijs"].to(DEV), a["c2ws"].to(DEV), cam, jitter=a["jitter"].to(DEV))
        out = st._model(a["knn_query"].to(DEV), a["positions"].to(DEV), a["orientations"].to(DEV), None, False)
    _close(out, a["knn_out"], 3e-5, 3e-5, "fieldset kNN")
    _close(p.rgbds[..., :3], a["out_rgbds"][..., :3], 3e-5, 3e-5, "colour")
    _close(p.rgbds[..., 3], a["out_rgbds"][..., 3], 1e-4, 3e-5, "depth")
    _close(p.color_vars, a["out_color_vars"], 3e-5, 3e-5, "colour var")
    _close(p.depth_vars, a["out_depth_vars"], 1e-4, 3e-5, "depth var")
    _close(p.term_probs, a["out_term_probs"], 3e-5, 3e-5, "term")


@pytest.mark.parametrize("side,K", [(12, 2), (45, 2), (45, 3)])
def test_fieldset_knn_large_map_rays(side, K):
    """A map of side x side fields on a grid and points along rays (neighbouring samples are spatially coherent, the
    renderer's access pattern): the candidate pruning of the assign kernel -- per warp, and per block on maps of more
    than a few hundred fields -- must select exactly the K nearest fields of the full scan."""
    import neural_graph_mapping_b200 as ngm

    F = side * side
    g = torch.Generator().manual_seed(side + K)
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 4}, 1, 4, 16, "no")
    n_tab = 8
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(n_tab)])
    idx = torch.arange(F)
    pos = torch.stack([(idx % side).float() * 1.1547, torch.zeros(F), -(idx // side).float() * 1.1547], -1)
    pos = pos + 0.05 * torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    fid = torch.randint(0, n_tab, (F,), generator=g)
    rays, S = 70, 64  # 4,480 points: the last block of the kernel is partial
    origin = torch.stack([torch.rand(rays, generator=g) * side * 1.1547, torch.rand(rays, generator=g) * 0.6 - 0.3,
                          -torch.rand(rays, generator=g) * side * 1.1547], -1)
    d = torch.randn(rays, 3, generator=g) * torch.tensor([1.0, 0.15, 1.0])
    d = d / d.norm(dim=-1, keepdim=True)
    t = torch.linspace(0.0, 4.0, S)
    pts = (origin[:, None] + t[None, :, None] * d[:, None]).reshape(-1, 3)
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube", num_knn=K, distance_factor=10.0, outside_value=1.0)
    ref = R.fieldset_forward_knn(pts, pos, ori, fid, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                "encoding_kwargs": {"dim_in": 3, "num_octaves": 4}, "num_layers": 1, "dim_out": 4,
                                "dim_mlp_out": 16}, K, 10.0, 1.0, field_radius=1.0, scale_mode="unit_cube").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), fid.to(DEV), False)
    dd = torch.cdist(pts, pos)
    srt = torch.sort(dd, dim=-1)[0]
    margin = (srt[:, 0] - 1.0).abs() > 1e-5
    for j in range(K):
        margin &= (srt[:, j + 1] - srt[:, j]).abs() > 1e-5
    assert margin.float().mean() > 0.95
    _close(y.cpu()[margin], ref[margin], 3e-5, 3e-5, "kNN fieldset, large map")
    inside = (ref != 1.0).any(-1).float().mean().item()
    assert 0.3 < inside <= 1.0


@pytest.mark.parametrize("F,K,n", [(1, 2, 300), (3, 1, 1000), (40, 2, 5000), (1500, 3, 4000), (9000, 2, 3000)])
def test_fieldset_knn_vs_oracle(F, K, n):
    """Field counts below K, K=1/3, more centres than one shared-memory chunk, more fields than the per-block
    histogram of the bucketing kernels holds (global counters then), field_ids remap."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(F * 10 + K)
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 4}, 1, 4, 16, "no")
    n_tab = min(F, 8)  # a small parameter table; field i uses row field_ids[i]
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(n_tab)])
    pos = torch.randn(F, 3, generator=g) * (1.0 if F < 100 else 6.0)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    fid = torch.randint(0, n_tab, (F,), generator=g)
    pts = torch.randn(n, 3, generator=g) * (1.5 if F < 100 else 6.0)
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube", num_knn=K, distance_factor=10.0, outside_value=1.0)
    ref = R.fieldset_forward_knn(pts, pos, ori, fid, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                "encoding_kwargs": {"dim_in": 3, "num_octaves": 4}, "num_layers": 1, "dim_out": 4,
                                "dim_mlp_out": 16}, K, 10.0, 1.0, field_radius=1.0, scale_mode="unit_cube").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), fid.to(DEV), False)
    # points whose two nearest centres are (nearly) equidistant may legitimately pick another neighbour order
    d = torch.cdist(pts, pos)
    srt = torch.sort(d, dim=-1)[0]
    margin = torch.ones(n, dtype=torch.bool)
    for j in range(min(K, F - 1)):
        margin &= (srt[:, j + 1] - srt[:, j]).abs() > 1e-5
    margin &= (srt[:, 0] - 1.0).abs() > 1e-5
    _close(y.cpu()[margin], ref[margin], 3e-5, 3e-5, "kNN fieldset")
    inside = (ref != 1.0).any(-1).float().mean().item()
    assert 0.0 < inside <= 1.0


def test_render_image_knn():
    """render_image = pixel grid in pixel_block_size chunks through _render_ijs (run_mapping.py:402-437)."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("knn_render")
    meta = dict(meta)
    meta["config"] = dict(meta["config"], pixel_block_size=1000)
    st = make_state(meta, a, DEV)
    st.eval()
    cam = ngm.Camera(64, 48, 55.4, 55.4, 31.5, 23.5)
    torch.manual_seed(0)
    rgbds, dvars = st.render_image(a["c2ws"].to(DEV), cam)
    assert rgbds.shape == (48, 64, 4) and dvars.shape == (48, 64)
    assert torch.isfinite(rgbds).all() and torch.isfinite(dvars).all()
    # same render with the noise injected: oracle parity on the full image
    g = torch.Generator().manual_seed(9)
    jit = torch.rand(48 * 64, meta["num_samples"], generator=g)
    ijs = torch.cartesian_prod(torch.arange(48), torch.arange(64))
    with torch.no_grad():
        p = st._render_ijs(ijs.to(DEV), a["c2ws"].to(DEV), cam, jitter=jit.to(DEV))
    fs, rs = G.field_spec(meta["field_kwargs"]), G.render_spec(meta)
    cs = R.CameraSpec(64, 48, 55.4, 55.4, 31.5, 23.5)
    ref = R.render_rays(ijs, a["c2ws"], cs, rs, fs, G.params(a), a["positions"], a["orientations"], jitter=jit)
    _close(p.rgbds, ref.rgbds, 1e-4, 5e-5, "image rgbd")
    psnr = R.psnr(p.rgbds[..., :3].cpu().reshape(48, 64, 3), ref.rgbds[..., :3].reshape(48, 64, 3))
    assert psnr > 80.0, psnr
