"""Round-2 parity hardening: skip mode `rezero` on the GPU, the reference's gate of the
behind-camera overwrite (ngm/run_mapping.py:494-495), output-buffer alignment errors, in-kernel jitter offsets."""
import pytest
import torch

import golden_util as G
from oracle import restatement as R
from tests_support import make_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


@pytest.mark.parametrize("skip", ["rezero", "add", "concat"])
def test_skip_modes_random_params_vs_oracle(skip):
    """models.py:160-180 with NON-zero rezero gains (the reference initialises them to 0, which would hide a wrong
    residual path) on the fp32 kernel, against the oracle restatement."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(7)
    F, n, W, L, O = 3, 515, 48, 3, 4
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": O}, L, 4, W, skip)
    per_field = [R.init_field_params(spec, g) for _ in range(F)]
    if skip == "rezero":
        for p in per_field:
            p["_rezero"] = torch.rand(L, generator=g) + 0.25
    params = R.stack_params(per_field)
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    pts = pos[:, None] + torch.rand(F, n, 3, generator=g) * 1.6 - 0.8
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube")
    ref = R.fieldset_forward_vmap(pts, pos, ori, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                "encoding_kwargs": {"dim_in": 3, "num_octaves": O}, "num_layers": L, "dim_out": 4,
                                "dim_mlp_out": W, "skip_mode": skip}, 2, 10.0, 1.0, field_radius=1.0,
                               scale_mode="unit_cube", precision="fp32").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    model.set_vmap_fields(None)
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), None, True)
    err = (y.cpu() - ref).abs().max().item()
    assert err < 5e-5 * max(1.0, ref.abs().max().item()), f"{skip}: max |cuda - oracle| = {err:.3e}"


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("some_near_negative", [False, True])
def test_behind_camera_overwrite_gate(prec, some_near_negative):
    """run_mapping.py:494-495: with every near >= 0 the reference keeps the network output even for depth-guided
    samples that fall behind the camera (gt < range_depth_guided); with one negative near, every sample behind the
    camera is overwritten.  Both against the oracle; the two cases must differ (negative control)."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    S, Sg = 8, 8
    meta = dict(meta, num_samples=S, num_samples_depth_guided=Sg)
    g = torch.Generator().manual_seed(21)
    F, Rr = 2, 48
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.zeros(F, Rr)
    far = torch.full((F, Rr), 1.5)
    gt = torch.rand(F, Rr, generator=g) * 0.08 + 0.005   # < range_depth_guided (0.1): guided window starts below 0
    if some_near_negative:
        near[0, 0] = -0.05
    jit = torch.rand(F, Rr, S, generator=g)
    jg = torch.rand(F, Rr, Sg, generator=g)
    fid = torch.tensor([1, 3])
    c2ws = a["c2ws"][:F, :1].expand(F, Rr, 4, 4).contiguous()
    cam = ngm.Camera(**meta["camera"])
    st = make_state(meta, a, DEV, prec)
    with torch.no_grad():
        p = st._render_ijs(ijs.to(DEV), c2ws.to(DEV), cam, fid.to(DEV), True, near.to(DEV), far.to(DEV), gt.to(DEV),
                           jitter=jit.to(DEV), jitter_guided=jg.to(DEV))
    fs, rs, cs = G.field_spec(meta["field_kwargs"]), G.render_spec(meta), G.camera_spec(meta["camera"])

    def oracle(near_):
        return R.render_rays(ijs, c2ws, cs, rs, fs, G.params(a), a["positions"], a["orientations"], field_ids=fid,
                             use_vmap=True, near_distances=near_, far_distances=far, gt_distances=gt.clone(),
                             jitter=jit, jitter_guided=jg)

    ref = oracle(near)
    tol_c, tol_d = (5e-5, 2e-4) if prec == "fp32" else (2e-2, 5e-2)
    assert (p.rgbds[..., :3].cpu() - ref.rgbds[..., :3]).abs().max().item() < tol_c
    assert (p.rgbds[..., 3].cpu() - ref.rgbds[..., 3]).abs().max().item() < tol_d
    assert p.tsdf_residuals.shape == ref.tsdf_residuals.shape
    assert (p.tsdf_residuals.cpu() - ref.tsdf_residuals).abs().max().item() < (1e-4 if prec == "fp32" else 2e-2)
    # negative control: the gate matters on this batch (samples behind the camera exist)
    other_near = near.clone()
    other_near[0, 0] = -0.05 if not some_near_negative else 0.0
    other = oracle(other_near)
    assert (other.rgbds[1] - ref.rgbds[1]).abs().max().item() > 1e-3, "fixture has no sample behind the camera"


def test_render_out_alignment_is_a_value_error():
    """A misaligned rgbd output buffer is a clean ValueError on both precisions (ngm_render_rays_fwd checks it)."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=8, num_samples_depth_guided=0)
    F, Rr = 1, 5
    ijs = torch.zeros(F, Rr, 2, dtype=torch.long, device=DEV)
    cam = ngm.Camera(**meta["camera"])
    for prec in ("fp32", "fp16"):
        st = make_state(meta, a, DEV, prec)
        flat = torch.empty(1 + 9 * F * Rr, device=DEV)
        base = flat[1:]  # 4-byte offset: not 16-byte aligned
        outs = (base[:4 * F * Rr].view(F, Rr, 4), base[4 * F * Rr:7 * F * Rr].view(F, Rr, 3),
                base[7 * F * Rr:8 * F * Rr].view(F, Rr), base[8 * F * Rr:9 * F * Rr].view(F, Rr))
        with pytest.raises(ValueError, match="16-byte aligned"), torch.no_grad():
            st._render_ijs(ijs, a["c2ws"][0, 0].to(DEV), cam, torch.tensor([0], device=DEV), True,
                           torch.full((F, Rr), 0.5, device=DEV), torch.full((F, Rr), 2.0, device=DEV), out=outs)


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_shard_with_seed_and_offset_equals_whole(prec):
    """Field shards rendered with the batch's seed and their global sample offset draw the in-kernel jitter of the
    whole batch: the single-process form of `distributed.render_rays_sharded`."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    S = 16
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0)
    g = torch.Generator().manual_seed(5)
    F, Rr = 4, 96
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1).to(DEV)
    near = (torch.rand(F, Rr, generator=g) * 0.5 + 0.3).to(DEV)
    far = near + 1.5
    fid = torch.tensor([0, 3, 1, 4], device=DEV)
    cam = ngm.Camera(**meta["camera"])
    st = make_state(meta, a, DEV, prec)
    c2w = a["c2ws"][0, 0].to(DEV)
    with torch.no_grad():
        whole = st._render_ijs(ijs, c2w, cam, fid, True, near, far, seed=1234)
        again = st._render_ijs(ijs, c2w, cam, fid, True, near, far, seed=1234)
        other = st._render_ijs(ijs, c2w, cam, fid, True, near, far, seed=1235)
        parts = [st._render_ijs(ijs[f0:f0 + 2], c2w, cam, fid[f0:f0 + 2], True, near[f0:f0 + 2], far[f0:f0 + 2],
                                seed=1234, sample_offset=f0 * Rr * S) for f0 in (0, 2)]
    assert torch.equal(whole.rgbds, again.rgbds)
    assert not torch.equal(whole.rgbds, other.rgbds)
    assert torch.equal(torch.cat([p.rgbds for p in parts]), whole.rgbds)
    assert torch.equal(torch.cat([p.depth_vars for p in parts]), whole.depth_vars)
