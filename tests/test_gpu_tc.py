"""GPU tests of the fp16-operand tensor-core (tcgen05) path.

Tolerances (fp16 operands, fp32 accumulation; north_star: "PSNR within 0.1 dB"):
  per-pixel colour L1 (mean) <= 2e-3, depth L1 (mean) <= 5e-3 m against the fp32 reference
  golden vectors; the raw MLP output within 2% of its scale (max) / 0.3% (mean).
"""

import pytest
import torch

import golden_util as G
from oracle import restatement as R
from tests_support import make_state, product_field_kwargs, run_vmap_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _gemm(n, k, rows, seed):
    from neural_graph_mapping_b200 import _lib

    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(rows, k, generator=g) * 0.5).half()
    w = torch.randn(n, k, generator=g) * 0.3
    b = torch.randn(n, generator=g)
    ad, wd, bd = a.to(DEV), w.to(DEV), b.to(DEV)
    out = torch.full((rows, n), float("nan"), device=DEV)
    ws = torch.zeros(256 * 1024, dtype=torch.uint8, device=DEV)
    rc = _lib.load_debug_lib().ngm_debug_tc_gemm(wd.data_ptr(), bd.data_ptr(), n, k, ad.data_ptr(), rows, out.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _lib.stream_ptr(torch.device(DEV)))
    _lib.check(rc)
    torch.cuda.synchronize()
    ref = a.float() @ w.half().float().T + b
    return out.cpu(), ref


@pytest.mark.parametrize("n,k,rows", [(128, 128, 128), (16, 128, 300), (128, 48, 257), (64, 64, 1000), (4, 32, 77),
                                      (128, 16, 128), (100, 112, 129)])
def test_tc_gemm_plumbing(n, k, rows):
    """Weight packing + SWIZZLE_128B smem descriptors + A operand in TMEM + tcgen05.mma + TMEM epilogue."""
    out, ref = _gemm(n, k, rows, seed=n * 1000 + k)
    err = (out - ref).abs()
    assert torch.isfinite(out).all(), "non-finite output (uninitialised rows?)"
    assert err.max().item() < 5e-3, f"n={n} k={k}: max err {err.max().item():.3e}; " \
        f"first bad {torch.nonzero(err > 5e-3)[:5].tolist()}"


def test_tc_gemm_one_hot_layout():
    """One-hot operands pin the exact (row, k) -> TMEM and (n, k) -> smem mappings."""
    from neural_graph_mapping_b200 import _lib

    n, k, rows = 128, 128, 128
    a = torch.zeros(rows, k, dtype=torch.half)
    a[torch.arange(rows), torch.arange(rows) % k] = 1.0  # row r selects k = r
    b = torch.zeros(n)
    for which, w in (("n", torch.arange(n)[:, None].expand(n, k).float()),
                     ("k", torch.arange(k)[None, :].expand(n, k).float())):  # exact in fp16
        ad, wd, bd = a.to(DEV), w.contiguous().to(DEV), b.to(DEV)
        out = torch.empty(rows, n, device=DEV)
        ws = torch.zeros(256 * 1024, dtype=torch.uint8, device=DEV)
        _lib.check(_lib.load_debug_lib().ngm_debug_tc_gemm(wd.data_ptr(), bd.data_ptr(), n, k, ad.data_ptr(), rows, out.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _lib.stream_ptr(torch.device(DEV))))
        ref = w.T[torch.arange(rows) % k]  # out[r, n] = W[n, r]
        assert torch.equal(out.cpu(), ref), \
            f"W[n,k]={which}: layout mismatch, out[1,:6]={out[1, :6].tolist()} out[5,:6]={out[5, :6].tolist()}"


def _field(fk, precision):
    import neural_graph_mapping_b200 as ngm

    return ngm.NeuralField(**product_field_kwargs(fk), precision=precision).to(DEV)


def test_field_fwd_fp16_vs_fp32_and_reference():
    meta, a = G.load("fields_forward")
    name = "nerf8_w128_l4"
    fk = meta["variants"][name]["field_kwargs"]
    sd = {k: v.to(DEV) for k, v in G.params(a, prefix=f"{name}:param:").items()}
    x = a[f"{name}:x"].to(DEV)
    f16, f32 = _field(fk, "fp16"), _field(fk, "fp32")
    f16.load_state_dict(sd, strict=False)
    f32.load_state_dict(sd, strict=False)
    with torch.no_grad():
        y16, y32 = f16(x), f32(x)
    ref = a[f"{name}:y"]
    scale = ref.abs().max().item()
    e = (y16.cpu() - ref).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)
    assert (y32.cpu() - ref).abs().max().item() < 1e-4 * max(scale, 1.0)


@pytest.mark.parametrize("W,L,O,n", [(32, 2, 4, 1000), (64, 1, 8, 129), (128, 4, 8, 4096), (16, 0, 4, 300), (48, 3, 4, 511)])
def test_fieldset_fp16_shapes_vs_oracle(W, L, O, n):
    """Tile tails, several fields (weight image switches), small/large widths."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(W * 7 + L)
    F = 5
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": O}, L, 4, W, "no")
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    pts = pos[:, None] + torch.rand(F, n, 3, generator=g) * 1.6 - 0.8
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube")
    ref = R.fieldset_forward_vmap(pts, pos, ori, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                "encoding_kwargs": {"dim_in": 3, "num_octaves": O}, "num_layers": L, "dim_out": 4,
                                "dim_mlp_out": W}, 2, 10.0, 1.0, field_radius=1.0, scale_mode="unit_cube",
                               precision="fp16").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    model.set_vmap_fields(None)
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), None, True)
    scale = ref.abs().max().item()
    e = (y.cpu() - ref).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)


@pytest.mark.parametrize("name", G.VMAP_CASES)
def test_render_fused_fp16_golden(name):
    """The fused tcgen05 render kernel against the fp32 reference's golden outputs."""
    meta, a = G.load(name)
    p = run_vmap_case(meta, a, DEV, precision="fp16")
    col = (p.rgbds[..., :3].cpu() - a["out_rgbds"][..., :3]).abs()
    dep = (p.rgbds[..., 3].cpu() - a["out_rgbds"][..., 3]).abs()
    assert col.mean().item() < 2e-3, f"colour L1 {col.mean().item():.2e} (max {col.max().item():.2e})"
    assert dep.mean().item() < 5e-3, f"depth L1 {dep.mean().item():.2e} (max {dep.max().item():.2e})"
    # per-pixel bounds: a single badly wrong pixel must fail, not vanish in the mean
    assert col.max().item() < 3e-2, f"colour max error {col.max().item():.2e} (mean {col.mean().item():.2e})"
    assert dep.max().item() < 8e-2, f"depth max error {dep.max().item():.2e} (mean {dep.mean().item():.2e})"
    assert (p.term_probs.cpu() - a["out_term_probs"]).abs().mean().item() < 3e-3
    assert (p.term_probs.cpu() - a["out_term_probs"]).abs().max().item() < 5e-2
    assert (p.color_vars.cpu() - a["out_color_vars"]).abs().mean().item() < 3e-3
    assert (p.depth_vars.cpu() - a["out_depth_vars"]).abs().mean().item() < 5e-3
    if "out_freespace" in a:
        assert p.freespace_geometry.shape == a["out_freespace"].shape
        assert (p.freespace_geometry.cpu() - a["out_freespace"]).abs().mean().item() < 2e-3
        assert (p.tsdf_residuals.cpu() - a["out_tsdf"]).abs().mean().item() < 2e-3


@pytest.mark.parametrize("S,Sg", [(128, 0), (100, 0), (5, 0), (16, 16), (70, 30), (200, 0)])
def test_render_fp16_vs_fp32_sizes(S, Sg):
    """Ray strides 8..128 (1, 2 and 4 warps per ray), padded strides, the >128-sample staged path."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=S, num_samples_depth_guided=Sg)
    g = torch.Generator().manual_seed(S + Sg)
    F, Rr = 3, 70
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    gt = near + (far - near) * torch.rand(F, Rr, generator=g)
    jit = torch.rand(F, Rr, S, generator=g)
    jg = torch.rand(F, Rr, max(Sg, 1), generator=g)[..., :Sg]
    cam = ngm.Camera(**meta["camera"])
    outs = {}
    for prec in ("fp32", "fp16"):
        st = make_state(meta, a, DEV, prec)
        with torch.no_grad():
            outs[prec] = st._render_ijs(ijs.to(DEV), a["c2ws"][:F, :1].expand(F, Rr, 4, 4).to(DEV), cam,
                                        torch.tensor([0, 3, 1], device=DEV), True, near.to(DEV), far.to(DEV),
                                        gt.to(DEV) if Sg else None, jitter=jit.to(DEV),
                                        jitter_guided=jg.to(DEV) if Sg else None)
    p32, p16 = outs["fp32"], outs["fp16"]
    assert (p16.rgbds[..., :3] - p32.rgbds[..., :3]).abs().mean().item() < 2e-3
    assert (p16.rgbds[..., 3] - p32.rgbds[..., 3]).abs().mean().item() < 5e-3
    assert (p16.term_probs - p32.term_probs).abs().mean().item() < 3e-3
    if Sg:
        assert p16.freespace_geometry.shape == p32.freespace_geometry.shape
        assert p16.tsdf_residuals.shape == p32.tsdf_residuals.shape


def test_psnr_trained_field_within_0p1_db():
    """north_star: PSNR within 0.1 dB of the reference.  Fixture: a field trained (by the oracle, CPU) on an
    analytic sphere scene; PSNR of the reference-arithmetic render vs the analytic ground truth is stored."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("trained_sphere")
    H, W = meta["image_hw"]
    S = meta["num_samples"]
    a = dict(a)
    a["jitter"] = torch.rand(1, H * W, S, generator=torch.Generator().manual_seed(meta["jitter_seed"]))
    gt = a["gt_rgb"][0].reshape(H, W, 3)
    ref_img = a["out_rgbds"][0, :, :3].reshape(H, W, 3)
    psnr_ref = R.psnr(ref_img, gt)
    assert abs(psnr_ref - meta["psnr_reference_db"]) < 1e-3
    out = {}
    for prec in ("fp32", "fp16"):
        p = run_vmap_case(meta, a, DEV, precision=prec)
        img = p.rgbds[0, :, :3].cpu().reshape(H, W, 3)
        out[prec] = (R.psnr(img, gt), R.psnr(img, ref_img), (p.rgbds[0, :, 3].cpu() - a["gt_depth"][0]).abs().mean().item())
    print("PSNR vs gt: reference %.3f dB, fp32 %.3f dB, fp16 %.3f dB; fp16 vs reference image %.1f dB; depth L1 %s" % (
        psnr_ref, out["fp32"][0], out["fp16"][0], out["fp16"][1], [round(out[k][2], 4) for k in out]))
    assert abs(out["fp32"][0] - psnr_ref) < 0.01, out
    assert abs(out["fp16"][0] - psnr_ref) < 0.1, out
    assert out["fp16"][1] > 45.0, out
    assert abs(out["fp16"][2] - meta["depth_l1_reference"]) < 5e-3


_PERMUTO_KW = dict(pos_dim=3, log2_hashmap_size=12, nr_levels=16, nr_feat_per_level=2, coarsest_scale=1.0,
                   finest_scale=1e-4, init_scale=0.5)


@pytest.mark.parametrize("W,L,levels,concat,n", [(128, 4, 16, False, 4096), (32, 1, 16, False, 1000), (64, 2, 6, True, 333),
                                               (48, 3, 10, False, 129)])
def test_fieldset_permuto_fp16_vs_oracle(W, L, levels, concat, n):
    """Permutohedral front end of the tcgen05 kernel (field stage): several fields, level counts that are not a
    multiple of four, concat_points, tile tails; against oracle/permuto.py + the oracle MLP."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(W + L + levels)
    F = 4
    ekw = dict(_PERMUTO_KW, nr_levels=levels, concat_points=concat, concat_points_scaling=0.5)
    spec = R.FieldSpec("permuto", ekw, L, 4, W, "no")
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    pts = pos[:, None] + torch.rand(F, n, 3, generator=g) * 1.6 - 0.8
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube")
    ref = R.fieldset_forward_vmap(pts, pos, ori, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PermutohedralEncoding",
                                "encoding_kwargs": ekw, "num_layers": L, "dim_out": 4, "dim_mlp_out": W}, 2, 10.0, 1.0,
                               field_radius=1.0, scale_mode="unit_cube", precision="fp16").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    model.set_vmap_fields(None)
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), None, True)
    scale = ref.abs().max().item()
    e = (y.cpu() - ref).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)


@pytest.mark.parametrize("inkernel", [False, True])
@pytest.mark.parametrize("L,W,S", [(4, 128, 64), (1, 32, 24), (2, 64, 16)])
def test_render_fused_permuto_fp16_vs_fp32(L, W, S, inkernel, monkeypatch):
    """Fused tcgen05 render with the permutohedral encoding (the reference's default encoding; 1 hidden layer x 32
    is its default MLP, neural_graph_map.yaml:6-17) against the fp32 kernels on identical rays and jitter."""
    import neural_graph_mapping_b200 as ngm

    # default: sampler kernel -> row encoder -> fused kernel reading the A operand; the env var keeps the
    # in-kernel front end (table gathers in the MMA waits)
    monkeypatch.setenv("NGM_TC_PERMUTO_INKERNEL", "1" if inkernel else "0")
    meta, a = G.load("vmap_guided_nrgbd")
    g = torch.Generator().manual_seed(L * 100 + W)
    F, Rr = 3, 300
    spec = R.FieldSpec("permuto", dict(_PERMUTO_KW), L, 4, W, "no")
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    params[f"_linears.{L}.weight"][:, 3] *= 10.0  # spread the geometry (x20 geometry_factor, nrgbd bump) so that
    params[f"_linears.{L}.bias"][:, 3] += 0.3      # rays neither all terminate at the first sample nor never
    fk = {"encoding_type": "neural_graph_mapping.positional_encodings.PermutohedralEncoding", "encoding_kwargs": dict(_PERMUTO_KW),
          "num_layers": L, "dim_out": 4, "dim_mlp_out": W, "skip_mode": "no", "initial_geometry_bias": 0.0,
          "neus_initial_sd": 1.0}
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0, field_kwargs=fk)
    arrays = {f"param:{k}": v for k, v in params.items()}
    arrays["positions"], arrays["orientations"] = a["positions"][:F], a["orientations"][:F]
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    jit = torch.rand(F, Rr, S, generator=g)
    cam = ngm.Camera(**meta["camera"])
    outs = {}
    for prec in ("fp32", "fp16"):
        st = make_state(meta, arrays, DEV, prec)
        with torch.no_grad():
            outs[prec] = st._render_ijs(ijs.to(DEV), a["c2ws"][:F, :1].expand(F, Rr, 4, 4).to(DEV), cam,
                                        torch.arange(F, device=DEV), True, near.to(DEV), far.to(DEV), None,
                                        jitter=jit.to(DEV))
    p32, p16 = outs["fp32"], outs["fp16"]
    assert torch.isfinite(p16.rgbds).all()
    assert (p16.rgbds[..., :3] - p32.rgbds[..., :3]).abs().mean().item() < 2e-3
    assert (p16.rgbds[..., 3] - p32.rgbds[..., 3]).abs().mean().item() < 5e-3
    assert (p16.term_probs - p32.term_probs).abs().mean().item() < 3e-3
    assert p32.term_probs.std().item() > 1e-3, "degenerate scene: the comparison would be vacuous"


@pytest.mark.parametrize("F,K,n,W,L", [(1, 2, 300, 16, 1), (40, 2, 5000, 128, 4), (300, 3, 4000, 64, 2)])
def test_fieldset_knn_fp16_vs_oracle(F, K, n, W, L):
    """kNN blend path with the tcgen05 field kernel in gather mode: fields without entries, a data-dependent tile
    count per field, field_ids remap; against the oracle (fp16 tolerance)."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(F * 10 + K)
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 4}, L, 4, W, "no")
    n_tab = min(F, 8)
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(n_tab)])
    pos = torch.randn(F, 3, generator=g) * (1.0 if F < 100 else 4.0)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    fid = torch.randint(0, n_tab, (F,), generator=g)
    pts = torch.randn(n, 3, generator=g) * (1.5 if F < 100 else 4.0)
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube", num_knn=K, distance_factor=10.0, outside_value=1.0)
    ref = R.fieldset_forward_knn(pts, pos, ori, fid, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
                                "encoding_kwargs": {"dim_in": 3, "num_octaves": 4}, "num_layers": L, "dim_out": 4,
                                "dim_mlp_out": W}, K, 10.0, 1.0, field_radius=1.0, scale_mode="unit_cube",
                               precision="fp16").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), fid.to(DEV), False)
    d = torch.cdist(pts, pos)
    srt = torch.sort(d, dim=-1)[0]
    margin = torch.ones(n, dtype=torch.bool)
    for j in range(min(K, F - 1)):
        margin &= (srt[:, j + 1] - srt[:, j]).abs() > 1e-5
    margin &= (srt[:, 0] - 1.0).abs() > 1e-5
    assert torch.isfinite(y).all()
    inside = (ref != 1.0).any(-1)
    assert torch.equal((y.cpu() != 1.0).any(-1)[margin], inside[margin]), "inside/outside classification differs"
    scale = ref[inside].abs().max().item()
    e = (y.cpu()[margin] - ref[margin]).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)


@pytest.mark.parametrize("name", G.KNN_CASES)
def test_render_knn_fp16_golden(name):
    """_render_ijs(use_vmap=False) -- what the untouched driver calls for keyframe renders and evaluation
    (run_mapping.py:429) -- with fp16 field evaluation, against the fp32 reference's golden outputs."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load(name)
    st = make_state(meta, a, DEV, "fp16")
    st.eval()
    cam = ngm.Camera(**meta["camera"])
    with torch.no_grad():
        p = st._render_ijs(a["ijs"].to(DEV), a["c2ws"].to(DEV), cam, jitter=a["jitter"].to(DEV))
    col = (p.rgbds[..., :3].cpu() - a["out_rgbds"][..., :3]).abs()
    dep = (p.rgbds[..., 3].cpu() - a["out_rgbds"][..., 3]).abs()
    assert col.mean().item() < 2e-3, f"colour L1 {col.mean().item():.2e} (max {col.max().item():.2e})"
    assert dep.mean().item() < 5e-3, f"depth L1 {dep.mean().item():.2e} (max {dep.max().item():.2e})"
    assert (p.term_probs.cpu() - a["out_term_probs"]).abs().mean().item() < 3e-3


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_fieldset_knn_permuto_vs_oracle(prec):
    """kNN blend path with the permutohedral encoding (the reference's default field): the (point, neighbour) entries
    are encoded by the whole-GPU row encoder and read back by the field kernels in gather mode."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(21)
    F, K, n, W, L = 40, 2, 5000, 32, 1
    ekw = dict(_PERMUTO_KW, concat_points=False)
    spec = R.FieldSpec("permuto", ekw, L, 4, W, "no")
    n_tab = 8
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(n_tab)])
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    fid = torch.randint(0, n_tab, (F,), generator=g)
    pts = torch.randn(n, 3, generator=g) * 1.5
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube", num_knn=K, distance_factor=10.0, outside_value=1.0)
    ref = R.fieldset_forward_knn(pts, pos, ori, fid, spec, params, rs)
    model = ngm.NeuralFieldSet(3, "neural_graph_mapping_b200.models.NeuralField",
                               {"encoding_type": "neural_graph_mapping_b200.positional_encodings.PermutohedralEncoding",
                                "encoding_kwargs": ekw, "num_layers": L, "dim_out": 4, "dim_mlp_out": W}, K, 10.0, 1.0,
                               field_radius=1.0, scale_mode="unit_cube", precision=prec).to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), fid.to(DEV), False)
    d = torch.cdist(pts, pos)
    srt = torch.sort(d, dim=-1)[0]
    margin = ((srt[:, 1] - srt[:, 0]).abs() > 1e-5) & ((srt[:, 0] - 1.0).abs() > 1e-5)
    inside = (ref != 1.0).any(-1)
    scale = ref[inside].abs().max().item()
    e = (y.cpu()[margin] - ref[margin]).abs()
    if prec == "fp32":
        # the finest lattice levels (cell 1e-4, features of order 0.5) have gradients of ~5e3 per unit length: a
        # last-bit difference in the rotated local coordinates (FMA contraction vs torch) shows up at ~1e-4
        assert e.max().item() < 2e-3 * max(scale, 1.0) and e.mean().item() < 1e-4 * max(scale, 1.0), (e.max().item(), e.mean().item())
    else:
        assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)


_ROW_ENCODINGS = [
    ("fourier", "PositionalEncodingFourier", {"dim_in": 3, "dim_out": 35, "mu": 0.0, "sigma": 2.0, "raw_coords": True}),
    ("triplane", "TriplaneEncoding", {"resolution": 16, "num_components": 32, "init_scale": 0.5, "mode": "sum"}),
    ("nerf", "PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 6}),
]


@pytest.mark.parametrize("kind,enc_cls,ekw", _ROW_ENCODINGS)
def test_row_encoded_fields_fp16(kind, enc_cls, ekw):
    """Encodings without an in-kernel front end (Fourier, Triplane, NeRF with 6 octaves) on the tcgen05 path through
    pre-encoded fp16 rows: dense field evaluation, the fused renderer and the kNN path, against the fp32 kernels."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    g = torch.Generator().manual_seed(len(kind))
    F, Rr, S, W, L = 3, 200, 32, 64, 2
    spec = R.FieldSpec(kind, dict(ekw), L, 4, W, "no")
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    params[f"_linears.{L}.weight"][:, 3] *= 10.0
    params[f"_linears.{L}.bias"][:, 3] += 0.3
    fk = {"encoding_type": f"neural_graph_mapping.positional_encodings.{enc_cls}", "encoding_kwargs": dict(ekw),
          "num_layers": L, "dim_out": 4, "dim_mlp_out": W, "skip_mode": "no", "initial_geometry_bias": 0.0,
          "neus_initial_sd": 1.0}
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0, field_kwargs=fk)
    arrays = {f"param:{k}": v for k, v in params.items()}
    arrays["positions"], arrays["orientations"] = a["positions"][:F], a["orientations"][:F]
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    jit = torch.rand(F, Rr, S, generator=g)
    cam = ngm.Camera(**meta["camera"])
    pts = arrays["positions"][:, None] + torch.rand(F, 700, 3, generator=g) * 1.6 - 0.8
    out = {}
    for prec in ("fp32", "fp16"):
        st = make_state(meta, arrays, DEV, prec)
        with torch.no_grad():
            pr = st._render_ijs(ijs.to(DEV), a["c2ws"][:F, :1].expand(F, Rr, 4, 4).to(DEV), cam, torch.arange(F, device=DEV),
                                True, near.to(DEV), far.to(DEV), None, jitter=jit.to(DEV))
            st._model.set_vmap_fields(None)
            dense = st._model(pts.to(DEV), arrays["positions"].to(DEV), arrays["orientations"].to(DEV), None, True)
            knn = st._model(pts.reshape(-1, 3).to(DEV), arrays["positions"].to(DEV), arrays["orientations"].to(DEV), None, False)
        out[prec] = (pr, dense, knn)
    p32, d32, k32 = out["fp32"]
    p16, d16, k16 = out["fp16"]
    scale = d32.abs().max().item()
    assert (d16 - d32).abs().max().item() < 2e-2 * scale and (d16 - d32).abs().mean().item() < 3e-3 * scale
    assert (k16 - k32).abs().mean().item() < 3e-3 * scale
    assert (p16.rgbds[..., :3] - p32.rgbds[..., :3]).abs().mean().item() < 2e-3
    assert (p16.rgbds[..., 3] - p32.rgbds[..., 3]).abs().mean().item() < 5e-3
    assert (p16.term_probs - p32.term_probs).abs().mean().item() < 3e-3


# ---- skip connections "add" / "concat" on the tcgen05 path (ngm/models.py:160-170) ----

def test_field_fwd_fp16_skip_add_golden():
    """The reference's own output of a NeRF-4 field with skip_mode "add" (2 x 48), through the tcgen05 kernel; and the
    concat golden (width 40, not a multiple of 16) must be refused loudly rather than silently run elsewhere."""
    meta, a = G.load("fields_forward")
    name = "nerf4_add"
    fk = meta["variants"][name]["field_kwargs"]
    assert fk["skip_mode"] == "add"
    f16 = _field(fk, "fp16")
    f16.load_state_dict({k: v.to(DEV) for k, v in G.params(a, prefix=f"{name}:param:").items()}, strict=False)
    with torch.no_grad():
        y16 = f16(a[f"{name}:x"].to(DEV))
    ref = a[f"{name}:y"]
    scale = ref.abs().max().item()
    e = (y16.cpu() - ref).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (e.max().item(), e.mean().item(), scale)
    fk = meta["variants"]["nerf4_concat"]["field_kwargs"]
    bad = _field(fk, "fp16")
    bad.load_state_dict({k: v.to(DEV) for k, v in G.params(a, prefix="nerf4_concat:param:").items()}, strict=False)
    with pytest.raises(NotImplementedError), torch.no_grad():
        bad(a["nerf4_concat:x"].to(DEV))


@pytest.mark.parametrize("skip", ["add", "concat"])
@pytest.mark.parametrize("W,L,O,n", [(32, 2, 4, 1000), (128, 3, 8, 700), (64, 1, 8, 129), (128, 4, 4, 2500)])
def test_fieldset_fp16_skip_vs_oracle(skip, W, L, O, n):
    """Stacked fields with skip connections on the tcgen05 kernel (pre-encoded rows; "concat" widens K of every linear
    after the first to W + E) against the oracle: tile tails, several fields, widths 32..128."""
    import neural_graph_mapping_b200 as ngm

    g = torch.Generator().manual_seed(W * 7 + L + (1 if skip == "add" else 2))
    F = 4
    spec = R.FieldSpec("nerf", {"dim_in": 3, "num_octaves": O}, L, 4, W, skip)
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
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
                               scale_mode="unit_cube", precision="fp16").to(DEV)
    model.all_fields_params = {k: v.to(DEV) for k, v in params.items()}
    model.set_vmap_fields(None)
    with torch.no_grad():
        y = model(pts.to(DEV), pos.to(DEV), ori.to(DEV), None, True)
    scale = ref.abs().max().item()
    e = (y.cpu() - ref).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (skip, e.max().item(), e.mean().item(), scale)
    # the kNN blend path runs the same kernel in gather mode
    rs2 = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube", num_knn=2, distance_factor=10.0, outside_value=1.0)
    q_pts = pts.reshape(-1, 3)[:: max(1, (F * n) // 1500)]
    fid = torch.arange(F)
    ref_k = R.fieldset_forward_knn(q_pts, pos, ori, fid, spec, params, rs2)
    with torch.no_grad():
        y_k = model(q_pts.to(DEV), pos.to(DEV), ori.to(DEV), fid.to(DEV), False)
    d = torch.sort(torch.cdist(q_pts, pos), dim=-1)[0]
    margin = ((d[:, 1] - d[:, 0]).abs() > 1e-5) & ((d[:, 0] - 1.0).abs() > 1e-5)
    if F > 2:
        margin &= (d[:, 2] - d[:, 1]).abs() > 1e-5
    e = (y_k.cpu()[margin] - ref_k[margin]).abs()
    assert e.max().item() < 2e-2 * scale and e.mean().item() < 3e-3 * scale, (skip, "knn", e.max().item(), e.mean().item())


@pytest.mark.parametrize("skip", ["add", "concat"])
def test_render_fused_fp16_skip_vs_fp32(skip):
    """The fused tcgen05 renderer with skip connections (sampler -> row encoder -> fused kernel, like the
    permutohedral path) against the fp32 kernels on identical rays and jitter."""
    import neural_graph_mapping_b200 as ngm

    L, W, S = 3, 64, 24
    meta, a = G.load("vmap_guided_nrgbd")
    g = torch.Generator().manual_seed(77)
    F, Rr = 3, 300
    ekw = {"dim_in": 3, "num_octaves": 8}
    spec = R.FieldSpec("nerf", dict(ekw), L, 4, W, skip)
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    params[f"_linears.{L}.weight"][:, 3] *= 3.0
    params[f"_linears.{L}.bias"][:, 3] += 0.3
    fk = {"encoding_type": "neural_graph_mapping.positional_encodings.PositionalEncodingNeRF", "encoding_kwargs": dict(ekw),
          "num_layers": L, "dim_out": 4, "dim_mlp_out": W, "skip_mode": skip, "initial_geometry_bias": 0.0,
          "neus_initial_sd": 1.0}
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0, field_kwargs=fk)
    arrays = {f"param:{k}": v for k, v in params.items()}
    arrays["positions"], arrays["orientations"] = a["positions"][:F], a["orientations"][:F]
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    jit = torch.rand(F, Rr, S, generator=g)
    cam = ngm.Camera(**meta["camera"])
    outs = {}
    for prec in ("fp32", "fp16"):
        st = make_state(meta, arrays, DEV, prec)
        with torch.no_grad():
            outs[prec] = st._render_ijs(ijs.to(DEV), a["c2ws"][:F, :1].expand(F, Rr, 4, 4).to(DEV), cam,
                                        torch.arange(F, device=DEV), True, near.to(DEV), far.to(DEV), None,
                                        jitter=jit.to(DEV))
    p32, p16 = outs["fp32"], outs["fp16"]
    assert torch.isfinite(p16.rgbds).all()
    assert (p16.rgbds[..., :3] - p32.rgbds[..., :3]).abs().mean().item() < 2e-3
    assert (p16.rgbds[..., 3] - p32.rgbds[..., 3]).abs().mean().item() < 5e-3
    assert (p16.term_probs - p32.term_probs).abs().mean().item() < 3e-3
    assert p32.term_probs.std().item() > 1e-3, "degenerate scene: the comparison would be vacuous"
