"""tcgen05 backward of the per-field MLP (csrc/field_tc_bwd.cu, `ngm_field_bwd`) against torch.autograd through the
oracle restatement (fp32, CPU): the gradients of every linear's weight and bias, per tensor, so a failure names the
GEMM family that is wrong (dW_L^T: N = 16 tile; chain: MN-major weight image; dW_l: MN-major activation buffers;
db_l: ones tile; db_L: fp32 shuffle sums)."""
import pytest
import torch

from oracle import restatement as R

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _proto(enc, ekw, L, W):
    import neural_graph_mapping_b200 as ngm

    cls = {"nerf": "PositionalEncodingNeRF", "permuto": "PermutohedralEncoding"}[enc]
    return ngm.NeuralField(f"neural_graph_mapping_b200.positional_encodings.{cls}", dict(ekw), L, 4, W)


def _case(enc, ekw, L, W, F, N, seed, gscale):
    g = torch.Generator().manual_seed(seed)
    spec = R.FieldSpec(enc, dict(ekw), L, 4, W, "no")
    params = R.stack_params([R.init_field_params(spec, g) for _ in range(F)])
    for k in params:  # trained-scale weights: keep activations alive through the layers
        if k.endswith("weight"):
            params[k] = params[k] * 1.5
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    pts = pos[:, None] + torch.rand(F, N, 3, generator=g) * 1.6 - 0.8
    d_out = torch.randn(F, N, 4, generator=g) * gscale
    d_out[:, ::7] *= 30.0  # a wide dynamic range inside one call
    return spec, params, pos, ori, pts, d_out


def _reference_grads(spec, params, pos, ori, pts, d_out):
    rs = R.RenderSpec(field_radius=1.0, scale_mode="unit_cube")
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and ("_linears" in k or "lattice_values" in k))
         for k, v in params.items()}
    out = R.fieldset_forward_vmap(pts, pos, ori, spec, p, rs)
    out.backward(d_out)
    return out.detach(), {k: v.grad for k, v in p.items() if v.grad is not None}


def _check(name, got, ref, rel_max, rel_mean):
    """Errors relative to max |ref|.  The forward is recomputed with fp16 operands, so a point whose pre-activation
    is within fp16 rounding of zero can land on the other side of the ReLU than in the fp32 reference and its whole
    contribution to one gradient entry flips: with a few hundred points per field and the 30x outliers of these
    fixtures a single flip is several percent of one entry.  Hence a loose bound on the worst entry and tight bounds
    on the mean and on the relative Frobenius error (6e-2: four fp16 layers deep the flips add up to ~3e-2), which a
    wrong GEMM layout would miss by orders of magnitude."""
    scale = ref.abs().max().item()
    assert scale > 0, name
    e = (got.cpu() - ref).abs()
    fro = (got.cpu() - ref).norm().item() / ref.norm().item()
    assert torch.isfinite(got).all(), f"{name}: non-finite gradient"
    assert e.max().item() <= rel_max * scale and e.mean().item() <= rel_mean * scale and fro <= 6e-2, \
        f"{name}: max err {e.max().item() / scale:.3e} mean err {e.mean().item() / scale:.3e} of max |ref| {scale:.3e}, " \
        f"relative Frobenius error {fro:.3e}; worst at {tuple(torch.nonzero(e == e.max())[0].tolist())}"


CASES = [
    # enc, kwargs, L, W, F, N, upstream scale
    ("nerf", {"dim_in": 3, "num_octaves": 4}, 1, 32, 2, 300, 1e-7),     # 64-wide buffers, M = 64 weight-gradient GEMMs
    ("nerf", {"dim_in": 3, "num_octaves": 8}, 1, 128, 2, 384, 1.0),     # one hidden layer of 128
    ("nerf", {"dim_in": 3, "num_octaves": 4}, 2, 64, 3, 257, 1e-3),
    ("nerf", {"dim_in": 3, "num_octaves": 8}, 3, 128, 2, 700, 1e-8),    # single launch, three accumulators
    ("nerf", {"dim_in": 3, "num_octaves": 8}, 4, 128, 3, 128 * 5 + 17, 1e-7),  # BASELINE field: two launches
    ("nerf", {"dim_in": 3, "num_octaves": 6}, 2, 96, 2, 200, 1e-5),     # NeRF rows from the library's row encoder
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}{c[1].get('num_octaves', '')}_L{c[2]}_W{c[3]}" for c in CASES])
@pytest.mark.parametrize("max_ctas", [0, 2])
def test_field_bwd_tc_vs_autograd(case, max_ctas, monkeypatch):
    from neural_graph_mapping_b200 import autograd as ag

    enc, ekw, L, W, F, N, gs = case
    if max_ctas:
        monkeypatch.setenv("NGM_TC_MAX_CTAS", str(max_ctas))  # few CTAs: several field segments and tiles per CTA
    spec, params, pos, ori, pts, d_out = _case(enc, ekw, L, W, F, N, seed=L * 100 + W, gscale=gs)
    ref_out, ref = _reference_grads(spec, params, pos, ori, pts, d_out)
    proto = _proto(enc, ekw, L, W)
    assert ag.tc_training_supported(proto)
    p = {k: v.to(DEV).requires_grad_(True) for k, v in params.items()}
    out = ag.field_forward_tc(proto, p, pts.to(DEV), pos.to(DEV), ori.to(DEV), "unit_cube", 1.0)
    sc = ref_out.abs().max().item()
    assert (out.detach().cpu() - ref_out).abs().max().item() < 2e-2 * sc
    out.backward(d_out.to(DEV))
    torch.cuda.synchronize()
    # order: last bias (fp32 sums), last weight, then down the chain
    for i in range(L, -1, -1):
        _check(f"d _linears.{i}.bias", p[f"_linears.{i}.bias"].grad, ref[f"_linears.{i}.bias"], 3.5e-1, 8e-3)
        _check(f"d _linears.{i}.weight", p[f"_linears.{i}.weight"].grad, ref[f"_linears.{i}.weight"], 3.5e-1, 8e-3)


def test_field_bwd_tc_permuto_table_gradient():
    """The reference's default field (permutohedral 16 x 2, one hidden layer of 32): encoding rows from
    ngm_encode_fwd, dLoss/d encoding from the tcgen05 backward, table gradient from ngm_encode_bwd."""
    from neural_graph_mapping_b200 import autograd as ag

    ekw = {"pos_dim": 3, "log2_hashmap_size": 12, "nr_levels": 16, "nr_feat_per_level": 2, "coarsest_scale": 1.0, "finest_scale": 0.0001,
           "init_scale": 0.3}
    L, W, F, N = 1, 32, 2, 500
    spec, params, pos, ori, pts, d_out = _case("permuto", ekw, L, W, F, N, seed=5, gscale=1e-6)
    ref_out, ref = _reference_grads(spec, params, pos, ori, pts, d_out)
    proto = _proto("permuto", ekw, L, W)
    p = {k: (v.to(DEV).requires_grad_(True) if k in ref else v.to(DEV)) for k, v in params.items()}
    out = ag.field_forward_tc(proto, p, pts.to(DEV), pos.to(DEV), ori.to(DEV), "unit_cube", 1.0)
    assert (out.detach().cpu() - ref_out).abs().max().item() < 2e-2 * ref_out.abs().max().item()
    out.backward(d_out.to(DEV))
    for k, r in ref.items():
        _check(f"d {k}", p[k].grad, r, 1.5e-1, 2e-3)
