"""Training path: gradients of the CUDA render path against torch.autograd through the oracle
restatement (CPU fp32) on the golden fixtures -- the quantity the reference's training step
consumes (ngm/run_mapping.py:1164-1186: loss.backward() must reach vmap_fields_params)."""
import pytest
import torch

import golden_util as G
from oracle import restatement as R
from tests_support import make_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _loss(pred, gen):
    """A fixed random linear functional of every Prediction member (plus squares, so that second-order
    terms of the compositor are exercised too)."""
    total = 0.0
    for t in pred:
        if t is None:
            continue
        w = torch.randn(t.shape, generator=gen).to(t.device)
        total = total + (w * t).sum() + 0.5 * (w * t * t).sum()
    return total


def _grad(t):
    """Parameters the loss does not depend on (e.g. _neus_sd outside neus mode) have no .grad."""
    return torch.zeros_like(t) if t.grad is None else t.grad.detach()


def _grads_ours(meta, a, requires):
    import neural_graph_mapping_b200 as ngm

    st = make_state(meta, a, DEV)
    cam = ngm.Camera(**meta["camera"])
    params = st._model.all_fields_params
    for k in requires:
        params[k].requires_grad_(True)
    d = lambda k: a[k].to(DEV) if k in a else None  # noqa: E731
    pred = st._render_ijs(d("ijs"), d("c2ws"), cam, d("field_ids"), True, d("near"), d("far"), d("gt"),
                          jitter=d("jitter"), jitter_guided=d("jitter_guided"))
    loss = _loss(pred, torch.Generator().manual_seed(5))
    loss.backward()
    return pred, {k: _grad(params[k]).cpu() for k in requires}


def _grads_oracle(meta, a, requires):
    fs, rs, cam = G.field_spec(meta["field_kwargs"]), G.render_spec(meta), G.camera_spec(meta["camera"])
    params = {k: v.clone() for k, v in G.params(a).items()}
    for k in requires:
        params[k].requires_grad_(True)
    g = lambda k: a[k] if k in a else None  # noqa: E731
    pred = R.render_rays(a["ijs"], a["c2ws"], cam, rs, fs, params, a["positions"], a["orientations"],
                         field_ids=a["field_ids"], use_vmap=True, near_distances=g("near"), far_distances=g("far"),
                         gt_distances=g("gt"), jitter=g("jitter"), jitter_guided=g("jitter_guided"))
    loss = _loss(pred, torch.Generator().manual_seed(5))
    loss.backward()
    return pred, {k: _grad(params[k]) for k in requires}


@pytest.mark.parametrize("name", ["c1_vmap_256x32", "vmap_guided_nrgbd", "vmap_behind_camera", "vmap_neus",
                                  "vmap_density", "vmap_occupancy"])
def test_render_gradients_vs_oracle(name):
    meta, a = G.load(name)
    requires = sorted(k for k, v in G.params(a).items() if v.dtype == torch.float32 and not k.endswith("random_shift_per_level"))
    pred, ours = _grads_ours(meta, a, requires)
    ref_pred, ref = _grads_oracle(meta, a, requires)
    # forward of the differentiable path = forward of the oracle
    for o, r_ in zip(pred, ref_pred):
        if r_ is None:
            assert o is None
            continue
        assert o.shape == r_.shape
        assert torch.allclose(o.detach().cpu(), r_.detach(), atol=1e-4, rtol=1e-4)
    for k in requires:
        r_ = ref[k]
        scale = r_.abs().max().item() + 1e-12
        err = (ours[k] - r_).abs().max().item()
        assert err <= 2e-4 * scale + 1e-6, f"{name}: d/d{k}: max abs err {err:.3e} at gradient scale {scale:.3e}"


def test_no_grad_keeps_fused_path():
    """Under torch.no_grad (evaluation, run_mapping.py:402) parameters that require grad must not
    switch the renderer to the differentiable path."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("c1_vmap_256x32")
    st = make_state(meta, a, DEV)
    for v in st._model.all_fields_params.values():
        if v.dtype == torch.float32:
            v.requires_grad_(True)
    cam = ngm.Camera(**meta["camera"])
    d = lambda k: a[k].to(DEV) if k in a else None  # noqa: E731
    with torch.no_grad():
        p = st._render_ijs(d("ijs"), d("c2ws"), cam, d("field_ids"), True, d("near"), d("far"), jitter=d("jitter"))
    assert not p.rgbds.requires_grad
    assert torch.allclose(p.rgbds.cpu(), a["out_rgbds"], atol=2e-4)


def test_permuto_table_gradient_vs_oracle():
    """ngm_encode_fwd / ngm_encode_bwd (permutohedral) against autograd through oracle/permuto.py."""
    from neural_graph_mapping_b200 import autograd as ag
    from neural_graph_mapping_b200.models import NeuralField
    from oracle import permuto as P

    torch.manual_seed(0)
    kw = dict(pos_dim=3, log2_hashmap_size=8, nr_levels=6, nr_feat_per_level=2, coarsest_scale=1.0,
              finest_scale=0.05, concat_points=True, concat_points_scaling=0.5, init_scale=1.0)
    proto = NeuralField("neural_graph_mapping_b200.positional_encodings.PermutohedralEncoding", kw, 1, 4, 16)
    F, N = 3, 500
    enc = proto._encoding
    table = (torch.rand(F, 6, 256, 2) * 2 - 1).requires_grad_(True)
    shift = torch.randn(F, 6, 3)
    pts = torch.rand(F, N, 3)
    w = torch.randn(F, N, enc.get_out_dim())
    # oracle
    ref = torch.stack([P.encode(pts[f], table[f], shift[f], enc.scale_factor, True, 0.5) for f in range(F)])
    (ref * w).sum().backward()
    g_ref = table.grad.clone()
    # CUDA
    t2 = table.detach().clone().to(DEV).requires_grad_(True)
    params = {"_encoding.lattice_values": t2, "_encoding.random_shift_per_level": shift.to(DEV),
              "_linears.0.weight": torch.zeros(F, 16, enc.get_out_dim(), device=DEV), "_linears.0.bias": torch.zeros(F, 16, device=DEV),
              "_linears.1.weight": torch.zeros(F, 4, 16, device=DEV), "_linears.1.bias": torch.zeros(F, 4, device=DEV)}
    proto.to(DEV)
    out = ag.encode(proto, params, pts.to(DEV))
    assert torch.allclose(out.detach().cpu(), ref.detach(), atol=1e-6)
    (out * w.to(DEV)).sum().backward()
    assert torch.allclose(t2.grad.cpu(), g_ref, atol=1e-5, rtol=1e-5)


def test_training_field_forward_golden():
    """The differentiable field evaluation (autograd.field_forward: torch encodings + batched GEMMs) against the
    reference's own outputs for every in-tree encoding and skip mode, and ngm_encode_fwd (CUDA) against the
    torch encodings on the same points."""
    import ctypes as C

    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import _lib
    from neural_graph_mapping_b200 import autograd as ag
    from tests_support import product_field_kwargs

    meta, a = G.load("fields_forward")
    for name, info in meta["variants"].items():
        fld = ngm.NeuralField(**product_field_kwargs(info["field_kwargs"])).to(DEV)
        sd = {k: v.to(DEV) for k, v in G.params(a, prefix=f"{name}:param:").items()}
        fld.load_state_dict(sd, strict=False)
        params = {k: v.detach()[None].contiguous().requires_grad_(v.dtype == torch.float32) for k, v in fld._own_params().items()}
        x = a[f"{name}:x"].to(DEV).reshape(1, -1, 3)
        y = ag.field_forward(fld, params, x)
        ref = a[f"{name}:y"].reshape(1, -1, y.shape[-1])
        scale = max(ref.abs().max().item(), 1.0)
        assert torch.allclose(y.detach().cpu(), ref, atol=3e-5 * scale, rtol=3e-5), name
        y.sum().backward()  # every parameter that takes part gets a finite gradient
        assert all(p.grad is None or torch.isfinite(p.grad).all() for p in params.values()), name
        # CUDA stand-alone encoding == torch encoding
        enc_t = ag.encode(fld, {k: v.detach() for k, v in params.items()}, x)
        e = _lib.NgmEncodeArgs()
        e.field, keep = fld.field_desc({k: v.detach() for k, v in params.items()}, True)
        out = torch.empty(enc_t.shape, device=DEV)  # (enc_t may be a transposed view)
        xx = x.contiguous()
        e.points_per_field, e.num_fields = xx.shape[1], 1
        e.points, e.out = xx.data_ptr(), out.data_ptr()
        _lib.check(_lib.lib.ngm_encode_fwd(C.byref(e), _lib.stream_ptr(torch.device(DEV))))
        assert torch.allclose(out, enc_t, atol=2e-5, rtol=2e-5), f"{name}: ngm_encode_fwd vs torch encoding"

