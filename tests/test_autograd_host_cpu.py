"""Host-side (pure torch) pieces of the training path on CPU: world->local, the PyTorch encodings and the
batched-GEMM MLP of neural_graph_mapping_b200.autograd against the oracle restatement, forward and gradients.
(The CUDA pieces -- sampler, permutohedral encode, compositor fwd/bwd -- are covered by the gpu tests.)"""
import pytest
import torch

from oracle import restatement as R


def _proto(enc_cls, enc_kwargs, L, W, skip):
    from neural_graph_mapping_b200.models import NeuralField

    return NeuralField(f"neural_graph_mapping_b200.positional_encodings.{enc_cls}", enc_kwargs, L, 4, W, skip)


CASES = [
    ("nerf", "PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 4}, 2, 32, "no"),
    ("nerf", "PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 4, "start_octave": 1}, 3, 40, "concat"),
    ("nerf", "PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 3}, 2, 24, "add"),
    ("nerf", "PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 3}, 3, 18, "rezero"),
    ("fourier", "PositionalEncodingFourier", {"dim_in": 3, "dim_out": 19, "mu": 0.0, "sigma": 2.0, "raw_coords": True}, 1, 16, "no"),
    ("triplane", "TriplaneEncoding", {"resolution": 8, "num_components": 6, "init_scale": 0.5, "mode": "sum"}, 1, 16, "no"),
    ("triplane", "TriplaneEncoding", {"resolution": 8, "num_components": 4, "init_scale": 0.5, "mode": "product"}, 1, 16, "no"),
    ("triplane", "TriplaneEncoding", {"resolution": 8, "num_components": 4, "init_scale": 0.5, "mode": "concat"}, 1, 16, "no"),
]


@pytest.mark.parametrize("kind,enc_cls,ekw,L,W,skip", CASES)
@pytest.mark.parametrize("scale_mode", ["unit_cube", "unit_ball", "no"])
def test_training_field_forward_and_gradients_vs_oracle(kind, enc_cls, ekw, L, W, skip, scale_mode):
    from neural_graph_mapping_b200 import autograd as ag

    g = torch.Generator().manual_seed(L * 31 + W)
    F, N = 3, 50
    spec = R.FieldSpec(kind, dict(ekw), L, 4, W, skip)
    per_field = [R.init_field_params(spec, g) for _ in range(F)]
    if skip == "rezero":
        for p in per_field:
            p["_rezero"] = torch.rand(L, generator=g)
    params = R.stack_params(per_field)
    pos = torch.randn(F, 3, generator=g)
    q = torch.randn(F, 4, generator=g)
    ori = q / q.norm(dim=-1, keepdim=True)
    span = 1.6 if kind != "triplane" else 2.6  # triplane: also outside [-1, 1] (border padding)
    pts = pos[:, None] + torch.rand(F, N, 3, generator=g) * span - span / 2
    rs = R.RenderSpec(field_radius=1.0, scale_mode=scale_mode)
    w_out = torch.randn(F, N, 4, generator=g)

    ref_params = {k: v.clone().requires_grad_(v.dtype == torch.float32) for k, v in params.items()}
    ref = R.fieldset_forward_vmap(pts, pos, ori, spec, ref_params, rs)
    (ref * w_out).sum().backward()

    proto = _proto(enc_cls, ekw, L, W, skip)
    our_params = {k: v.clone().requires_grad_(v.dtype == torch.float32) for k, v in params.items()}
    local = ag.world_to_local(pts, pos, ori, scale_mode, 1.0)
    out = ag.field_forward(proto, our_params, local)
    (out * w_out).sum().backward()

    assert torch.allclose(out, ref, atol=2e-5, rtol=2e-5)
    for k, v in ref_params.items():
        if v.grad is None:
            assert our_params[k].grad is None or float(our_params[k].grad.abs().max()) == 0.0, k
            continue
        scale = v.grad.abs().max().item() + 1e-12
        assert (our_params[k].grad - v.grad).abs().max().item() <= 1e-4 * scale + 1e-7, k


def test_unknown_scale_mode_raises_like_reference():
    from neural_graph_mapping_b200 import autograd as ag

    with pytest.raises(ValueError, match="is not available"):  # models.py:285
        ag.world_to_local(torch.zeros(1, 2, 3), torch.zeros(1, 3), torch.tensor([[1.0, 0, 0, 0]]), "cube", 1.0)
