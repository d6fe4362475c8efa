"""Edge cases of the render surface: empty batches, a single ray / a single sample, strided and broadcast
(expand-ed) inputs the reference passes (camera.py:265-267, run_mapping.py:547), ray counts that do not fill a
tile, the 128-sample limit of the fused kernel."""
import pytest
import torch

import golden_util as G
from oracle import restatement as R
from tests_support import make_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _scene(F, Rr, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1)
    near = torch.rand(F, Rr, generator=g) * 0.5 + 0.3
    far = near + 1.5
    jit = torch.rand(F, Rr, S, generator=g)
    return ijs, near, far, jit


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("F,Rr", [(0, 16), (2, 0)])
def test_empty_batches(prec, F, Rr):
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    st = make_state(meta, a, DEV, prec)
    cam = ngm.Camera(**meta["camera"])
    ijs = torch.zeros(F, Rr, 2, dtype=torch.int64, device=DEV)
    with torch.no_grad():
        p = st._render_ijs(ijs, a["c2ws"][0, 0].to(DEV), cam, torch.arange(F, device=DEV), True,
                           torch.ones(F, Rr, device=DEV), torch.full((F, Rr), 2.0, device=DEV))
    assert p.rgbds.shape == (F, Rr, 4) and p.color_vars.shape == (F, Rr, 3)
    assert p.depth_vars.shape == (F, Rr) and p.term_probs.shape == (F, Rr)
    torch.cuda.synchronize()


@pytest.mark.parametrize("F,Rr,S", [(1, 1, 1), (3, 1, 7), (2, 5, 128), (2, 3, 129), (5, 17, 33)])
def test_small_and_limit_shapes_all_paths_agree(F, Rr, S):
    """fp32 kernels vs the oracle, and fp16 (fused up to 128 samples, staged beyond) vs fp32."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0)
    ijs, near, far, jit = _scene(F, Rr, S, seed=F * 100 + S)
    cam = ngm.Camera(**meta["camera"])
    fid = torch.arange(F)
    c2w = a["c2ws"][0, 0]
    outs = {}
    for prec in ("fp32", "fp16"):
        st = make_state(meta, a, DEV, prec)
        with torch.no_grad():
            outs[prec] = st._render_ijs(ijs.to(DEV), c2w.to(DEV), cam, fid.to(DEV), True, near.to(DEV), far.to(DEV),
                                        jitter=jit.to(DEV))
    fs, rs, cs = G.field_spec(meta["field_kwargs"]), G.render_spec(meta), G.camera_spec(meta["camera"])
    ref = R.render_rays(ijs, c2w, cs, rs, fs, G.params(a), a["positions"], a["orientations"], field_ids=fid,
                        use_vmap=True, near_distances=near, far_distances=far, jitter=jit)
    p32, p16 = outs["fp32"], outs["fp16"]
    assert torch.allclose(p32.rgbds.cpu(), ref.rgbds, atol=1e-4, rtol=1e-4)
    assert torch.allclose(p32.term_probs.cpu(), ref.term_probs, atol=3e-5, rtol=3e-5)
    assert torch.allclose(p32.color_vars.cpu(), ref.color_vars, atol=3e-5, rtol=3e-5)
    assert (p16.rgbds - p32.rgbds).abs().max().item() < 3e-2
    assert (p16.term_probs - p32.term_probs).abs().max().item() < 3e-2


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_strided_and_broadcast_inputs(prec):
    """ijs as a strided view, near/far as expand-ed scalars, c2ws as an expand-ed (F,R,4,4) view of one pose:
    same result as with contiguous copies."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    S = 16
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0)
    F, Rr = 3, 40
    ijs, near, far, jit = _scene(F, 2 * Rr, S, seed=5)
    st = make_state(meta, a, DEV, prec)
    cam = ngm.Camera(**meta["camera"])
    fid = torch.arange(F, device=DEV)
    ijs_view = ijs.to(DEV)[:, ::2]                       # stride 2 along the ray axis
    near_b = torch.tensor(0.5, device=DEV).expand(F, Rr)  # stride-0 tensors
    far_b = torch.tensor(2.0, device=DEV).expand(F, Rr)
    c2w_b = a["c2ws"][0, 0].to(DEV).expand(F, Rr, 4, 4)
    jv = jit.to(DEV)[:, ::2]
    with torch.no_grad():
        p_view = st._render_ijs(ijs_view, c2w_b, cam, fid, True, near_b, far_b, jitter=jv)
        p_cont = st._render_ijs(ijs_view.contiguous(), a["c2ws"][0, 0].to(DEV), cam, fid, True, near_b.contiguous(),
                                far_b.contiguous(), jitter=jv.contiguous())
    assert torch.equal(p_view.rgbds, p_cont.rgbds) and torch.equal(p_view.term_probs, p_cont.term_probs)
    assert not ijs_view.is_contiguous() and not near_b.is_contiguous()


def test_knn_empty_and_outside():
    """use_vmap=False: zero query points, and points outside every field radius get outside_value in all 4 channels
    (models.py:401) on both field kernels."""
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("knn_render")
    for prec in ("fp32", "fp16"):
        st = make_state(meta, a, DEV, prec)
        st.eval()
        pos, ori = a["positions"].to(DEV), a["orientations"].to(DEV)
        with torch.no_grad():
            empty = st._model(torch.zeros(0, 3, device=DEV), pos, ori, None, False)
            far_pts = torch.full((300, 3), 1e3, device=DEV) + torch.rand(300, 3, device=DEV)
            out = st._model(far_pts, pos, ori, None, False)
        assert empty.shape == (0, 4)
        assert torch.equal(out, torch.full_like(out, float(st._model._outside_value)))
