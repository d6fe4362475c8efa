"""Persistent kernel-friendly weight layout (`ngm_pack_weights`): the tensor-core kernels read cached
pre-swizzled fp16 images of the stacked tables (ngm/models.py:254-264 stays the source of truth); the cache follows
in-place torch updates (version counters), the CUDA Adam step (rows re-packed) and new tables (checkpoint load)."""
import pytest
import torch

import golden_util as G
from tests_support import make_state

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _setup():
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=16, num_samples_depth_guided=0)
    g = torch.Generator().manual_seed(3)
    F, Rr = 3, 80
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1).to(DEV)
    near = (torch.rand(F, Rr, generator=g) * 0.5 + 0.3).to(DEV)
    far = near + 1.5
    jit = torch.rand(F, Rr, 16, generator=g).to(DEV)
    fid = torch.tensor([4, 0, 2], device=DEV)
    cam = ngm.Camera(**meta["camera"])
    c2w = a["c2ws"][0, 0].to(DEV)

    def render(st):
        with torch.no_grad():
            return st._render_ijs(ijs, c2w, cam, fid, True, near, far, jitter=jit)

    return meta, a, render, fid


def _fresh_like(meta, a, st):
    """A new state holding copies of st's current tables (no cache)."""
    st2 = make_state(meta, a, DEV, "fp16")
    st2._model.all_fields_params = {k: v.detach().clone() for k, v in st._model.all_fields_params.items()}
    return st2


def test_cache_is_used_and_follows_updates(tmp_path, render_pipeline):
    from neural_graph_mapping_b200 import _lib, optim

    meta, a, render, fid = _setup()
    st = make_state(meta, a, DEV, "fp16")
    n0 = _lib.lib.ngm_launch_count()
    p1 = render(st)
    n1 = _lib.lib.ngm_launch_count()
    p2 = render(st)
    n2 = _lib.lib.ngm_launch_count()
    cache = st._model._packed_cache
    assert cache is not None and cache[1] is not None
    assert torch.equal(p1.rgbds, p2.rgbds)
    # first call: pack + render; second call: render only (one fused kernel, or sampler + field kernel + compositor)
    per_render = 1 if render_pipeline == "fused" else 3
    assert n2 - n1 == per_render and n1 - n0 == per_render + 1, (n0, n1, n2)
    # in-place torch update: version counters change -> images re-packed on the next call
    with torch.no_grad():
        st._model.all_fields_params["_linears.0.weight"].mul_(1.05)
    p3 = render(st)
    assert not torch.equal(p3.rgbds, p1.rgbds)
    assert torch.equal(p3.rgbds, render(_fresh_like(meta, a, st)).rgbds)
    # the CUDA Adam step writes through raw pointers: update_step re-packs exactly the rows it touched
    st._reference_flow = True
    pred = st._render_ijs(*_train_batch(st, fid))
    st._update_step({"combined": pred.rgbds.square().mean()}, fid)
    st._reference_flow = False
    p4 = render(st)
    assert not torch.equal(p4.rgbds, p3.rgbds)
    assert torch.equal(p4.rgbds, render(_fresh_like(meta, a, st)).rgbds)
    # checkpoint round trip: new tables, new cache, same pixels
    path = str(tmp_path / "ckpt.pt")
    st.save_model(path)
    st5 = make_state(meta, a, DEV, "fp16")
    st5.load_model(path)
    assert torch.equal(render(st5).rgbds, p4.rgbds)
    assert st5._model._packed_cache[1] is not st._model._packed_cache[1]


def _train_batch(st, fid):
    import neural_graph_mapping_b200 as ngm

    meta, a = G.load("vmap_guided_nrgbd")
    g = torch.Generator().manual_seed(9)
    F, Rr = len(fid), 32
    ijs = torch.stack([torch.randint(0, 480, (F, Rr), generator=g), torch.randint(0, 640, (F, Rr), generator=g)], -1).to(DEV)
    near = (torch.rand(F, Rr, generator=g) * 0.5 + 0.3).to(DEV)
    return ijs, a["c2ws"][0, 0].to(DEV), ngm.Camera(**meta["camera"]), fid, True, near, near + 1.5


def test_pack_rows_matches_full_pack():
    """ngm_pack_weights on a subset of rows writes exactly those rows' images (the others stay byte-identical)."""
    import ctypes as C

    from neural_graph_mapping_b200 import _lib

    meta, a, render, fid = _setup()
    st = make_state(meta, a, DEV, "fp16")
    render(st)
    model = st._model
    full = model._packed_cache[1].clone()
    with torch.no_grad():
        for k, v in model.all_fields_params.items():
            if k.startswith("_linears"):
                v[fid] = v[fid] * 1.5  # index_put_: bumps the version counters -> the cache key no longer matches
    desc, keep = model._prototype_field.field_desc(model.all_fields_params, True)
    per = C.c_size_t(0)
    _lib.check(_lib.lib.ngm_packed_weights_bytes(C.byref(desc), C.byref(per)))
    rows = model.all_fields_params["_linears.0.weight"].shape[0]
    assert full.numel() == rows * per.value
    part = full.clone()
    _lib.check(_lib.lib.ngm_pack_weights(C.byref(desc), fid.data_ptr(), len(fid), part.data_ptr(), _lib.stream_ptr(torch.device(DEV))))
    whole = torch.empty_like(full)
    _lib.check(_lib.lib.ngm_pack_weights(C.byref(desc), None, rows, whole.data_ptr(), _lib.stream_ptr(torch.device(DEV))))
    torch.cuda.synchronize()
    assert torch.equal(part, whole)
    untouched = [r for r in range(rows) if r not in fid.tolist()]
    for r in untouched:
        assert torch.equal(part[r * per.value:(r + 1) * per.value], full[r * per.value:(r + 1) * per.value])
    assert not torch.equal(part, full)
