"""Checkpoint format: the file the unmodified reference's save_model wrote
(tests/golden/checkpoint_ref.pt, made by oracle/make_checkpoint_fixture.py; ngm/run_mapping.py:2147-2173) loads into
this package and renders the reference's outputs; a file this package writes has the same keys, names and shapes."""
import os

import pytest
import torch

import golden_util as G
from tests_support import product_config

CKPT = os.path.join(G.GOLDEN_DIR, "checkpoint_ref.pt")


def test_reference_checkpoint_layout_and_roundtrip(tmp_path):
    """CPU: key set / names / shapes of the reference's file, and our save_model writes the same layout."""
    import neural_graph_mapping_b200 as ngm

    meta, _ = G.load("checkpoint_render")
    ref = torch.load(CKPT, map_location="cpu")
    assert set(ref) == {"map_dict", "all_fields_params", "state_dict"}
    n = meta["num_fields"]
    assert ref["map_dict"]["num"] == n and ref["map_dict"]["positions"].shape == (32, 3)  # over-allocated tables
    st = ngm.RenderState(product_config(meta, "cpu"))
    st.load_model(CKPT)
    assert st._global_map_dict["num"] == n
    assert set(st._model.all_fields_params) == set(ref["all_fields_params"])
    assert all(v.shape[0] == n for v in st._model.all_fields_params.values())
    # state-dict names are the reference's (prototype field under `_prototype_field.`), strict load succeeded
    assert set(st._model.state_dict()) == set(ref["state_dict"])
    out = str(tmp_path / "ours.pt")
    st.save_model(out)
    mine = torch.load(out, map_location="cpu")
    assert set(mine) == set(ref)
    for k, v in ref["all_fields_params"].items():
        assert torch.equal(mine["all_fields_params"][k], v)
    for k, v in ref["state_dict"].items():
        assert torch.equal(mine["state_dict"][k], v)
    for k, v in ref["map_dict"].items():
        assert torch.equal(torch.as_tensor(mine["map_dict"][k]), torch.as_tensor(v)), k


def test_our_checkpoint_loads_into_live_reference(tmp_path):
    """Where /root/reference exists: the reference's own load_model accepts a file written by this package."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present")
    import neural_graph_mapping_b200 as ngm

    meta, _ = G.load("checkpoint_render")
    st = ngm.RenderState(product_config(meta, "cpu"))
    st.load_model(CKPT)
    out = str(tmp_path / "ours.pt")
    st.save_model(out)
    ref = ref_loader.load()
    cfg = ref_loader.default_config(model_kwargs={"field_kwargs": meta["field_kwargs"]})
    m = ref.run_mapping.NeuralGraphMap(cfg)
    m.load_model(out)
    assert m._global_map_dict["num"] == meta["num_fields"]
    assert torch.equal(m._model.all_fields_params["_linears.0.weight"], st._model.all_fields_params["_linears.0.weight"])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_render_from_reference_checkpoint(precision):
    """GPU: load the reference's checkpoint, render the per-field and the kNN batch, compare with the reference's
    renders of the same state (stale rows of the over-allocated map tables must not be read)."""
    import neural_graph_mapping_b200 as ngm

    dev = "cuda:0"
    meta, a = G.load("checkpoint_render")
    st = ngm.RenderState(product_config(meta, dev, precision))
    st.load_model(CKPT)
    cam = ngm.Camera(**meta["camera"])
    d = lambda k: a[k].to(dev)  # noqa: E731
    tol = 3e-5 if precision == "fp32" else 6e-3
    with torch.no_grad():
        p = st._render_ijs(d("ijs"), d("c2ws"), cam, d("field_ids"), True, d("near"), d("far"), jitter=d("jitter"))
        st.eval()
        k = st._render_ijs(d("knn_ijs"), d("knn_c2w"), cam, jitter=d("knn_jitter"))
    for ours, ref, what in ((p.rgbds, a["out_rgbds"], "vmap rgbd"), (p.term_probs, a["out_term_probs"], "vmap term"),
                            (k.rgbds, a["knn_out_rgbds"], "knn rgbd"), (k.term_probs, a["knn_out_term_probs"], "knn term")):
        err = (ours.cpu() - ref).abs()
        if precision == "fp32":
            assert err.max().item() < 1e-4 + tol * ref.abs().max().item(), (what, err.max().item())
        else:
            assert err.mean().item() < tol, (what, err.mean().item())
