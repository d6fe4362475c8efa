"""Helpers shared by the GPU parity tests, smoke() and bench.py: build a stand-alone RenderState
from a golden fixture's stored reference config and run the CUDA path on it."""
import torch

import golden_util as G

_ENC_CLS = {
    "PositionalEncodingNeRF": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
    "PositionalEncodingFourier": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingFourier",
    "TriplaneEncoding": "neural_graph_mapping_b200.positional_encodings.TriplaneEncoding",
    "PermutohedralEncoding": "neural_graph_mapping_b200.positional_encodings.PermutohedralEncoding",
}


def product_field_kwargs(field_kwargs):
    """Reference YAML field_kwargs -> the same kwargs with this package's type strings."""
    fk = dict(field_kwargs)
    fk["encoding_type"] = _ENC_CLS[fk["encoding_type"].split(".")[-1]]
    return fk


def product_config(meta, device, precision="fp32"):
    cfg = dict(meta["config"])
    mk = dict(cfg["model_kwargs"])
    mk["field_type"] = "neural_graph_mapping_b200.models.NeuralField"
    mk["field_kwargs"] = product_field_kwargs(meta["field_kwargs"])
    cfg["model_kwargs"] = mk
    cfg["model_type"] = "neural_graph_mapping_b200.models.NeuralFieldSet"
    cfg["device"] = device
    cfg["precision"] = precision
    cfg["num_samples_coarse"] = meta["num_samples"]
    cfg["num_samples_depth_guided"] = meta.get("num_samples_depth_guided", 0)
    if "near_distance" in meta:
        cfg["eval_near_distance"] = meta["near_distance"]
        cfg["eval_far_distance"] = meta["far_distance"]
        cfg["eval_num_samples"] = meta["num_samples"]
    return cfg


def make_state(meta, arrays, device, precision="fp32"):
    import neural_graph_mapping_b200 as ngm

    st = ngm.RenderState(product_config(meta, device, precision))
    st.set_fields(G.params(arrays), arrays["positions"], arrays["orientations"])
    return st


def run_vmap_case(meta, a, device, precision="fp32"):
    import neural_graph_mapping_b200 as ngm

    st = make_state(meta, a, device, precision)
    cam = ngm.Camera(**meta["camera"])
    d = lambda k: a[k].to(device) if k in a else None  # noqa: E731
    with torch.no_grad():
        return st._render_ijs(d("ijs"), d("c2ws"), cam, d("field_ids"), True, d("near"), d("far"), d("gt"),
                              jitter=d("jitter"), jitter_guided=d("jitter_guided"))
