"""Where the single fused render kernel pays off: frame time (CUDA events, frames back to back) of the fused kernel
against the stage kernels (sampler -> [row encoder] -> tcgen05 field kernel -> compositor, NGM_RENDER_FUSED=0) over MLP
sizes, for the NeRF-8 and the permutohedral encoding.  One JSON line per configuration."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench  # noqa: E402
import bench_variants as bv  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)
ENC = {"nerf8": ("PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 8}, 48),
       "permuto": ("PermutohedralEncoding", bv.PERMUTO, 32)}
for enc_key in ("nerf8", "permuto"):
    enc, ekw, E = ENC[enc_key]
    for L, W in ((1, 32), (1, 64), (2, 64), (4, 64), (1, 128), (2, 128), (4, 128)):
        sc = bv.scene(E, L, W, enc)
        dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
        cfg = copy.deepcopy(bench.config_dict(dev, "fp16"))
        cfg["model_kwargs"]["field_kwargs"].update(
            encoding_type=f"neural_graph_mapping_b200.positional_encodings.{enc}", encoding_kwargs=dict(ekw), num_layers=L,
            dim_mlp_out=W)
        st = ngm.RenderState(cfg)
        st.set_fields(sc["params"], sc["positions"], sc["orientations"])
        res = {}
        for mode in ("1", "0"):
            os.environ["NGM_RENDER_FUSED"] = mode
            with torch.no_grad():
                for _ in range(3):
                    st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(8):
                    st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
                e1.record()
                torch.cuda.synchronize()
            res["fused_ms" if mode == "1" else "staged_ms"] = round(e0.elapsed_time(e1) / 8, 3)
        macs = E * W + (L - 1) * W * W + W * 4
        print(json.dumps({"encoding": enc_key, "layers": L, "width": W, "macs_per_point": macs, **res}), flush=True)
os.environ.pop("NGM_RENDER_FUSED", None)
