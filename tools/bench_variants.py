"""Secondary measurements (not the bench.py headline): the 640x480x64 keyframe render of bench.py with other
field configurations -- SURVEY.md 8d "C2 secondary": permutohedral encoding with the 4x128 MLP, and the
reference's default field (permutohedral 16x2, one hidden layer of 32: neural_graph_map.yaml:6-17).
    python tools/bench_variants.py            # prints one line per (variant, precision)
"""
import copy
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

PERMUTO = {"pos_dim": 3, "log2_hashmap_size": 12, "nr_levels": 16, "nr_feat_per_level": 2, "coarsest_scale": 1.0,
           "finest_scale": 1e-4, "init_scale": 0.5}
VARIANTS = {
    "nerf8_4x128": ("PositionalEncodingNeRF", {"dim_in": 3, "num_octaves": 8}, 48, 4, 128),
    "permuto_4x128": ("PermutohedralEncoding", PERMUTO, 32, 4, 128),
    "permuto_1x32 (reference default field)": ("PermutohedralEncoding", PERMUTO, 32, 1, 32),
}


def scene(E, L, W, enc_name, seed=1234):
    sc = bench.synthetic_scene(seed)
    g = torch.Generator().manual_seed(seed + 1)
    F = bench.F_FIELDS
    params = {}
    dims_in, dims_out = [E] + [W] * L, [W] * L + [4]
    for i, (di, do) in enumerate(zip(dims_in, dims_out)):
        b = 1.0 / math.sqrt(di)
        params[f"_linears.{i}.weight"] = (torch.rand(F, do, di, generator=g) * 2 - 1) * b
        params[f"_linears.{i}.bias"] = (torch.rand(F, do, generator=g) * 2 - 1) * b
    params[f"_linears.{L}.bias"][:, 3] += 0.3
    params["_neus_sd"] = torch.ones(F)
    if enc_name == "PermutohedralEncoding":
        params["_encoding.lattice_values"] = (torch.rand(F, 16, 4096, 2, generator=g) * 2 - 1) * 0.5
        params["_encoding.random_shift_per_level"] = torch.randn(F, 16, 3, generator=g) * 10.0
    sc["params"] = params
    return sc


def main():
    dev = "cuda:0"
    cam = ngm.Camera(**bench.CAMERA)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (enc, ekw, E, L, W) in VARIANTS.items():
        sc = scene(E, L, W, enc)
        dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
        for prec in ("fp16", "fp32"):
            if prec == "fp32" and W == 128:
                continue  # 200 ms/frame on the FFMA path: measured in round 1, not repeated here
            cfg = copy.deepcopy(bench.config_dict(dev, prec))
            fk = cfg["model_kwargs"]["field_kwargs"]
            fk.update(encoding_type=f"neural_graph_mapping_b200.positional_encodings.{enc}", encoding_kwargs=dict(ekw),
                      num_layers=L, dim_mlp_out=W)
            st = ngm.RenderState(cfg)
            st.set_fields(sc["params"], sc["positions"], sc["orientations"])
            ts = []
            with torch.no_grad():
                for i in range(13):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            rays = bench.F_FIELDS * bench.R_RAYS
            flops = 2 * (E * W + (L - 1) * W * W + W * 4) * rays * bench.S
            print(json.dumps({"variant": name, "precision": prec, "ms_per_frame": round(ms, 4),
                              "rays_per_s": round(rays / ms * 1e3), "mlp_tflops": round(flops / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
