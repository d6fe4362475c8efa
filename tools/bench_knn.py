"""Secondary measurement: the driver's eval path -- render_image / _render_ijs(use_vmap=False), K=2 blend over all
fields (ngm/run_mapping.py:402-437, models.py:347-405) -- on the bench.py scene: 640x480 pixels x 64 samples,
75 fields, NeRF-8 + 4x128 MLP.  fp32 = FFMA field kernel, fp16 = tcgen05 field kernel in gather mode."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

import copy  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_variants as bv  # noqa: E402

dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)
runs = [("nerf8_4x128", "fp16", 5), ("nerf8_4x128", "fp32", 1),
        ("permuto_1x32 (reference default field)", "fp16", 3), ("permuto_1x32 (reference default field)", "fp32", 3)]
if "--only" in sys.argv:  # e.g. --only nerf8_4x128:fp16 (profiling one configuration under ncu)
    v_, p_ = sys.argv[sys.argv.index("--only") + 1].split(":")
    runs = [r for r in runs if r[0].startswith(v_) and r[1] == p_]
for variant, prec, reps in runs:
    enc, ekw, E, L, W = bv.VARIANTS[variant]
    sc = bv.scene(E, L, W, enc)
    ijs = sc["ijs"].reshape(-1, 2).to(dev)
    near, far = sc["near"].reshape(-1).to(dev), sc["far"].reshape(-1).to(dev)
    c2w = sc["c2w"].to(dev)
    cfg = copy.deepcopy(bench.config_dict(dev, prec))
    cfg["model_kwargs"]["field_kwargs"].update(
        encoding_type=f"neural_graph_mapping_b200.positional_encodings.{enc}", encoding_kwargs=dict(ekw), num_layers=L,
        dim_mlp_out=W)
    st = ngm.RenderState(cfg)
    st.set_fields(sc["params"], sc["positions"], sc["orientations"])
    ts = []
    with torch.no_grad():
        for i in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p = st._render_ijs(ijs, c2w, cam, None, False, near, far)
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                ts.append(e0.elapsed_time(e1))
    ms_sync = sum(ts) / len(ts)
    # frames back to back (what a render loop does): the host prepares frame i+1 while the GPU runs frame i
    with torch.no_grad():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            p = st._render_ijs(ijs, c2w, cam, None, False, near, far)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"path": "kNN (use_vmap=False), K=2, 75 fields, 307200 rays x 64", "field": variant, "precision": prec,
                      "ms_per_frame": round(ms, 3), "ms_per_frame_host_synced": round(ms_sync, 3),
                      "rays_per_s": round(307200 / ms * 1e3),
                      "inside_fraction": round(float((p.term_probs > 0).float().mean()), 3)}), flush=True)


if "--only" in sys.argv:
    sys.exit(0)
# render_image (the driver's call, run_mapping.py:402-437): whole 640x480 frame, reference block size vs ours
from neural_graph_mapping_b200 import renderer  # noqa: E402

enc, ekw, E, L, W = bv.VARIANTS["nerf8_4x128"]
sc = bv.scene(E, L, W, enc)
cfg = copy.deepcopy(bench.config_dict(dev, "fp16"))
cfg["eval_near_distance"], cfg["eval_far_distance"], cfg["eval_num_samples"] = 1.0, 3.0, 64
st = ngm.RenderState(cfg)
st.set_fields(sc["params"], sc["positions"], sc["orientations"])
st.eval()
for blk in (0, renderer.IMAGE_BLOCK_BYTES):  # 0: exactly the reference's 8,192-pixel blocks
    renderer.IMAGE_BLOCK_BYTES = blk
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img, dv = st.render_image(sc["c2w"].to(dev), cam)
        e1.record()
        torch.cuda.synchronize()
        if i > 0:
            ts.append(e0.elapsed_time(e1))
    print(json.dumps({"call": "render_image 640x480, eval samples", "block_budget_bytes": blk, "block_pixels": renderer.image_block_pixels(st),
                      "ms_per_frame": round(sum(ts) / len(ts), 3)}), flush=True)
