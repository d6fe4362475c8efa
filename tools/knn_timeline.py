"""Where a kNN frame's time goes (not under ncu: CUPTI activity records through torch.profiler): per-kernel device
time of one `_render_ijs(use_vmap=False)` frame on the bench scene, and the span of the frame on the GPU."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import bench_variants as bv  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "nerf8_4x128"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
use_vmap = len(sys.argv) > 3 and sys.argv[3] == "vmap"
dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)
c5_fields = int(sys.argv[3][2:]) if len(sys.argv) > 3 and sys.argv[3].startswith("c5") else 0
if c5_fields:  # a map of c5_fields fields on the reference's grid (bench.c4_scene), full frame from inside it, eval samples
    sc = bench.c4_scene(777, num_fields=c5_fields, rays=2)
    cfg = bench.config_dict(dev, prec)
    cfg.update(eval_near_distance=0.5, eval_far_distance=4.5, eval_num_samples=bench.S)
    st = ngm.RenderState(cfg)
    st.set_fields(sc["params"], sc["positions"], sc["orientations"])
    st.eval()
    side = int(c5_fields ** 0.5 + 0.999)
    c2w = torch.eye(4, device=dev)
    c2w[:3, 3] = torch.tensor([side * 0.577, 0.0, 1.0], device=dev)
    ijs = torch.cartesian_prod(torch.arange(bench.H, device=dev), torch.arange(bench.W_IMG, device=dev))
    call = lambda: st._render_ijs(ijs, c2w, cam)  # noqa: E731
else:
    enc, ekw, E, L, W = bv.VARIANTS[variant]
    sc = bv.scene(E, L, W, enc)
    cfg = copy.deepcopy(bench.config_dict(dev, prec))
    cfg["model_kwargs"]["field_kwargs"].update(
        encoding_type=f"neural_graph_mapping_b200.positional_encodings.{enc}", encoding_kwargs=dict(ekw), num_layers=L,
        dim_mlp_out=W)
    st = ngm.RenderState(cfg)
    st.set_fields(sc["params"], sc["positions"], sc["orientations"])
if c5_fields:
    pass
elif use_vmap:
    ijs, near, far = sc["ijs"].to(dev), sc["near"].to(dev), sc["far"].to(dev)
    fid = sc["field_ids"].to(dev)
    call = lambda: st._render_ijs(ijs, c2w, cam, fid, True, near, far)  # noqa: E731
else:
    ijs = sc["ijs"].reshape(-1, 2).to(dev)
    near, far = sc["near"].reshape(-1).to(dev), sc["far"].reshape(-1).to(dev)
    call = lambda: st._render_ijs(ijs, c2w, cam, None, False, near, far)  # noqa: E731
if not c5_fields:
    c2w = sc["c2w"].to(dev)
with torch.no_grad():
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            call()
            torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
frame = evs[-n:]
rows = [{"kernel": e.name[:90], "us": round(e.time_range.elapsed_us(), 1),
         "gap_before_us": round(e.time_range.start - (frame[i - 1].time_range.end if i else e.time_range.start), 1)}
        for i, e in enumerate(frame)]
span = frame[-1].time_range.end - frame[0].time_range.start
print(json.dumps({"variant": variant, "precision": prec, "path": "vmap" if use_vmap else "knn", "map_fields": c5_fields or 75, "launches": n,
                  "span_us": round(span, 1), "kernel_sum_us": round(sum(r["us"] for r in rows), 1), "kernels": rows}, indent=1))
