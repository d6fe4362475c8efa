"""Secondary measurement: one training render step (forward + backward into the field parameters) of the
differentiable path (autograd.py), at the reference's default training shape (32 fields x 512 rays x 8+16
samples, neural_graph_map.yaml:60-63) and at BASELINE config 4's per-GPU shard (32 fields x 4096 rays x 64
samples), 4-layer x 128 MLP + NeRF-8.  The CPU column is the oracle port under torch.autograd on the host cores."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)


def run(F, R, S, G, prec, reps=5):
    sc = bench.synthetic_scene(7, F, R)
    cfg = bench.config_dict(dev, prec)
    cfg["num_samples_coarse"], cfg["num_samples_depth_guided"] = S, G
    st = ngm.RenderState(cfg)
    st.set_fields(sc["params"], sc["positions"], sc["orientations"])
    st._reference_flow = True  # the driver's flow: _render_ijs first gathers the active fields into leaves (:500)
    dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
    gt = (dz["near"] + dz["far"]) * 0.5 if G else None
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p = st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"], gt)
        loss = p.rgbds.square().mean() + p.depth_vars.mean() + p.term_probs.mean()
        if p.freespace_geometry is not None:
            loss = loss + p.freespace_geometry.square().mean() + p.tsdf_residuals.square().mean()
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def run_cpu(F, R, S, G):
    from oracle import restatement as Rr

    sc = bench.synthetic_scene(7, F, R)
    params = {k: v.clone().requires_grad_(True) for k, v in sc["params"].items()}
    fs = Rr.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 8}, bench.L_MLP, 4, bench.W_MLP, "no")
    rs = Rr.RenderSpec(num_samples=S, num_samples_depth_guided=G, geometry_mode="nrgbd", geometry_factor=20.0,
                       truncation_distance=0.1, range_depth_guided=0.1, freespace_weight=40.0, tsdf_weight=50.0)
    cs = Rr.CameraSpec(**bench.CAMERA)
    g = torch.Generator().manual_seed(3)
    gt = (sc["near"] + sc["far"]) * 0.5 if G else None
    t0 = time.perf_counter()
    p = Rr.render_rays(sc["ijs"], sc["c2w"], cs, rs, fs, params, sc["positions"], sc["orientations"],
                       field_ids=sc["field_ids"], use_vmap=True, near_distances=sc["near"], far_distances=sc["far"],
                       gt_distances=gt, jitter=torch.rand(F, R, S, generator=g),
                       jitter_guided=torch.rand(F, R, max(G, 1), generator=g)[..., :G] if G else None)
    loss = p.rgbds.square().mean() + p.depth_vars.mean() + p.term_probs.mean()
    loss.backward()
    return (time.perf_counter() - t0) * 1e3


SHAPES = {"default training batch (32 x 512 x 8+16)": (32, 512, 8, 16),
          "config-4 shard (32 x 4096 x 64)": (32, 4096, 64, 0)}
if "--profile" in sys.argv:  # under ncu: python tools/bench_train.py --profile {default|c4}: one fp16 step, no CPU column
    which = sys.argv[sys.argv.index("--profile") + 1]
    F, R, S, G = SHAPES[[k for k in SHAPES if (which == "c4") == k.startswith("config-4")][0]]
    print(json.dumps({"profiled": which, "ms": run(F, R, S, G, "fp16", reps=1)}))
    sys.exit(0)
for name, (F, R, S, G) in SHAPES.items():
    row = {"shape": name, "rays": F * R, "points": F * R * (S + G)}
    for prec in ("fp16", "fp32"):  # fp16: tcgen05 forward + tcgen05 backward; fp32: reference arithmetic
        ms = run(F, R, S, G, prec)
        row[f"gpu_{prec}_ms"] = round(ms, 3)
        row[f"gpu_{prec}_rays_per_s"] = round(F * R / ms * 1e3)
    if F * R * (S + G) <= 500000:
        torch.set_num_threads(os.cpu_count() or 1)
        row["cpu_oracle_ms"] = round(min(run_cpu(F, R, S, G) for _ in range(2)), 1)
        row["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(row), flush=True)
