"""BASELINE config 3: the mapping loop on a synthetic RGB-D stream with gt poses, at the reference's default loop
shape (640x480 NRGBD intrinsics, 32 train fields x 512 rays x 8 + 16 samples, 5 iterations per frame,
ngm/config/neural_graph_map.yaml:50-64), for two fields: the reference's default (permutohedral 16 x 2, one hidden
layer of 32) and the BASELINE metric field (NeRF-8, 4 x 128).  Prints one JSON line per run: frames/s of the whole
loop (wall clock, as ngm/run_mapping.py:1249-1251 computes it), the final losses, PSNR / depth L1 of rendered frames.
    python tools/mapping_loop.py [--frames 60] [--precision fp16|fp32]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402
from neural_graph_mapping_b200 import mapping  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--fields", default="permuto_1x32,nerf8_4x128")
args = ap.parse_args()
dev = "cuda:0"
P = "neural_graph_mapping_b200.positional_encodings."
FIELDS = {
    "permuto_1x32": {"encoding_type": P + "PermutohedralEncoding",
                     "encoding_kwargs": {"pos_dim": 3, "log2_hashmap_size": 12, "nr_levels": 16, "nr_feat_per_level": 2,
                                         "coarsest_scale": 1.0, "finest_scale": 0.0001, "init_scale": 0.00001},
                     "num_layers": 1, "dim_out": 4, "dim_mlp_out": None},
    "nerf8_4x128": {"encoding_type": P + "PositionalEncodingNeRF", "encoding_kwargs": {"dim_in": 3, "num_octaves": 8},
                    "num_layers": 4, "dim_out": 4, "dim_mlp_out": 128},
}
for name in args.fields.split(","):
    cfg = bench.config_dict(dev, args.precision)
    fk = dict(FIELDS[name], skip_mode="no", initial_geometry_bias=0.0, neus_initial_sd=1.0)
    cfg["model_kwargs"] = dict(cfg["model_kwargs"], field_kwargs=fk)
    cfg.update(learning_rate=1e-3, adam_eps=1e-15, adam_weight_decay=1e-5, num_train_fields=32, num_rays_per_field=512,
               num_samples_coarse=8, num_samples_depth_guided=16, num_iterations_per_frame=5, eval_num_samples=128,
               termination_weight=0.0, photometric_weight=1.0, depth_weight=1.0, max_keyframes=64)
    cam = ngm.Camera(**bench.CAMERA)
    stream = mapping.SyntheticStream(cam, dev, num_frames=args.frames, keyframe_every=5)
    torch.manual_seed(0)
    loop = mapping.MappingLoop(cfg, cam, stream)
    if name.startswith("nerf"):  # small output head: term_prob ~1 at the start, like the reference's default field
        with torch.no_grad():
            head = loop._model._prototype_field._linears[-1]
            head.weight.mul_(0.05)
            head.bias.mul_(0.05)
    loop._current_frame_optimization(0)  # warm-up frame: field growth, allocator, kernels' first launches
    loop._total_optimization_time = 0.0
    t_frames = list(range(1, args.frames))
    for f in t_frames:
        loop._current_frame_optimization(f)
    fps = len(t_frames) / loop._total_optimization_time
    ev = [loop.evaluate_frame(f) for f in (2, args.frames // 2, args.frames - 3)]
    row = {"workload": "configs[2]: mapping loop, synthetic RGB-D stream (box room + sphere, NRGBD intrinsics 640x480, gt poses, "
                       "keyframe every 5 frames), 5 iterations/frame of 32 fields x 512 rays x 8+16 samples",
           "field": name, "precision": args.precision, "frames": len(t_frames), "fps": round(fps, 2),
           "ms_per_iteration": round(1e3 * loop._total_optimization_time / (5 * len(t_frames)), 3),
           "fields": loop._num_fields, "iterations": loop._current_iteration,
           "psnr_db": [round(e["psnr"], 2) for e in ev], "depth_l1_m": [round(e["depth_l1"], 4) for e in ev]}
    print(json.dumps(row), flush=True)
