"""One launch of every kernel on the bench workload inside a cudaProfilerStart/Stop range.
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/prof python tools/prof_stages.py
(B200_PROFILING.md; numbers printed under ncu are never bench values.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402
from neural_graph_mapping_b200 import models, renderer  # noqa: E402
from neural_graph_mapping_b200.camera import sample_rays  # noqa: E402

dev = "cuda:0"
S, F, R = bench.S, bench.F_FIELDS, bench.R_RAYS
sc = bench.synthetic_scene(1234)
st = ngm.RenderState(bench.config_dict(dev, "fp16"))
st.set_fields(sc["params"], sc["positions"], sc["orientations"])
cam = ngm.Camera(**bench.CAMERA)
dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
model = st._model
pos, ori = st._global_map_dict["positions"], st._global_map_dict["orientations"]


def once():
    with torch.no_grad():
        _, dist, world_pts, depth = sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], c2ws=dz["c2w"], seed=1,
                                                want_world=True, want_depth=True)
        q = world_pts.view(F, R * S, 3)
        outs = models.field_forward(model._prototype_field, model.all_fields_params, True, q, pos, ori, dz["field_ids"],
                                    model._scale_mode, model._field_radius, "fp16")
        o = outs.view(F * R, S, 4)
        renderer.composite(o, o[..., 3], dist.view(F * R, S), depth.view(F * R, S), "nrgbd", 20.0, color_stride=4,
                           geometry_stride=4)
        st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
    torch.cuda.synchronize()


for _ in range(3):
    once()
torch.cuda.profiler.start()
once()
torch.cuda.profiler.stop()
print("profiled one launch of each kernel")
