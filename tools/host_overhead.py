"""Host-side cost of the public calls (wall clock with a device synchronise around N calls, and cProfile of the
callers): the default training-batch render under no_grad (32 fields x 512 rays x 24 samples) and one kNN frame."""
import cProfile
import copy
import io
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench  # noqa: E402
import bench_variants as bv  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402

dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)
enc, ekw, E, L, W = bv.VARIANTS["nerf8_4x128"]
sc = bv.scene(E, L, W, enc)
cfg = copy.deepcopy(bench.config_dict(dev, "fp16"))
st = ngm.RenderState(cfg)
st.set_fields(sc["params"], sc["positions"], sc["orientations"])
F, R = 32, 512
ijs = sc["ijs"][:F, :R].to(dev).contiguous()
near, far = sc["near"][:F, :R].to(dev).contiguous(), sc["far"][:F, :R].to(dev).contiguous()
fid = sc["field_ids"][:F].to(dev)
c2w = sc["c2w"].to(dev)
ij_all = sc["ijs"].reshape(-1, 2).to(dev)
n_all, f_all = sc["near"].reshape(-1).to(dev), sc["far"].reshape(-1).to(dev)


def small():
    return st._render_ijs(ijs, c2w, cam, fid, True, near, far)


def knn_frame():
    return st._render_ijs(ij_all, c2w, cam, None, False, n_all, f_all)


out = {}
with torch.no_grad():
    for name, fn, n in (("default_batch_render", small, 300), ("knn_frame", knn_frame, 10)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t_issue = (time.perf_counter() - t0) / n
        torch.cuda.synchronize()
        t_total = (time.perf_counter() - t0) / n
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(n):
            fn()
            torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
        out[name] = {"host_issue_us_per_call": round(t_issue * 1e6, 1), "wall_us_per_call_back_to_back": round(t_total * 1e6, 1)}
        print(name, out[name])
        print("\n".join(s.getvalue().splitlines()[:40]))
print(json.dumps(out))
