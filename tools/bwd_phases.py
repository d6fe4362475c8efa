"""Where the tcgen05 backward kernel spends its time: cycles per phase of CTA 0 (issuer lane and one row thread),
config-4 shard, diagnostics build.    python tools/bwd_phases.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NGM_B200_LIB", os.path.join(ROOT, "neural_graph_mapping_b200", "libngm_b200_debug.so"))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0], "--profile", "none"]
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402
from neural_graph_mapping_b200 import _lib  # noqa: E402

dev = "cuda:0"
cam = ngm.Camera(**bench.CAMERA)
F, R, S = 32, 4096, 64
sc = bench.synthetic_scene(7, F, R)
cfg = bench.config_dict(dev, "fp16")
st = ngm.RenderState(cfg)
st.set_fields(sc["params"], sc["positions"], sc["orientations"])
st._reference_flow = True
dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
dbg = _lib.load_debug_lib()
buf = (C.c_uint64 * 16)()
NAMES = {0: "front end", 1: "wait previous tile's dW", 2: "wait forward MMA", 3: "forward epilogue (+arrive)", 4: "wait chain MMA",
         5: "chain epilogue", 6: "wait this step's dW", 7: "store g + arrive", 8: "flush", 9: "spill epilogue (g_lo -> HBM)", 10: "issuer: wait for operands",
         11: "issuer: issue"}
for it in range(3):
    p = st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
    (p.rgbds.square().mean() + p.depth_vars.mean()).backward()
    dbg.ngm_debug_bwd_phases(buf)
tot_row = sum(buf[i] for i in range(10))
tiles = F * R * S // 128 // 148
print(f"CTA 0, {tiles} tiles per launch pass (2 groups x 2 launches); row-thread cycles {tot_row}, per tile {tot_row / tiles:.0f}")
for i, n in NAMES.items():
    base = tot_row if i < 10 else (buf[10] + buf[11])
    print(f"  {n:32s} {buf[i]:12d}  {100.0 * buf[i] / max(base, 1):5.1f}%   per tile {buf[i] / tiles:8.0f}")
