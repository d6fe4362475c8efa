"""Phase trace of the fused tcgen05 render kernel (diagnostics).  Run on a GPU box:
    NGM_TC_TRACE=1 python tools/tc_trace.py
Prints, for CTA 0, the cycles between consecutive phase events of each role/slot."""
import collections
import ctypes as C
import os
import sys

os.environ["NGM_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the traced twin of the kernel and the ngm_debug_* entry points live in the diagnostics build only
os.environ.setdefault("NGM_B200_LIB", os.path.join(ROOT, "neural_graph_mapping_b200", "libngm_b200_debug.so"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import neural_graph_mapping_b200 as ngm  # noqa: E402
from neural_graph_mapping_b200 import _lib  # noqa: E402

dev = "cuda:0"
sc = bench.synthetic_scene(1)
st = ngm.RenderState(bench.config_dict(dev, "fp16"))
st.set_fields(sc["params"], sc["positions"], sc["orientations"])
cam = ngm.Camera(**bench.CAMERA)
dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
with torch.no_grad():
    for _ in range(3):
        st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"])
torch.cuda.synchronize()
buf = (C.c_uint64 * 16384)()
n = _lib.load_debug_lib().ngm_debug_tc_trace(buf, 16384)
ev = sorted(((buf[i] & 0xFFFFFFFFFFFF), buf[i] >> 48) for i in range(n))
print("events", n)
PH = {0: "FE start", 1: "sample ready (pre-encode)", 2: "A0 stored+arrived", 3: "d_ready (hidden)",
      4: "hidden epilogue done", 5: "d_ready (last)", 6: "compositor done", 7: "accumulator drained"}
ROLE = {0: "issuer", 1: "front-end half", 2: "compositor half"}


class _N(dict):
    def __missing__(self, k):
        role, ph = k
        if role == 0:
            return "issuer: " + ("a_ready wait done" if ph == 0 else "issued+commit")
        return f"{ROLE.get(role, role)}: {PH.get(ph, ph)}"


names = _N()
last = {}
dur = collections.defaultdict(list)
for clk, e in ev:
    role, slot, phase, layer = e >> 12, (e >> 8) & 1, (e >> 4) & 15, e & 15
    key = (role, slot)
    if key in last:
        pclk, pphase, player = last[key]
        dur[(role, slot, pphase, player, phase, layer)].append(clk - pclk)
    last[key] = (clk, phase, layer)
for k in sorted(dur):
    role, slot, p0, l0, p1, l1 = k
    v = sorted(dur[k])
    med = v[len(v) // 2]
    print(f"role {role} slot {slot}: [{names[(role, p0)]} L{l0}] -> [{names[(role, p1)]} L{l1}]  n={len(v):4d} "
          f"median {med:7d}  p10 {v[len(v) // 10]:7d}  p90 {v[len(v) * 9 // 10]:7d}")
t0, t1 = ev[0][0], ev[-1][0]
tiles = sum(1 for _, e in ev if (e >> 12) == 2 and ((e >> 4) & 15) == 6)
print(f"CTA 0: {tiles} tiles in {t1 - t0} cycles -> {(t1 - t0) / max(tiles, 1):.0f} cycles/tile (both slots)")
# raw timeline of a window in the middle of the trace (relative cycles), for reading the overlap by eye
mid = len(ev) // 2
base = ev[mid][0] if ev else 0
print("-- timeline (cycle, role/slot, phase, layer)")
for clk, e in ev[mid:mid + 260]:
    role, slot, phase, layer = e >> 12, (e >> 8) & 1, (e >> 4) & 15, e & 15
    print(f"{clk - base:8d}  {'  ' * (2 * slot + (role != 0) + (role == 1))}{'I' if role == 0 else ('F' if role == 1 else 'C')}{slot} "
          f"{names[(role, phase)].split(': ')[-1]} L{layer}")
