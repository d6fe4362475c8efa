"""TMEM bandwidth micro-benchmark (diagnostics): bytes per SM clock for tcgen05.ld / st."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

from neural_graph_mapping_b200 import _lib  # noqa: E402

torch.zeros(1, device="cuda")
iters = 2000
for mode, name, bytes_per in [(0, "ld.x32 (wait each)", 32 * 32 * 4), (1, "2x ld.x32 in flight", 2 * 32 * 32 * 4),
                              (3, "ld.x16 (wait each)", 16 * 32 * 4), (2, "st.x16 (wait each)", 16 * 32 * 4)]:
    for warps in (1, 4, 8, 16):
        cyc = C.c_uint64(0)
        rc = _lib.lib.ngm_debug_tmem_bw(warps, iters, mode, C.byref(cyc))
        assert rc == 0, _lib.lib.ngm_last_error()
        total = warps * iters * bytes_per
        print(f"{name:24s} warps={warps:2d}: {cyc.value / iters:8.1f} cyc/iter  {total / cyc.value:8.1f} B/clk/SM")
