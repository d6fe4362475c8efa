#!/usr/bin/env python
"""bench.py -- rays/sec of the per-field ray renderer (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] -- one 640x480 keyframe render, 64 samples/ray,
4-layer x 128 MLP, NeRF-8 encoding (E=48): the 307,200 pixels pre-bucketed to 75 posed fields
x 4,096 rays (SURVEY.md 8d, C2 primary).  One *step* = one full keyframe (307,200 rays,
19,660,800 sample points).  N>1 (weak scaling): every rank renders one keyframe of its own and
the rendered tiles (36 B/ray) are all-gathered over NCCL inside the step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision auto|fp16|fp32]
    python bench.py --impl reference ...      # the CPU arm: oracle port of the reference path

One JSON line on stdout (rank 0).  `value` = whole-job rays/s with inputs resident in HBM,
`e2e` = same metric through the public Python API with HOST (pinned) buffers, H2D of the rays
and D2H of the rendered Prediction inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "rays/sec at 64 samples/ray, 4-layer x 128 MLP"
H, W_IMG, S = 480, 640, 64
F_FIELDS, R_RAYS = 75, 4096  # 75 * 4096 = 307,200 = 640 * 480
E_ENC, W_MLP, L_MLP = 48, 128, 4
FLOPS_PER_POINT = 2 * (E_ENC * W_MLP + (L_MLP - 1) * W_MLP * W_MLP + W_MLP * 4)
SAMPLER_BYTES_PER_RAY = 24 + 20 * S
COMPOSITE_BYTES_PER_RAY = S * (4 * 4 + 8) + 36  # fp32 MLP output
FUSED_BYTES_PER_RAY = 60

FIELD_KWARGS = {
    "encoding_type": "neural_graph_mapping_b200.positional_encodings.PositionalEncodingNeRF",
    "encoding_kwargs": {"dim_in": 3, "num_octaves": 8},
    "num_layers": L_MLP, "dim_out": 4, "dim_mlp_out": W_MLP, "skip_mode": "no",
    "initial_geometry_bias": 0.0, "neus_initial_sd": 1.0,
}
CAMERA = dict(width=W_IMG, height=H, fx=554.2562584220408, fy=554.2562584220408, cx=319.5, cy=239.5,
              pixel_center=0.0)


def config_dict(device, precision):
    return {
        "model_type": "neural_graph_mapping_b200.models.NeuralFieldSet",
        "model_kwargs": {
            "dim_points": 3, "field_type": "neural_graph_mapping_b200.models.NeuralField",
            "field_kwargs": FIELD_KWARGS, "num_knn": 2, "distance_factor": 10.0, "field_radius": 1.0,
            "scale_mode": "unit_cube", "outside_value": 1.0,
        },
        "color_factor": 1.0, "geometry_factor": 20.0, "device": device, "field_radius": 1.0,
        "freespace_weight": 40.0, "tsdf_weight": 50.0, "near_distance": 0.0, "far_distance": 8.0,
        "pixel_block_size": 8192, "block_size": 3000000, "geometry_mode": "nrgbd",
        "truncation_distance": 0.1, "num_samples_coarse": S, "num_samples_depth_guided": 0,
        "single_field_id": None, "precision": precision,
    }


def synthetic_scene(seed, num_fields=F_FIELDS, rays=R_RAYS):
    """Seeded synthetic keyframe: CPU tensors (ijs, near, far, c2w, stacked params, poses)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    n = num_fields * rays
    pix = torch.arange(n) % (H * W_IMG)
    ijs = torch.stack([pix // W_IMG, pix % W_IMG], -1).view(num_fields, rays, 2)
    c2w = torch.eye(4)
    # field f sits on the central ray of its pixel chunk at distance 2; radius 1
    cx0, cy0, fx = CAMERA["cx"], CAMERA["cy"], CAMERA["fx"]
    mid = ijs[:, rays // 2].float()
    d = torch.stack([(mid[:, 1] - cx0) / fx, -(mid[:, 0] - cy0) / fx, -torch.ones(num_fields)], -1)
    d = d / d.norm(dim=-1, keepdim=True)
    positions = d * 2.0
    q = torch.randn(num_fields, 4, generator=g)
    orientations = q / q.norm(dim=-1, keepdim=True)
    near = torch.full((num_fields, rays), 1.0) + 0.05 * torch.rand(num_fields, rays, generator=g)
    far = torch.full((num_fields, rays), 3.0) - 0.05 * torch.rand(num_fields, rays, generator=g)
    dims_in = [E_ENC] + [W_MLP] * L_MLP
    dims_out = [W_MLP] * L_MLP + [4]
    params = {}
    for i, (di, do) in enumerate(zip(dims_in, dims_out)):
        b = 1.0 / math.sqrt(di)
        params[f"_linears.{i}.weight"] = (torch.rand(num_fields, do, di, generator=g) * 2 - 1) * b
        params[f"_linears.{i}.bias"] = (torch.rand(num_fields, do, generator=g) * 2 - 1) * b
    params[f"_linears.{L_MLP}.bias"][:, 3] += 0.3  # keep occupancies non-degenerate
    params["_neus_sd"] = torch.ones(num_fields)
    return dict(ijs=ijs, c2w=c2w, near=near, far=far, positions=positions, orientations=orientations,
                params=params, field_ids=torch.arange(num_fields))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Start nvidia-smi early (it needs ~0.5 s to produce its first line); mark() brackets the timed region."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        ok = [(t, r) for t, r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        # a line printed at time t reports the clocks of the preceding sampling period
        inside = [r for t, r in ok if self.t0 is not None and self.t0 <= t <= self.t1 + 0.03]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the samples of the whole GPU-busy run (warm-up included)
            inside, window = [r for _, r in ok], "whole run (timed region shorter than one nvidia-smi period)"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------
_BEST_THREADS = None


def _best_thread_count():
    """The reference path is many small GEMMs + elementwise ops; past a few dozen threads torch's
    intra-op pool gets slower, not faster.  Probe {all, 64, 32, 16, 8} host threads on a small sample
    and keep the fastest, so the CPU arm gets the best configuration it can use on this host."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    import torch

    from oracle import restatement as Rr

    total = os.cpu_count() or 1
    cands = sorted({c for c in (total, 64, 32, 16, 8) if c <= total}, reverse=True)
    sc = synthetic_scene(3, 2, 512)
    fs = Rr.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 8}, L_MLP, 4, W_MLP, "no")
    rs = Rr.RenderSpec(num_samples=S, geometry_mode="nrgbd", geometry_factor=20.0)
    cam = Rr.CameraSpec(**CAMERA)
    jit = torch.rand(2, 512, S, generator=torch.Generator().manual_seed(3))
    best, best_t = total, float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                Rr.render_rays(sc["ijs"], sc["c2w"], cam, rs, fs, sc["params"], sc["positions"], sc["orientations"],
                               field_ids=sc["field_ids"], use_vmap=True, near_distances=sc["near"],
                               far_distances=sc["far"], jitter=jit)
                ts.append(time.perf_counter() - t0)
            if min(ts[1:]) < best_t:
                best, best_t = c, min(ts[1:])
    _BEST_THREADS = best
    return best


def cpu_reference_rate(num_fields, rays, repeats, seed=7):
    """rays/s of oracle.restatement.render_rays (vmap path, fp32, all host threads) on a bounded
    sample of the SAME workload (same MLP / encoding / samples per ray)."""
    import torch

    from oracle import restatement as Rr

    sc = synthetic_scene(seed, num_fields, rays)
    cores = _best_thread_count()
    torch.set_num_threads(cores)
    fs = Rr.FieldSpec("nerf", {"dim_in": 3, "num_octaves": 8}, L_MLP, 4, W_MLP, "no")
    rs = Rr.RenderSpec(num_samples=S, geometry_mode="nrgbd", geometry_factor=20.0)
    cam = Rr.CameraSpec(**CAMERA)
    g = torch.Generator().manual_seed(seed)
    jit = torch.rand(num_fields, rays, S, generator=g)
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats + 1):  # first = warm-up
            t0 = time.perf_counter()
            Rr.render_rays(sc["ijs"], sc["c2w"], cam, rs, fs, sc["params"], sc["positions"], sc["orientations"],
                           field_ids=sc["field_ids"], use_vmap=True, near_distances=sc["near"],
                           far_distances=sc["far"], jitter=jit)
            dt = time.perf_counter() - t0
            if _ > 0:
                best = min(best, dt)
    return num_fields * rays / best, cores, f"{num_fields} fields x {rays} rays x {S} samples, best of {repeats}"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    nf, rr = 4, 2048  # 8,192 rays = 524,288 points per step (~0.5 s on 8 cores)
    rate = None
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_reference_rate(nf, rr, repeats=1, seed=7 + i)
        if i >= args.warmup:
            times.append(nf * rr / r)
    ms = 1e3 * sum(times) / len(times)
    rate = nf * rr / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 640x480 keyframe render, 64 samples/ray, 4-layer x 128 MLP, NeRF-8 "
                               f"(per-field path); each step a bounded sample of {nf}x{rr} rays"},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{nf} fields x {rr} rays x {S} samples per step (oracle/restatement.py, torch CPU fp32)"},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    import __graft_entry__

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    clocks = ClockSampler(local_rank)
    clocks.start()
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    __graft_entry__.build()
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import _lib, distributed

    # ---- precision: fp16 tensor-core path when the build has it, else the fp32 path ----
    precision = args.precision
    sc = synthetic_scene(1234 + rank)
    cam = ngm.Camera(**CAMERA)

    def make_state(prec):
        st = ngm.RenderState(config_dict(dev, prec))
        st.set_fields(sc["params"], sc["positions"], sc["orientations"])
        return st

    dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
    if precision == "auto":
        try:
            st = make_state("fp16")
            with torch.no_grad():
                st._render_ijs(dz["ijs"][:1, :128], dz["c2w"], cam, dz["field_ids"][:1], True, dz["near"][:1, :128],
                               dz["far"][:1, :128])
            torch.cuda.synchronize()
            precision = "fp16"
        except NotImplementedError:
            precision = "fp32"
    st = make_state(precision)
    rays_per_step = F_FIELDS * R_RAYS
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident():
        with torch.no_grad():
            return distributed.render_rays_gathered(st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"], dz["far"])

    # pinned host buffers for the end-to-end arm
    hz = {k: sc[k].pin_memory() for k in ("ijs", "near", "far", "c2w")}
    h_out = torch.empty(world, rays_per_step, 9, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(hz[k].numel() * hz[k].element_size() for k in hz)
    d2h_bytes = h_out.numel() * 4

    def step_e2e():
        with torch.no_grad():
            d = {k: hz[k].to(dev, non_blocking=True) for k in hz}
            packed = distributed.render_rays_gathered(st, d["ijs"], d["c2w"], cam, dz["field_ids"], d["near"], d["far"],
                                                      return_packed=True)
            h_out.copy_(packed.view(world, rays_per_step, 9), non_blocking=True)

    def timed_e2e_two_streams(steps, warmup):
        """End to end through the public API with HOST buffers, double-buffered over two streams (what a serving loop
        does): every step still copies its inputs from pinned host memory and its result back inside the timed region."""
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        h_outs = [h_out, torch.empty_like(h_out).pin_memory()]

        def one(i):
            with torch.cuda.stream(streams[i % 2]), torch.no_grad():
                d = {k: hz[k].to(dev, non_blocking=True) for k in hz}
                flush.fill_(1)
                packed = distributed.render_rays_gathered(st, d["ijs"], d["c2w"], cam, dz["field_ids"], d["near"], d["far"],
                                                          return_packed=True)
                h_outs[i % 2].copy_(packed.view(world, rays_per_step, 9), non_blocking=True)

        for i in range(warmup):
            one(i)
        torch.cuda.synchronize()
        e0, e1, em = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(streams[0])
        streams[1].wait_event(e0)
        for i in range(steps):
            one(i)
        em.record(streams[1])
        streams[0].wait_event(em)
        e1.record(streams[0])
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, clk=None):
        for _ in range(warmup):
            fn()
            flush.fill_(1)
        barrier()
        if clk:
            clk.mark_begin()
        evs = []
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the per-step events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        if clk:
            clk.mark_end()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

    launches0 = _lib.lib.ngm_launch_count()
    total_ms = timed(step_resident, args.steps, args.warmup, clocks)
    clock_info = clocks.stop()
    launches = (_lib.lib.ngm_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = total_ms / args.steps
    value = world * rays_per_step / (ms_per_step / 1e3)

    e2e_seq_ms = None
    if world == 1:
        e2e_seq_ms = timed(step_e2e, max(args.steps // 4, 3), args.warmup) / max(args.steps // 4, 3)
        e2e_ms = timed_e2e_two_streams(args.steps, args.warmup) / args.steps
        e2e_mode = "2 CUDA streams, steps alternate: step i+1's H2D and step i-1's D2H overlap step i's render; one " \
                   "device-timed bracket around all steps (the L2 flushes between them included)"
    else:
        e2e_ms = timed(step_e2e, args.steps, args.warmup) / args.steps
        e2e_mode = "one stream, per-step events"
    e2e_value = world * rays_per_step / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel (field MLP; tensor-bound), measured live with events ----
    peaks = measured_peaks()
    stage = stage_breakdown(st, dz, cam, precision, steps=max(3, min(args.steps, 10)), flush=flush)
    roofline = stage["dominant"]
    roofline["peak"] = peaks["tflops_sustained"] if precision == "fp16" else roofline.get("peak")
    if roofline.get("peak"):
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
    roofline["peak_source"] = peaks["source"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, cores, sample = cpu_reference_rate(8, 4096, repeats=3)
        cpu = {"value": r, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if precision == "fp16" else "f32", "data": "synthetic",
            "config": {
                "workload": "configs[1]: single 640x480 keyframe render, 64 samples/ray, 4-layer x 128 MLP, NeRF-8 "
                            "encoding (E=48), 75 fields x 4096 rays per keyframe, nrgbd compositing; one keyframe per GPU",
                "rays_per_step_per_gpu": rays_per_step, "samples_per_ray": S, "precision": precision,
                "l2": "flushed between timed iterations (256 MiB write)",
                "parallelism": f"rays sharded by keyframe x{world}, one NCCL all-gather of rendered tiles" if world > 1 else "single GPU",
            },
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "how": e2e_mode,
                    # the same step with nothing overlapped (one stream: H2D -> render -> D2H back to back)
                    "single_stream_ms_per_step": e2e_seq_ms,
                    "single_stream_value": (world * rays_per_step / (e2e_seq_ms / 1e3)) if e2e_seq_ms else None},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_stages": stage["stages"],
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def stage_breakdown(st, dz, cam, precision, steps, flush):
    """Per-kernel device time with CUDA events on the launching stream (torch's current stream).
    fp32 path: the three stage kernels through the stage entry points; fp16 path: the fused kernel."""
    import torch

    from neural_graph_mapping_b200 import models, renderer
    from neural_graph_mapping_b200.camera import sample_rays

    peaks = measured_peaks()
    n_rays = F_FIELDS * R_RAYS
    points = n_rays * S
    dev = dz["ijs"].device

    def ev_time(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    BURST = 10

    def ev_time_burst(entry, args):
        """Average duration of one launch over BURST back-to-back C-ABI calls per event pair.  For the ~0.1 ms
        HBM-bound stage kernels a single launch between two events also times the event/launch gaps and any
        host-side bubble (ncu's gpu__time_duration of the same launches is ~12% shorter); every launch streams
        its > L2-sized working set from HBM again, so no flush is needed inside the burst."""
        import ctypes as C

        from neural_graph_mapping_b200 import _lib
        sp = _lib.stream_ptr(torch.device(dev))
        ts = []
        for _ in range(max(steps // BURST, 3) + 1):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(BURST):
                _lib.check(entry(C.byref(args), sp))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / BURST)
        return sum(ts[1:]) / len(ts[1:])

    stages = {}
    with torch.no_grad():
        from neural_graph_mapping_b200 import _lib
        from neural_graph_mapping_b200.camera import sample_rays_args

        # sampler stage (HBM-bound)
        t1 = ev_time(lambda: sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], c2ws=dz["c2w"], seed=1,
                                         want_world=True, want_depth=True))
        sa, s_outs, s_keep = sample_rays_args(cam, dz["ijs"], S, dz["near"], dz["far"], c2ws=dz["c2w"], seed=1,
                                              want_world=True, want_depth=True)
        t = ev_time_burst(_lib.lib.ngm_sample_rays, sa)
        del s_outs, s_keep
        # the stage call also writes points_cam (12 B/sample) on top of the SURVEY figure
        b = n_rays * (SAMPLER_BYTES_PER_RAY + 12 * S)
        how = f"{BURST} back-to-back launches per event pair, working set > L2"
        stages["sampler"] = {"bound": "hbm", "ms": t, "achieved": b / t / 1e6, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": b / t / 1e6 / peaks["hbm_gbs"], "bytes_per_launch": b,
                             "timing": how, "ms_single_launch_after_flush": t1}
        pts_cam, dist, world_pts, depth = sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], c2ws=dz["c2w"], seed=1,
                                                      want_world=True, want_depth=True)
        del pts_cam
        # field stage (tensor-bound)
        model = st._model
        slots = dz["field_ids"]
        pos, ori = st._global_map_dict["positions"], st._global_map_dict["orientations"]
        q = world_pts.view(F_FIELDS, R_RAYS * S, 3)

        def field():
            return models.field_forward(model._prototype_field, model.all_fields_params, True, q, pos, ori, slots,
                                        model._scale_mode, model._field_radius, precision)

        t = ev_time(field)
        fl = points * FLOPS_PER_POINT
        peak = peaks["tflops_sustained"] if precision == "fp16" else None
        stages["field_mlp"] = {"bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak, "unit": "TFLOP/s",
                               "frac": (fl / t / 1e9 / peak) if peak else None, "flops_per_launch": fl,
                               "note": None if peak else "fp32 FFMA path: no tensor-pipe peak applies"}
        outs = field()
        # compositor stage (HBM-bound)
        o = outs.view(n_rays, S, 4)
        t1 = ev_time(lambda: renderer.composite(o, o[..., 3], dist.view(n_rays, S), depth.view(n_rays, S), "nrgbd", 20.0,
                                                color_stride=4, geometry_stride=4))
        ca, c_outs = renderer.composite_args(o, o[..., 3], dist.view(n_rays, S), depth.view(n_rays, S), "nrgbd", 20.0,
                                             color_stride=4, geometry_stride=4)
        t = ev_time_burst(_lib.lib.ngm_composite, ca)
        del c_outs
        b = n_rays * COMPOSITE_BYTES_PER_RAY
        stages["composite"] = {"bound": "hbm", "ms": t, "achieved": b / t / 1e6, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": b / t / 1e6 / peaks["hbm_gbs"], "bytes_per_launch": b,
                               "timing": how, "ms_single_launch_after_flush": t1}
        if precision == "fp16":
            # the product path: ONE fused tcgen05 kernel per render batch (sampler+encode+MLP+composite)
            t = ev_time(lambda: st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"]))
            stages["render_fused"] = {"bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak,
                                      "unit": "TFLOP/s", "frac": fl / t / 1e9 / peak, "flops_per_launch": fl,
                                      "hbm_bytes_per_launch_algorithmic": n_rays * FUSED_BYTES_PER_RAY}
    if precision == "fp16":
        dom = dict(stages["render_fused"])
        dom["kernel"] = "tc_kernel<0,8> fused render (tcgen05 MLP + sampler + composite)"
    else:
        dom = dict(stages["field_mlp"])
        dom["kernel"] = "field_fwd_simt_kernel (encode + MLP, fp32 FFMA)"
    dom["traffic"] = ncu_traffic("tc_kernel<0" if precision == "fp16" else "field_fwd_simt_kernel")
    dom["traffic_source"] = "profiles/r1_ncu_summary.json (committed ncu --set full capture of this kernel on this workload; not live)"
    return {"stages": stages, "dominant": dom}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (or None)."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_summary.json")
    try:
        for k in json.load(open(path))["kernels"]:
            if kernel_substr in k["kernel"]:
                return k.get("dram_read_bytes", 0.0) + k.get("dram_write_bytes", 0.0)
    except (OSError, ValueError, KeyError):
        pass
    return None


if __name__ == "__main__":
    main()
