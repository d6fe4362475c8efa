#!/usr/bin/env python
"""bench.py -- rays/sec of the per-field ray renderer (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] -- one 640x480 keyframe render, 64 samples/ray,
4-layer x 128 MLP, NeRF-8 encoding (E=48): the 307,200 pixels pre-bucketed to 75 posed fields
x 4,096 rays (SURVEY.md 8d, C2 primary).  One *step* = one full keyframe (307,200 rays,
19,660,800 sample points).  N>1 (weak scaling): every rank renders one keyframe of its own and
the rendered tiles (36 B/ray) reach every rank inside the step (mirrored stores of the compositor kernel over
NVSwitch, or an NCCL all-gather).  A step is three kernel launches: sample_rays -> tcgen05 field kernel -> compositor.

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
C4_FIELDS, C4_RAYS = 256, 4096  # BASELINE configs[3]: 256 anchored fields, 1M-ray batch, sharded by field


def c4_scene(seed, num_fields=C4_FIELDS, rays=C4_RAYS, keyframes=64):
    """BASELINE configs[3]: fields on the reference's grid (cell 2r/sqrt(3),
    run_mapping.py:299), per-ray c2ws (F, R, 4, 4) drawn from `keyframes` seeded poses, the caller shape of the
    training step (run_mapping.py:1164).  CPU tensors; the same on every rank."""
    import torch

    g = torch.Generator().manual_seed(seed)
    side = int(math.ceil(num_fields ** 0.5))
    cell = 2.0 / math.sqrt(3.0)
    idx = torch.arange(num_fields)
    positions = torch.stack([(idx % side).float() * cell, torch.zeros(num_fields), -(idx // side).float() * cell - 2.0], -1)
    q = torch.randn(num_fields, 4, generator=g)
    orientations = q / q.norm(dim=-1, keepdim=True)
    # keyframe poses: small random rotations about the camera axes, eyes 2 m in front of a random field
    ang = (torch.rand(keyframes, 3, generator=g) - 0.5) * 0.6
    cx, sx, cy, sy, cz, sz = ang[:, 0].cos(), ang[:, 0].sin(), ang[:, 1].cos(), ang[:, 1].sin(), ang[:, 2].cos(), ang[:, 2].sin()
    Rm = torch.stack([cy * cz, sx * sy * cz - cx * sz, cx * sy * cz + sx * sz,
                      cy * sz, sx * sy * sz + cx * cz, cx * sy * sz - sx * cz,
                      -sy, sx * cy, cx * cy], -1).view(keyframes, 3, 3)
    kf = torch.eye(4).repeat(keyframes, 1, 1)
    kf[:, :3, :3] = Rm
    which = torch.randint(0, keyframes, (num_fields, rays), generator=g)
    c2ws = kf[which].clone()                                    # (F, R, 4, 4)
    c2ws[..., :3, 3] = positions[:, None] + c2ws[..., :3, 2] * 2.0  # eye = field centre + 2 m along camera +z (looks down -z)
    ijs = torch.stack([torch.randint(120, 360, (num_fields, rays), generator=g),
                       torch.randint(160, 480, (num_fields, rays), generator=g)], -1)
    near = torch.full((num_fields, rays), 1.0) + 0.05 * torch.rand(num_fields, rays, generator=g)
    far = torch.full((num_fields, rays), 3.0) - 0.05 * torch.rand(num_fields, rays, generator=g)
    base = synthetic_scene(seed, num_fields, 2)["params"]
    return dict(ijs=ijs, c2ws=c2ws, near=near, far=far, positions=positions, orientations=orientations, params=base,
                field_ids=torch.arange(num_fields))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 / strong-scaling / training extras")
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

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- precision: fp16 tensor-core path when the build has it, else the fp32 path ----
    precision = args.precision
    sc = synthetic_scene(1234 + rank)
    cam = ngm.Camera(**CAMERA)

    def make_state(prec, scene=None):
        scene = scene or sc
        st = ngm.RenderState(config_dict(dev, prec))
        st.set_fields(scene["params"], scene["positions"], scene["orientations"])
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
    tile_floats = distributed.FLOATS_PER_RAY * rays_per_step

    # ---- multi-GPU parity inside the bench (runs under torch.distributed.run on the multi-GPU box) ----
    # (on a scene that is the same on every rank: the timed scenes below differ per rank)
    parity = None
    if world > 1:
        try:
            parity = multi_gpu_parity(make_state(precision, synthetic_scene(4321)), cam, dev, world, rank, precision)
        except Exception as e:  # the headline line must still be printed; the failure is in it
            parity = {"all_ranks": False, "error": f"{type(e).__name__}: {e}"[:300]}

    # ---- device-timed step, inputs resident: render + all-gather of the rendered tiles ----
    # The all-gather of step i is left in flight (NCCL's own stream) and overlaps step i+1's render; its buffers are
    # double-buffered, and the step that reuses a buffer first orders itself after the collective that filled it.
    bufs = [(torch.empty(tile_floats, device=dev), torch.empty(world, tile_floats, device=dev)) for _ in range(2)]
    pending = [None, None]
    counter = [0]

    # The exchange fused into the compositor kernel (distributed.TileExchange: mirrored stores over NVSwitch multicast /
    # NVLink peer memory + a signal-pad barrier) replaces the NCCL all-gather when symmetric memory is available AND
    # one step through it equals the NCCL all-gather of the same step bit for bit on every rank.
    ex_dev = ex_e2e = None
    exchange_info = None
    if world > 1:
        exchange_info = {"mode": "nccl all-gather"}
        if os.environ.get("NGM_BENCH_EXCHANGE", "fused") == "fused":
            try:
                ex_dev, ex_e2e, exchange_info = fused_exchange(distributed, st, dz, cam, dev, rays_per_step, bufs[0])
            except Exception as e:
                ex_dev = ex_e2e = None
                exchange_info = {"mode": "nccl all-gather", "fused_unavailable": f"{type(e).__name__}: {e}"[:300]}

    def step_resident():
        i = counter[0] & 1
        counter[0] += 1
        if pending[i] is not None:
            pending[i].wait()
        with torch.no_grad():
            pending[i] = distributed.render_rays_gathered(st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"],
                                                          dz["far"], async_gather=True, buffers=bufs[i], exchange=ex_dev)

    def drain():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    def timed(fn, steps, warmup, clk=None, after=None):
        """W warm-up steps, then K steps each bracketed by CUDA events on the launching stream (the L2 flush
        between them is outside the brackets); `after` (e.g. waiting for the last collectives) is timed too.
        Barrier + synchronize on both sides; returns the max over ranks of the summed device time in ms."""
        for _ in range(warmup):
            fn()
            flush.fill_(1)
        if after:
            after()
        barrier()
        if clk:
            clk.mark_begin()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        if after:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            after()
            e1.record()
            evs.append((e0, e1))
        barrier()
        if clk:
            clk.mark_end()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    launches0 = _lib.lib.ngm_launch_count()
    total_ms = timed(step_resident, args.steps, args.warmup, clocks, after=drain)
    clock_info = clocks.stop()
    launches = (_lib.lib.ngm_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_per_step = total_ms / args.steps
    value = world * rays_per_step / (ms_per_step / 1e3)

    # ---- end to end: HOST buffers in, HOST result out, through the public API; the same method at every N ----
    # Two CUDA streams, steps alternate: step i+1's H2D and step i-1's D2H overlap step i's render (what a serving
    # loop does).  Every step copies its rays from pinned host memory, renders, all-gathers the tiles over NCCL
    # (every rank ends with every rank's tile on the device, in flight behind the next step) and copies ITS OWN
    # rendered tile back to pinned host memory.
    hz = {k: sc[k].pin_memory() for k in ("ijs", "near", "far", "c2w")}
    h_outs = [torch.empty(tile_floats, dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_bytes = sum(hz[k].numel() * hz[k].element_size() for k in hz)
    d2h_bytes = h_outs[0].numel() * 4
    e_bufs = [(torch.empty(tile_floats, device=dev), torch.empty(world, tile_floats, device=dev)) for _ in range(2)]

    def timed_e2e(steps, warmup, n_streams):
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        pend = [None, None]
        rendered = [None]  # event after the previous step's last render kernel

        def one(i):
            b = i % 2
            with torch.cuda.stream(streams[i % n_streams]), torch.no_grad():
                if pend[b] is not None:
                    pend[b].wait()
                d = {k: hz[k].to(dev, non_blocking=True) for k in hz}
                flush.fill_(1)
                # A step is three kernels.  With two streams the block scheduler may start step i+1's 2.2 ms field
                # kernel before step i's compositor; on one or two GPUs that interleaving is harmless (measured: e2e
                # 2.33 ms/step against 2.40 with the renders kept in step order), but it holds back step i's tile and
                # with it every rank that waits for it in the exchange barrier -- at N = 8 the unordered loop ran at
                # 2.91 ms/step against 2.40 device-timed.  So from N = 4 on the renders are kept in step order and only
                # the copies overlap them.
                if world > 2 and rendered[0] is not None:
                    torch.cuda.current_stream().wait_event(rendered[0])
                pend[b] = distributed.render_rays_gathered(st, d["ijs"], d["c2w"], cam, dz["field_ids"], d["near"], d["far"],
                                                           async_gather=True, buffers=e_bufs[b], exchange=ex_e2e)
                rendered[0] = torch.cuda.Event()
                rendered[0].record()
                h_outs[b].copy_(pend[b].local, non_blocking=True)

        def finish():
            for b in range(2):
                if pend[b] is not None:
                    with torch.cuda.stream(streams[b % n_streams]):
                        pend[b].wait()
                    pend[b] = None

        for i in range(warmup):
            one(i)
        finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event() for _ in streams]
        e0.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_event(e0)
        for i in range(steps):
            one(i)
        finish()
        for s_, m in zip(streams[1:], marks[1:]):
            m.record(s_)
            streams[0].wait_event(m)
        e1.record(streams[0])
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    e2e_seq_ms = timed_e2e(max(args.steps // 4, 3), args.warmup, 1) / max(args.steps // 4, 3)
    e2e_ms = timed_e2e(args.steps, args.warmup, 2) / args.steps
    e2e_mode = ("2 CUDA streams, steps alternate: step i+1's H2D and step i-1's D2H overlap step i's render; every rank "
                "copies its rays H2D and ITS OWN rendered tile D2H each step; the exchange of the tiles (config.parallelism) stays "
                "in flight behind the next step; one device-timed bracket around all steps (L2 flushes included), max over ranks"
                + ("; the renders of the two streams are kept in step order (only the copies overlap them)" if world > 2 else ""))
    e2e_value = world * rays_per_step / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel (field MLP; tensor-bound), measured live with events ----
    peaks = measured_peaks()
    stage = stage_breakdown(st, dz, cam, precision, steps=max(3, min(args.steps, 10)), flush=flush, full=(world == 1))
    roofline = stage["dominant"]
    if precision == "fp16":
        # a kernel timed in isolation at full clocks is measured against the BURST cuBLAS figure; the sustained
        # figure (measured under the 1 kW power cap at ~1.3 GHz) applies when this run's clocks were capped too
        capped = "sw_power_cap" in (clock_info.get("reasons") or []) or (
            clock_info.get("sm_mhz") and clock_info.get("sm_max_mhz") and clock_info["sm_mhz"] < 0.9 * clock_info["sm_max_mhz"])
        roofline["peak"] = peaks["tflops_sustained"] if capped else peaks["tflops_burst"]
        roofline["peak_kind"] = "sustained (clocks capped during this run)" if capped else "burst (full clocks, no cap during this run)"
        roofline["frac"] = roofline["achieved"] / roofline["peak"]
        roofline["frac_of_burst"] = roofline["achieved"] / peaks["tflops_burst"]
        roofline["frac_of_sustained"] = roofline["achieved"] / peaks["tflops_sustained"]
    roofline["peak_source"] = peaks["source"]

    extras = {}
    if not args.no_extras:
        try:
            extras = run_extras(ngm, distributed, make_state, cam, dev, world, rank, precision, flush, timed, barrier,
                                steps=max(5, min(args.steps, 20)), warmup=3, fused=ex_dev is not None)
        except Exception as e:  # never lose the headline line to an extra
            extras = {"extras_error": f"{type(e).__name__}: {e}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, cores, sample = cpu_reference_rate(8, 4096, repeats=3)
        cpu = {"value": r, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": sample + " (oracle/restatement.py: the reference's _render_ijs restated in torch CPU fp32 and pinned "
                                  "to golden outputs of the unmodified reference; the reference itself cannot travel to the GPU "
                                  "box, and ran 2-3x slower than this port when the survey timed it)"}

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
                "parallelism": (f"rays sharded by keyframe x{world}; "
                                + ("the rendered tiles reach every rank as mirrored stores of the compositor kernel itself "
                                   "(symmetric memory over NVSwitch) + one signal-pad barrier per step on a side stream"
                                   if ex_dev is not None else
                                   "one NCCL all-gather of rendered tiles per step, left in flight behind the next "
                                   "step's render")) if world > 1 else "single GPU",
            },
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "how": e2e_mode,
                    # the same step with nothing overlapped (one stream: H2D -> render -> gather -> D2H back to back)
                    "single_stream_ms_per_step": e2e_seq_ms,
                    "single_stream_value": world * rays_per_step / (e2e_seq_ms / 1e3)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_stages": stage["stages"],
            "cpu_baseline": cpu,
        }
        if parity is not None:
            line["parity"] = parity
        if exchange_info is not None:
            line["exchange"] = exchange_info
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def fused_exchange(distributed, st, dz, cam, dev, rays_per_step, nccl_bufs):
    """Set up the fused tile exchange and check it against the NCCL all-gather of the same step (same seed) on every
    rank; returns (exchange of the device-timed loop, exchange of the e2e loop, info) or raises."""
    import torch
    import torch.distributed as dist

    # 3 slots: one step of slack (single-stream loop); 4 slots: a slot is reused on the stream that last read it
    ex_dev = distributed.TileExchange(rays_per_step, dev, slots=3)
    ex_e2e = distributed.TileExchange(rays_per_step, dev, slots=4)
    ok = torch.ones(1, device=dev, dtype=torch.int32)
    with torch.no_grad():
        for ex in (ex_dev, ex_e2e):
            for _ in range(ex.slots + 1):  # every slot, and one reuse
                want = distributed.render_rays_gathered(st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"], dz["far"],
                                                        return_packed=True, buffers=nccl_bufs, seed=99)
                got = distributed.render_rays_gathered(st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"], dz["far"],
                                                       return_packed=True, exchange=ex, seed=99)
                if not torch.equal(got, want):
                    ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        raise RuntimeError("the fused exchange did not reproduce the NCCL all-gather bit for bit")
    info = {"mode": "fused into the kernel that writes the Prediction (compositor stage): mirrored stores over "
                    + ("the NVSwitch multicast mapping" if ex_dev.multicast else "NVLink peer mappings")
                    + " of a symmetric tile buffer + signal-pad barrier",
            "multicast": bool(ex_dev.multicast), "mirrors_per_store": len(ex_dev.mirrors),
            "equals_nccl_all_gather": True}
    return ex_dev, ex_e2e, info


def multi_gpu_parity(st, cam, dev, world, rank, precision):
    """Bit-identity of the multi-GPU partitions against the same batch rendered whole on this GPU, checked on
    every rank before anything is timed: (a) fields sharded over ranks with in-kernel jitter from one broadcast seed
    (`render_rays_sharded`, the C4 partition), (b) one keyframe split by ray with injected jitter (`render_rays_split`)."""
    import torch

    from neural_graph_mapping_b200 import distributed

    g = torch.Generator().manual_seed(99)  # the same batch on every rank
    F = 8 * world if F_FIELDS >= 8 * world else world
    F = min(F, F_FIELDS) // world * world
    R = 64 * world
    ijs = torch.stack([torch.randint(0, H, (F, R), generator=g), torch.randint(0, W_IMG, (F, R), generator=g)], -1).to(dev)
    near = (1.0 + 0.05 * torch.rand(F, R, generator=g)).to(dev)
    far = (3.0 - 0.05 * torch.rand(F, R, generator=g)).to(dev)
    c2ws = torch.eye(4).repeat(F, R, 1, 1)
    c2ws[..., :3, 3] = 0.01 * torch.randn(F, R, 3, generator=g)
    c2ws = c2ws.to(dev)
    fid = torch.arange(F, device=dev)
    jit = torch.rand(F, R, S, generator=g).to(dev)
    res = {}
    with torch.no_grad():
        seed = distributed.shared_seed(torch.device(dev))
        whole = st._render_ijs(ijs, c2ws, cam, fid, True, near, far, seed=seed)
        shard = distributed.render_rays_sharded(st, ijs, c2ws, cam, fid, near, far, seed=seed)
        res["sharded_by_field_equals_whole"] = all(torch.equal(a, b) for a, b in zip(whole[:4], shard[:4]))
        whole_j = st._render_ijs(ijs, c2ws, cam, fid, True, near, far, jitter=jit)
        split = distributed.render_rays_split(st, ijs, c2ws, cam, fid, near, far, jitter=jit)
        res["split_by_ray_equals_whole"] = all(torch.equal(a, b) for a, b in zip(whole_j[:4], split[:4]))
    ok = torch.tensor([int(all(res.values()))], device=dev)
    import torch.distributed as dist

    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["all_ranks"] = bool(ok.item())
    res["batch"] = f"{F} fields x {R} rays x {S} samples, per-ray c2ws, precision {precision}"
    if rank == 0:
        tag = "OK" if res["all_ranks"] else "FAILED"
        print(f"[bench] multi-GPU parity {tag} on {world} ranks: {res}", file=sys.stderr, flush=True)
    return res


def run_extras(ngm, distributed, make_state, cam, dev, world, rank, precision, flush, timed, barrier, steps, warmup,
               fused=False):
    """Workloads BASELINE.json names beside the headline one, as extra keys of the same JSON line:
    `c4` -- configs[3]: 256 fields x 4096 rays x 64 samples with per-ray poses, the batch sharded by FIELD over the N
            GPUs (strong scaling: the total work is fixed), one all-gather of the rendered tiles;
    `strong_c2` -- ONE 640x480 keyframe split by ray over the N GPUs (strong scaling of the headline workload)."""
    import torch

    out = {}
    # ---- C4 ----
    c4 = c4_scene(4242)
    st4 = make_state(precision, c4)
    d4 = {k: c4[k].to(dev) for k in ("ijs", "c2ws", "near", "far", "field_ids")}
    n4 = C4_FIELDS * C4_RAYS
    how = ("tiles exchanged by the compositor kernel itself (distributed.TileExchange)" if fused
           else "NCCL all-gather of the tiles")
    if C4_FIELDS % world == 0:
        seed_box = [0]
        ex4 = distributed.TileExchange(n4 // world, dev, slots=3) if fused else None

        def step_c4():
            seed_box[0] += 1
            with torch.no_grad():
                distributed.render_rays_sharded(st4, d4["ijs"], d4["c2ws"], cam, d4["field_ids"], d4["near"], d4["far"],
                                                seed=seed_box[0], exchange=ex4)

        ms = timed(step_c4, steps, warmup) / steps
        out["c4"] = {"workload": "configs[3]: 256 anchored fields x 4096 rays (1,048,576 rays) x 64 samples, per-ray c2ws, "
                                 "4-layer x 128 MLP NeRF-8; fields sharded over the GPUs, " + how + "; the "
                                 "forward render of the batch (the training step's fwd+bwd is `train`)",
                     "value": n4 / (ms / 1e3), "unit": "rays/s", "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                     "rays_per_step": n4, "steps": steps}
    # ---- the training step on the C4 batch: fields sharded over the GPUs, NO collective (a field's gradient is local) ----
    if C4_FIELDS % world == 0 and precision == "fp16":
        f0, f1 = distributed.shard_range(C4_FIELDS, world, rank)
        st4._reference_flow = True  # the driver's flow: _render_ijs gathers the active fields into leaves (:500)
        sl = {k: d4[k][f0:f1].contiguous() for k in ("ijs", "c2ws", "near", "far", "field_ids")}

        def step_train():
            p = st4._render_ijs(sl["ijs"], sl["c2ws"], cam, sl["field_ids"], True, sl["near"], sl["far"])
            loss = p.rgbds.square().mean() + p.depth_vars.mean() + p.term_probs.mean()
            st4._update_step({"combined": loss}, sl["field_ids"])

        tsteps = max(3, steps // 4)
        ms = timed(step_train, tsteps, 2) / tsteps
        out["train_c4"] = {"workload": "configs[3] training step: render under autograd (tcgen05 forward) -> loss -> backward "
                                       "(tcgen05 MLP backward, compositor backward) -> Adam on the rank's own fields; "
                                       f"{C4_FIELDS // world} fields x {C4_RAYS} rays x {S} samples per GPU, no collective",
                           "value": n4 / (ms / 1e3), "unit": "rays/s", "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                           "rays_per_step": n4, "steps": tsteps}
    del st4, d4, c4
    # ---- C5: pose-graph update (field poses move) + full re-render through the kNN path, pixel rows over the GPUs ----
    if H % world == 0:
        c5 = c4_scene(777)
        cfg5 = config_dict(dev, precision)
        cfg5.update(eval_near_distance=0.5, eval_far_distance=4.5, eval_num_samples=S)
        st5 = ngm.RenderState(cfg5)
        st5.set_fields(c5["params"], c5["positions"], c5["orientations"])
        st5.eval()
        c2w5 = torch.eye(4, device=dev)
        c2w5[:3, 3] = torch.tensor([8.0, 0.0, 1.0], device=dev)  # inside the 16 x 16 grid of fields, looking down -z
        kf = torch.arange(C4_FIELDS, device=dev) % 3             # a third of the fields hangs on each of 3 keyframes
        ang = torch.tensor(0.01, device=dev)
        rot = torch.eye(3, device=dev)
        rot[0, 0] = rot[2, 2] = torch.cos(ang)
        rot[0, 2] = torch.sin(ang)
        rot[2, 0] = -torch.sin(ang)
        dq = torch.tensor([math.cos(0.005), 0.0, math.sin(0.005), 0.0], device=dev)  # the same rotation as a quaternion
        pivot = c5["positions"].mean(0).to(dev)
        seed5 = [0]

        def quat_mul(a_, b_):
            aw, ax, ay, az = a_.unbind(-1)
            bw, bx, by, bz = b_.unbind(-1)
            return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                                aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)

        def step_c5():
            # loop closure: the fields anchored to keyframe (step mod 3) move rigidly (run_mapping.py:937-952 changes
            # exactly these two tensors), then the keyframe is re-rendered
            seed5[0] += 1
            m = kf == (seed5[0] % 3)
            g_ = st5._global_map_dict
            with torch.no_grad():
                g_["positions"][m] = (g_["positions"][m] - pivot) @ rot.T + pivot
                g_["orientations"][m] = quat_mul(dq.expand(int(m.sum()), 4), g_["orientations"][m])
            distributed.render_image_sharded(st5, c2w5, cam, seed=seed5[0])

        ms = timed(step_c5, steps, warmup) / steps
        out["c5"] = {"workload": "configs[4]: pose-graph update (a third of 256 field poses moves) + full 640x480 x 64 re-render "
                                 "through the kNN path (K=2 blend over all fields), pixel rows split over the GPUs, all-gather",
                     "value": H * W_IMG / (ms / 1e3), "unit": "rays/s", "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                     "steps": steps}
        del st5, c5
    # ---- strong scaling of the headline keyframe ----
    if world > 1 and R_RAYS % world == 0:
        sc = synthetic_scene(1234)  # ONE keyframe, the same on every rank
        st = make_state(precision, sc)
        dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
        n_local = F_FIELDS * R_RAYS // world
        tf = distributed.FLOATS_PER_RAY * n_local
        bufs = [(torch.empty(tf, device=dev), torch.empty(world, tf, device=dev)) for _ in range(2)]
        pend = [None, None]
        cnt = [0]
        exs = distributed.TileExchange(n_local, dev, slots=3) if fused else None

        def step_split():
            i = cnt[0] & 1
            cnt[0] += 1
            if pend[i] is not None:
                pend[i].wait()
            with torch.no_grad():
                pend[i] = distributed.render_rays_split(st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"], dz["far"],
                                                        async_gather=True, buffers=bufs[i], exchange=exs)

        def drain():
            for i in range(2):
                if pend[i] is not None:
                    pend[i].wait()
                    pend[i] = None

        ms = timed(step_split, steps, warmup, after=drain) / steps
        out["strong_c2"] = {"workload": "configs[1] keyframe (307,200 rays x 64) split by ray over the GPUs, " + how,
                            "value": F_FIELDS * R_RAYS / (ms / 1e3), "unit": "rays/s", "ms_per_step": ms, "n_gpus": world,
                            "scaling": "strong", "steps": steps}
    return out


def stage_breakdown(st, dz, cam, precision, steps, flush, full=True):
    """Per-kernel device time with CUDA events on the launching stream (torch's current stream).
    The three stage kernels through the stage entry points (fp32: FFMA field kernel, fp16: tcgen05 field kernel), the
    whole step as `_render_ijs` runs it, and the opt-in single fused kernel.
    `full=False` (N > 1): only the dominant kernel."""
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
    fl = points * FLOPS_PER_POINT
    peak = peaks["tflops_burst"] if precision == "fp16" else None
    with torch.no_grad():
        from neural_graph_mapping_b200 import _lib
        from neural_graph_mapping_b200.camera import sample_rays_args

        if full:
            # sampler stage (HBM-bound) with exactly the outputs the renderer needs (world points,
            # distances, depths = 24 + 20 S bytes per ray; points_cam, which only Camera.sample_ijs_uniform returns, off)
            kw = dict(c2ws=dz["c2w"], seed=1, want_world=True, want_depth=True, want_cam=False)
            t1 = ev_time(lambda: sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], **kw))
            sa, s_outs, s_keep = sample_rays_args(cam, dz["ijs"], S, dz["near"], dz["far"], **kw)
            t = ev_time_burst(_lib.lib.ngm_sample_rays, sa)
            del s_outs, s_keep
            b = n_rays * SAMPLER_BYTES_PER_RAY
            how = f"{BURST} back-to-back launches per event pair, working set > L2"
            stages["sampler"] = {"bound": "hbm", "ms": t, "achieved": b / t / 1e6, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": b / t / 1e6 / peaks["hbm_gbs"], "bytes_per_launch": b,
                                 "timing": how, "ms_single_launch_after_flush": t1}
            _, dist, world_pts, depth = sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], **kw)
            # field stage (tensor-bound)
            model = st._model
            slots = dz["field_ids"]
            pos, ori = st._global_map_dict["positions"], st._global_map_dict["orientations"]
            q = world_pts.view(F_FIELDS, R_RAYS * S, 3)

            def field():
                return models.field_forward(model._prototype_field, model.all_fields_params, True, q, pos, ori, slots,
                                            model._scale_mode, model._field_radius, precision)

            t = ev_time(field)
            stages["field_mlp"] = {"bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak, "unit": "TFLOP/s",
                                   "frac": (fl / t / 1e9 / peak) if peak else None, "flops_per_launch": fl,
                                   "note": "peak = burst cuBLAS bf16" if peak else "fp32 FFMA path: no tensor-pipe peak applies"}
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
        else:
            # N > 1: only the dominant kernel (the field stage), on the sampler's world points
            _, _, world_pts, _ = sample_rays(cam, dz["ijs"], S, dz["near"], dz["far"], c2ws=dz["c2w"], seed=1, want_world=True,
                                             want_depth=True, want_cam=False)
            model = st._model
            pos, ori = st._global_map_dict["positions"], st._global_map_dict["orientations"]
            q = world_pts.view(F_FIELDS, R_RAYS * S, 3)
            t = ev_time(lambda: models.field_forward(model._prototype_field, model.all_fields_params, True, q, pos, ori,
                                                     dz["field_ids"], model._scale_mode, model._field_radius, precision))
            stages["field_mlp"] = {"bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak, "unit": "TFLOP/s",
                                   "frac": (fl / t / 1e9 / peak) if peak else None, "flops_per_launch": fl}
        if precision == "fp16":
            # the whole step as the public call runs it: sample_rays -> tcgen05 field kernel -> compositor
            t = ev_time(lambda: st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"]))
            stages["render_step"] = {"bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak,
                                     "unit": "TFLOP/s", "frac": fl / t / 1e9 / peak, "flops_per_launch": fl,
                                     "kernels": "sample_rays_kernel + tc_kernel<1,8> + composite_staged_kernel"}
            if full:
                # the opt-in single fused kernel (NGM_RENDER_FUSED=1: sampler + encoding + MLP + compositor in one
                # persistent tcgen05 kernel), for comparison
                os.environ["NGM_RENDER_FUSED"] = "1"
                try:
                    t = ev_time(lambda: st._render_ijs(dz["ijs"], dz["c2w"], cam, dz["field_ids"], True, dz["near"], dz["far"]))
                finally:
                    os.environ.pop("NGM_RENDER_FUSED", None)
                stages["render_single_fused_kernel_opt_in"] = {
                    "bound": "tensor", "ms": t, "achieved": fl / t / 1e9, "peak": peak, "unit": "TFLOP/s",
                    "frac": fl / t / 1e9 / peak, "flops_per_launch": fl,
                    "hbm_bytes_per_launch_algorithmic": n_rays * FUSED_BYTES_PER_RAY}
    if precision == "fp16":
        dom = dict(stages["field_mlp"])
        dom["kernel"] = "tc_kernel<1,8> field stage (NeRF-8 encoding + 4x128 MLP on tcgen05; 92 % of the step's device time)"
        dom["hbm_bytes_per_launch_algorithmic"] = points * (12 + 16)  # world point in, rgb + geometry out
    elif "field_mlp" in stages:
        dom = dict(stages["field_mlp"])
        dom["kernel"] = "field_fwd_simt_kernel (encode + MLP, fp32 FFMA)"
    else:
        dom = {"bound": "tensor", "kernel": "field_fwd_simt_kernel", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None}
    dom["traffic"] = ncu_traffic("tc_kernel<1, 8" if precision == "fp16" else "field_fwd_simt_kernel")
    dom["traffic_source"] = "profiles/ ncu summary (committed ncu --set full capture of this kernel on this workload; not live)"
    return {"stages": stages, "dominant": dom}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed ncu summary (or None)."""
    for name in ("r2_ncu_summary_final.json", "r2_ncu_summary.json", "r1_ncu_summary.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            for k in json.load(open(path))["kernels"]:
                if kernel_substr in k["kernel"]:
                    return k.get("dram_read_bytes", 0.0) + k.get("dram_write_bytes", 0.0)
        except (OSError, ValueError, KeyError):
            pass
    return None


if __name__ == "__main__":
    main()
