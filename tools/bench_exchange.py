"""Cost of the tile exchange at N GPUs (run under torch.distributed.run): device time of one configs[1] keyframe per
rank with (a) no exchange, (b) the exchange fused into the render kernel -- per-segment bulk repeat or per-ray stores,
through the NVSwitch multicast mapping or plain peer mappings -- and (c) an NCCL all-gather behind the render.
Every variant: L2 flushed between steps, the exchange's completion (barrier / collective) inside the timed bracket.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_exchange.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = f"cuda:{int(os.environ['LOCAL_RANK'])}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import distributed as D

    steps = int(os.environ.get("STEPS", "50"))
    sc = bench.synthetic_scene(1234 + rank)
    cam = ngm.Camera(**bench.CAMERA)
    st = ngm.RenderState(bench.config_dict(dev, "fp16"))
    st.set_fields(sc["params"], sc["positions"], sc["orientations"])
    dz = {k: sc[k].to(dev) for k in ("ijs", "c2w", "near", "far", "field_ids")}
    n = bench.F_FIELDS * bench.R_RAYS
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    local = torch.empty(D.FLOATS_PER_RAY * n, device=dev)
    gathered = torch.empty(world, D.FLOATS_PER_RAY * n, device=dev)

    def run(fn):
        for _ in range(5):
            fn()
            flush.fill_(1)
        dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()), 4)

    args = (st, dz["ijs"], dz["c2w"], cam, dz["field_ids"], dz["near"], dz["far"])
    out = {"n_gpus": world, "steps": steps, "unit": "ms per keyframe (max over ranks), exchange complete inside the bracket"}
    with torch.no_grad():
        out["render_only"] = run(lambda: D.render_rays_gathered(*args, return_packed=True, buffers=(local, gathered), group=None)
                                 if False else ngm.renderer.render_rays(*args[:5], True, *args[5:], out=tuple(
                                     v.view(bench.F_FIELDS, bench.R_RAYS, *v.shape[1:]) for v in D.packed_views(local, n))))
        out["nccl_all_gather"] = run(lambda: D.render_rays_gathered(*args, return_packed=True, buffers=(local, gathered)))
        for multicast in (True, False):
            ex = D.TileExchange(n, dev, slots=3, multicast=multicast)
            tag = "multicast" if ex.multicast else f"{len(ex.mirrors)}_peer_mappings"
            for per_ray in ("0", "1"):
                os.environ["NGM_MIRROR_PER_RAY"] = per_ray
                out[f"fused_{'per_ray' if per_ray == '1' else 'bulk'}_{tag}"] = run(
                    lambda: D.render_rays_gathered(*args, return_packed=True, exchange=ex))
            os.environ["NGM_MIRROR_PER_RAY"] = "0"
            out[f"barrier_only_{tag}"] = run(lambda: ex.finish(0))
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
