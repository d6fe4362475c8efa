"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): one process per GPU over NCCL.
`render_rays_sharded` (fields partitioned over ranks, one all-gather of rendered tiles) must equal
the single-GPU render of the whole batch."""
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import golden_util as G
    import neural_graph_mapping_b200 as ngm
    from neural_graph_mapping_b200 import distributed as D
    from tests_support import make_state

    dev = f"cuda:{rank}"
    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=16, num_samples_depth_guided=0)
    g = torch.Generator().manual_seed(5)
    F, R = 4, 96
    ijs = torch.stack([torch.randint(0, 480, (F, R), generator=g), torch.randint(0, 640, (F, R), generator=g)], -1)
    near = torch.rand(F, R, generator=g) * 0.5 + 0.3
    far = near + 1.5
    jit = torch.rand(F, R, 16, generator=g)
    fid = torch.tensor([0, 3, 1, 4])
    cam = ngm.Camera(**meta["camera"])
    for prec in ("fp32", "fp16"):
        st = make_state(meta, a, dev, prec)
        c2w = a["c2ws"][0, 0].to(dev)
        with torch.no_grad():
            full = st._render_ijs(ijs.to(dev), c2w, cam, fid.to(dev), True, near.to(dev), far.to(dev), jitter=jit.to(dev))
            f0, f1 = D.shard_range(F, world, rank)
            shard = D.render_rays_sharded(st, ijs.to(dev), c2w, cam, fid.to(dev), near.to(dev), far.to(dev),
                                          jitter=jit[f0:f1].to(dev))       # pre-sliced jitter
            shard2 = D.render_rays_sharded(st, ijs.to(dev), c2w, cam, fid.to(dev), near.to(dev), far.to(dev),
                                           jitter=jit.to(dev))             # whole-batch jitter: sliced by the library
            split = D.render_rays_split(st, ijs.to(dev), c2w, cam, fid.to(dev), near.to(dev), far.to(dev),
                                        jitter=jit.to(dev))                # the same batch split by ray
            seed = D.shared_seed(torch.device(dev))
            full_s = st._render_ijs(ijs.to(dev), c2w, cam, fid.to(dev), True, near.to(dev), far.to(dev), seed=seed)
            shard_s = D.render_rays_sharded(st, ijs.to(dev), c2w, cam, fid.to(dev), near.to(dev), far.to(dev), seed=seed)
            shard_auto = D.render_rays_sharded(st, ijs.to(dev), c2w, cam, fid.to(dev), near.to(dev), far.to(dev))
        for other in (shard, shard2, split):
            for x, y in zip(full[:4], other[:4]):
                assert x.shape == y.shape
                assert torch.equal(x, y), (prec, (x - y).abs().max().item())
        for x, y in zip(full_s[:4], shard_s[:4]):  # in-kernel jitter: one seed, global sample indices
            assert torch.equal(x, y), (prec, "seeded", (x - y).abs().max().item())
        assert torch.isfinite(shard_auto.rgbds).all()
        # an asynchronous gather returns this rank's tile at once and the full buffer after wait()
        pend = D.render_rays_gathered(st, ijs[f0:f1].to(dev), c2w, cam, fid[f0:f1].to(dev), near[f0:f1].to(dev),
                                      far[f0:f1].to(dev), async_gather=True, jitter=jit[f0:f1].to(dev))
        buf = pend.wait()
        torch.cuda.synchronize()
        assert buf.shape[0] == world and torch.equal(buf[rank], pend.local)
        # config-5 re-render: full frame through the kNN path, pixel rows split over the ranks, one shared seed
        small = ngm.Camera(width=64, height=48, fx=55.4, fy=55.4, cx=31.5, cy=23.5)
        st.eval()
        with torch.no_grad():
            seed = D.shared_seed(torch.device(dev))
            rgbd_s, dvar_s = D.render_image_sharded(st, c2w, small, seed=seed)
            ij = torch.cartesian_prod(torch.arange(48, device=dev), torch.arange(64, device=dev))
            whole = st._render_ijs(ij, c2w, small, seed=seed)
        st.train()
        assert torch.equal(rgbd_s.view(-1, 4), whole.rgbds) and torch.equal(dvar_s.view(-1), whole.depth_vars), prec
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
