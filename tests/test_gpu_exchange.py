"""The multi-GPU tile exchange fused into the render kernel (`NgmRenderArgs.mirror_delta`, `distributed.TileExchange`):
every Prediction store is repeated at byte offsets that lead to peer / multicast mappings of the same tile.

* one GPU: a mirror that points at a second local buffer must receive exactly the rendered tile (the kernel side: the
  compositor of the stage pipeline, and the opt-in single fused kernel -- every test runs under both, conftest.py);
* two GPUs (skipped otherwise): the exchange equals the NCCL all-gather of the same step bit for bit, through the
  multicast mapping and through plain peer mappings, over more steps than the ring has slots."""
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _scene(dev, F=4, R=96, S=16):
    import golden_util as G
    import neural_graph_mapping_b200 as ngm
    from tests_support import make_state

    meta, a = G.load("vmap_guided_nrgbd")
    meta = dict(meta, num_samples=S, num_samples_depth_guided=0)
    g = torch.Generator().manual_seed(11)
    ijs = torch.stack([torch.randint(0, 480, (F, R), generator=g), torch.randint(0, 640, (F, R), generator=g)], -1).to(dev)
    near = (torch.rand(F, R, generator=g) * 0.5 + 0.3).to(dev)
    far = near + 1.5
    fid = torch.tensor([0, 3, 1, 4][:F]).to(dev)
    cam = ngm.Camera(**meta["camera"])
    c2w = a["c2ws"][0, 0].to(dev)
    return meta, a, make_state, cam, ijs, c2w, fid, near, far


def test_mirrored_stores_single_gpu():
    from neural_graph_mapping_b200 import distributed as D
    from neural_graph_mapping_b200.renderer import render_rays

    dev = "cuda:0"
    meta, a, make_state, cam, ijs, c2w, fid, near, far = _scene(dev)
    F, R = ijs.shape[:2]
    n = F * R
    st = make_state(meta, a, dev, "fp16")
    both = torch.zeros(3, D.FLOATS_PER_RAY * n, device=dev)
    views = [D.packed_views(both[i], n) for i in range(3)]
    out = tuple(v.view(F, R, *v.shape[1:]) for v in views[0])
    delta = both[1].data_ptr() - both[0].data_ptr()
    with torch.no_grad():
        want = render_rays(st, ijs, c2w, cam, fid, True, near, far, seed=5)
        render_rays(st, ijs, c2w, cam, fid, True, near, far, seed=5, out=out, mirrors=[delta, 2 * delta])
    torch.cuda.synchronize()
    assert torch.equal(out[0], want.rgbds) and torch.equal(out[3], want.term_probs)
    assert torch.equal(both[1], both[0]) and torch.equal(both[2], both[0])
    assert both[0].abs().sum() > 0


def test_mirrors_are_refused_where_they_do_not_exist():
    from neural_graph_mapping_b200 import distributed as D
    from neural_graph_mapping_b200.renderer import render_rays

    dev = "cuda:0"
    meta, a, make_state, cam, ijs, c2w, fid, near, far = _scene(dev)
    F, R = ijs.shape[:2]
    n = F * R
    both = torch.zeros(2, D.FLOATS_PER_RAY * n, device=dev)
    out = tuple(v.view(F, R, *v.shape[1:]) for v in D.packed_views(both[0], n))
    delta = both[1].data_ptr() - both[0].data_ptr()
    with torch.no_grad():
        # the fp32 path ends in the same compositor kernel: mirrors work there too
        render_rays(make_state(meta, a, dev, "fp32"), ijs, c2w, cam, fid, True, near, far, out=out, mirrors=[delta])
        torch.cuda.synchronize()
        assert torch.equal(both[1], both[0]) and both[0].abs().sum() > 0
        both.zero_()
        st = make_state(meta, a, dev, "fp16")
        with pytest.raises(ValueError):  # 16-byte granularity of the rgbd vector store
            render_rays(st, ijs, c2w, cam, fid, True, near, far, out=out, mirrors=[delta + 4])
        with pytest.raises(ValueError):
            render_rays(st, ijs, c2w, cam, fid, True, near, far, mirrors=[delta])  # no out= to mirror
        with pytest.raises(ValueError):
            render_rays(st, ijs, c2w, cam, fid, True, near, far, out=out, mirrors=[delta] * 9)
    torch.cuda.synchronize()
    assert both[1].abs().sum() == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from neural_graph_mapping_b200 import distributed as D

    dev = f"cuda:{rank}"
    meta, a, make_state, cam, ijs, c2w, fid, near, far = _scene(dev, F=4, R=100 + 0)  # 9 * 400 floats: needs no padding
    st = make_state(meta, a, dev, "fp16")
    ijs = (ijs + rank) % 480  # every rank renders its own rays
    n = ijs.shape[0] * ijs.shape[1]
    modes = []
    for multicast in (True, False):
        ex = D.TileExchange(n, dev, slots=3, multicast=multicast)
        modes.append(ex.multicast)
        assert len(ex.mirrors) == (1 if ex.multicast else world - 1)
        with torch.no_grad():
            for step in range(7):  # more steps than slots: the ring is reused
                want = D.render_rays_gathered(st, ijs, c2w, cam, fid, near, far, return_packed=True, seed=100 + step)
                if step % 2:
                    got = D.render_rays_gathered(st, ijs, c2w, cam, fid, near, far, return_packed=True, seed=100 + step,
                                                 exchange=ex)
                else:
                    pend = D.render_rays_gathered(st, ijs, c2w, cam, fid, near, far, async_gather=True, seed=100 + step,
                                                  exchange=ex)
                    got = pend.wait()
                    assert torch.equal(pend.local, got[rank])
                assert got.shape == want.shape
                assert torch.equal(got, want), (multicast, step, (got - want).abs().max().item())
                assert not torch.equal(got[0], got[1])  # the ranks really rendered different tiles
            preds = D.render_rays_gathered(st, ijs, c2w, cam, fid, near, far, seed=3, exchange=ex)
            assert len(preds) == world and preds[rank].rgbds.shape == (ijs.shape[0], ijs.shape[1], 4)
        torch.cuda.synchronize()
        dist.barrier()
    # the other partitions ride on the same exchange: one batch split by ray, one batch sharded by field
    F, R = ijs.shape[0], ijs.shape[1]
    ex_split = D.TileExchange(F * R // world, dev)
    ex_shard = D.TileExchange(F // world * R, dev)
    g = torch.Generator().manual_seed(3)
    jit = torch.rand(F, R, 16, generator=g).to(dev)
    ijs0 = (ijs - rank) % 480  # the SAME batch on every rank
    with torch.no_grad():
        for _ in range(4):
            a_ = D.render_rays_split(st, ijs0, c2w, cam, fid, near, far, jitter=jit)
            b_ = D.render_rays_split(st, ijs0, c2w, cam, fid, near, far, jitter=jit, exchange=ex_split)
            c_ = D.render_rays_sharded(st, ijs0, c2w, cam, fid, near, far, seed=17)
            d_ = D.render_rays_sharded(st, ijs0, c2w, cam, fid, near, far, seed=17, exchange=ex_shard)
            for x, y in zip(a_[:4] + c_[:4], b_[:4] + d_[:4]):
                assert torch.equal(x, y)
    torch.cuda.synchronize()
    if rank == 0:
        print("multicast used:", modes)
    dist.barrier()
    dist.destroy_process_group()


def test_fused_exchange_equals_all_gather():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
