"""Golden output of the reference's multi-view target sampling (generated HERE; cannot travel to the GPU box).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Builds the keyframe store the driver keeps (contiguous poses
``_c_c2w_tensor``, the over-allocated RGB-D buffer ``_nc_rgbd_tensor`` and ``_frame_cid_to_ncid``,
ngm/run_mapping.py:1674-1713) for a small synthetic room on CPU, calls the UNMODIFIED
``NeuralGraphMap._sample_target_mv`` (:1261-1459) while recording what ``torch.multinomial`` / ``torch.randn`` /
``torch.rand`` return, and stores the store, the draws and the resulting ``Target`` in
``tests/golden/target_mv.npz``.

    python oracle/make_target_fixture.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_loader  # noqa: E402

CAM = dict(width=64, height=48, fx=55.0, fy=55.0, cx=31.5, cy=23.5, pixel_center=0.0)


class recorded_draws:
    """Record (not replace) the reference's random draws, in call order."""

    def __init__(self):
        self.calls = []

    def __enter__(self):
        self._orig = {n: getattr(torch, n) for n in ("multinomial", "randn", "rand")}
        for n, fn in self._orig.items():
            def wrap(*a, _n=n, _fn=fn, **k):
                out = _fn(*a, **k)
                self.calls.append((_n, out.clone()))
                return out
            setattr(torch, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self._orig.items():
            setattr(torch, n, fn)
        return False


def main():
    ref = ref_loader.load()
    torch.manual_seed(4242)
    g = torch.Generator().manual_seed(77)
    cfg = ref_loader.default_config(model_kwargs={"field_kwargs": MG.NERF4}, num_train_fields=8, num_rays_per_field=32)
    m = ref.run_mapping.NeuralGraphMap(cfg)
    m._camera = ref.camera.Camera(**CAM)
    H, W = CAM["height"], CAM["width"]
    # 10 fields in front of the cameras, two far off to the side / behind (never visible)
    n = 10
    pos = torch.zeros(32, 3)
    pos[:n] = torch.randn(n, 3, generator=g) * torch.tensor([1.2, 0.8, 0.6]) + torch.tensor([0.0, 0.0, -3.0])
    pos[8] = torch.tensor([0.0, 0.0, 6.0])     # behind every camera
    pos[9] = torch.tensor([40.0, 0.0, -3.0])   # far outside every frustum
    pos[n:] = torch.randn(32 - n, 3, generator=g)  # stale rows
    m._global_map_dict["positions"] = pos
    m._global_map_dict["num"] = n
    # 5 contiguous frames stored in rows (0, 2, 3, 5, 7) of an 8-row buffer
    K = 5
    c2ws = torch.eye(4).repeat(K, 1, 1)
    ang = (torch.rand(K, generator=g) - 0.5) * 0.5
    c2ws[:, 0, 0], c2ws[:, 0, 2], c2ws[:, 2, 0], c2ws[:, 2, 2] = ang.cos(), ang.sin(), -ang.sin(), ang.cos()
    c2ws[:, :3, 3] = (torch.rand(K, 3, generator=g) - 0.5) * torch.tensor([1.5, 0.6, 0.8])
    m._c_c2w_tensor = c2ws
    rows = torch.tensor([0, 2, 3, 5, 7])
    m._frame_cid_to_ncid = rows
    rgbd = torch.rand(8, H, W, 4, generator=g)
    ii, jj = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    for r in range(8):  # smooth depth: a tilted back wall between 3 and 5 m, a nearer blob
        d = 4.0 + 0.6 * torch.sin(jj / 9.0 + r) + 0.4 * torch.cos(ii / 7.0 - r)
        d = torch.where(((ii - 24) ** 2 + (jj - 20 - 3 * r) ** 2) < 90, torch.full_like(d, 2.2), d)
        rgbd[r, ..., 3] = d
    rgbd[..., 3][torch.rand(8, H, W, generator=g) < 0.08] = 0.0       # missing depth
    dark = torch.rand(8, H, W, generator=g) < 0.05
    rgbd[..., 0][dark] = 0.0
    rgbd[..., 1][dark] = 0.0                                            # rgb_mask false
    m._nc_rgbd_tensor = rgbd
    current = torch.tensor([1, 4, 7])
    with recorded_draws() as rec:
        t = m._sample_target_mv(current)
    names = [c[0] for c in rec.calls]
    assert names == ["multinomial", "multinomial", "randn", "multinomial", "rand"], names
    draws = dict(zip(["subset_observed", "subset_random", "probe_offsets", "frame_cids", "uv"], [c[1] for c in rec.calls]))
    print("target fields:", t.field_ids.tolist(), "rays:", tuple(t.ijs.shape),
          "depth_mask %.2f term_mask %.2f rgb_mask %.2f term_probs %.2f" % (
              t.depth_mask.float().mean(), t.term_mask.float().mean(), t.rgb_mask.float().mean(), t.term_probs.mean()))
    assert 8 not in t.field_ids.tolist() and 9 not in t.field_ids.tolist()
    arrays = dict(positions=MG._np(pos), c2ws=MG._np(c2ws), rgbds=MG._np(rgbd), frame_to_store=MG._np(rows),
                  current_field_ids=MG._np(current))
    arrays.update({"draw:" + k: MG._np(v) for k, v in draws.items()})
    arrays.update({"out:" + k: MG._np(getattr(t, k)) for k in t._fields})
    # ---- _get_observed_fields (:1643-1670) on one stored frame, from the pose of keyframe 2
    rgbd_image = rgbd[3]
    with recorded_draws() as rec2:
        obs = m._get_observed_fields(rgbd_image, c2ws[2])
    assert [c[0] for c in rec2.calls] == ["multinomial"]
    print("observed fields:", obs.tolist())
    assert 0 < len(obs) < n
    arrays.update({"observed:rgbd": MG._np(rgbd_image), "observed:c2w": MG._np(c2ws[2]),
                   "observed:draw_subset": MG._np(rec2.calls[0][1]), "observed:out": MG._np(obs)})
    MG._save("target_mv", {"case": "target_mv", "camera": CAM, "num_fields": n, "num_train_fields": 8,
                           "num_rays_per_field": 32, "field_radius": cfg["field_radius"]}, arrays)


if __name__ == "__main__":
    main()
