"""Golden checkpoint written by the UNMODIFIED reference's own ``save_model`` (generated HERE).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Builds a 5-field map on CPU the way the driver holds it
(map tables over-allocated to 32 rows with ``num`` valid, ngm/run_mapping.py:231-246), calls
``NeuralGraphMap.save_model`` (:2147-2164; yoco is stubbed, so only the ``torch.save`` half takes effect) and copies
the ``.pt`` to ``tests/golden/checkpoint_ref.pt``; then renders one vmap batch and one kNN batch from that state
with the reference and stores inputs and outputs in ``tests/golden/checkpoint_render.npz``.

    python oracle/make_checkpoint_fixture.py
"""
import os
import shutil
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main():
    ref = ref_loader.load()
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(515)
    n = 5
    m, cfg = MG._make_map(ref, g, MG.NERF4, n, {"num_samples_coarse": 16, "num_samples_depth_guided": 0})
    # the driver's own over-allocated tables (:231-246): rows >= num are stale / zero
    for key, width in (("positions", 3), ("orientations", 4)):
        full = torch.zeros(32, width)
        full[:n] = m._global_map_dict[key]
        full[n:] = torch.randn(32 - n, width, generator=g)  # stale rows must never be read
        m._global_map_dict[key] = full
    m._global_map_dict["kf_ids"] = torch.arange(32)
    m._global_map_dict["training_iterations"] = torch.randint(0, 50, (32,), generator=g)
    with tempfile.TemporaryDirectory() as tmp:
        m._run_dir, m._run_name, m._metrics = tmp, "checkpoint_ref", None
        m.save_model()
        shutil.copy(os.path.join(tmp, "checkpoint_ref.pt"), os.path.join(MG.OUT, "checkpoint_ref.pt"))
    print("wrote checkpoint_ref.pt:", os.path.getsize(os.path.join(MG.OUT, "checkpoint_ref.pt")), "bytes")
    cam = ref.camera.Camera(**MG.CAM)
    F, R, S = 3, 40, 16
    ijs = torch.stack([torch.randint(0, 480, (F, R), generator=g), torch.randint(0, 640, (F, R), generator=g)], -1)
    c2ws = MG._rand_c2w(g, (F, R))
    near = torch.rand(F, R, generator=g) * 0.5 + 0.5
    far = near + 2.0
    jit = torch.rand(F, R, S, generator=g)
    fid = torch.tensor([4, 0, 2])
    with ref_loader.injected_jitter(jit), torch.no_grad():
        pv = m._render_ijs(ijs, c2ws, cam, fid, True, near.clone(), far.clone(), None)
    N = 96
    ijs_k = torch.stack([torch.randint(0, 480, (N,), generator=g), torch.randint(0, 640, (N,), generator=g)], -1)
    c2w_k = MG._rand_c2w(g)
    m._near_distance, m._far_distance = 0.3, 4.0
    jit_k = torch.rand(N, S, generator=g)
    with ref_loader.injected_jitter(jit_k), torch.no_grad():
        pk = m._render_ijs(ijs_k, c2w_k, cam, None, False)
    MG._save("checkpoint_render",
             {"case": "checkpoint", "camera": MG.CAM, "field_kwargs": MG.NERF4,
              "config": {k: cfg[k] for k in MG.RENDER_KEYS}, "num_samples": S, "near_distance": 0.3,
              "far_distance": 4.0, "num_fields": n},
             dict(ijs=MG._np(ijs), c2ws=MG._np(c2ws), near=MG._np(near), far=MG._np(far), jitter=MG._np(jit),
                  field_ids=MG._np(fid), out_rgbds=MG._np(pv.rgbds), out_term_probs=MG._np(pv.term_probs),
                  knn_ijs=MG._np(ijs_k), knn_c2w=MG._np(c2w_k), knn_jitter=MG._np(jit_k),
                  knn_out_rgbds=MG._np(pk.rgbds), knn_out_term_probs=MG._np(pk.term_probs)))


if __name__ == "__main__":
    main()
