"""Golden trajectory of the reference's TRAINING iteration (generated HERE; cannot travel to the GPU box).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Runs the UNMODIFIED reference on CPU for a few mapping
iterations -- ``_add_fields`` (ngm/run_mapping.py:365-389), ``_set_vmap_fields`` (:668-707), ``_render_ijs``
(:440-666), ``_compute_losses`` (:1769-1871) and ``_update_step`` (:1183-1221: backward, torch.optim.Adam, scatter of
parameters and moments) -- with changing sets of active fields and a field-growth step in between, and stores
every input, the loss, the gradients and the parameter / optimizer-state tables after every iteration in
``tests/golden/train_steps.npz``.

Only deviation from ``fit()``: the throw-away optimizer of ``_init_optimizer`` (:357-362) is created with its
dummy variable on the CPU (the reference hard-codes ``device="cuda"`` for it); hyper-parameters are the driver's.

    python oracle/make_training_fixture.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as MG  # noqa: E402
from oracle import ref_loader  # noqa: E402

SC, SG = 8, 16  # neural_graph_map.yaml:62-63
RAYS = 24
ACTIVE = [[0, 2, 3], [1, 2], "grow", [5, 0, 4], [2, 5], [0, 1, 2, 3, 4, 5]]


def _randomise(m, g, rows, geometry_scale=0.12, geometry_bias=0.0):
    """Independent seeded parameters for fields ``rows`` (add_fields replicates the prototype), geometry channel
    calibrated so that rays terminate (the depth / colour losses are masked by term_probs > 0.8, :1787)."""
    ap = m._model.all_fields_params
    last = f"_linears.{m._model._prototype_field._num_layers}"
    x = torch.rand(2048, 3, generator=g)
    with torch.no_grad():
        for k, v in ap.items():
            if v.dtype.is_floating_point and v.dim() > 1:
                scale = v.abs().max().clamp_min(1e-3)
                v[rows] = (torch.rand(v[rows].shape, generator=g) * 2 - 1) * scale
        for f in rows:
            params = {k: v[f] for k, v in ap.items()}
            gch = torch.func.functional_call(m._model._prototype_field, (params, {}), x)[:, 3]
            s = geometry_scale / gch.std().clamp_min(1e-6)
            ap[last + ".weight"][f, 3] *= s
            ap[last + ".bias"][f, 3] = (ap[last + ".bias"][f, 3] - gch.mean()) * s + geometry_bias


def main():
    ref = ref_loader.load()
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(909)
    cfg = ref_loader.default_config(model_kwargs={"field_kwargs": MG.NERF4}, num_samples_coarse=SC,
                                    num_samples_depth_guided=SG, termination_weight=0.5)
    m = ref.run_mapping.NeuralGraphMap(cfg)
    m._optimizer = torch.optim.Adam([torch.autograd.Variable(torch.tensor(0.0))], lr=m._learning_rate,
                                    eps=m._adam_eps, weight_decay=m._adam_weight_decay)
    n = 4
    m._global_map_dict["num"] = n
    m._add_fields(n)
    _randomise(m, g, list(range(n)))
    pos = torch.randn(6, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, -2.0])
    m._global_map_dict["positions"] = pos
    m._global_map_dict["orientations"] = MG._rand_quat(g, 6)
    m._global_map_dict["training_iterations"] = torch.zeros(6, dtype=torch.long)
    cam = ref.camera.Camera(**MG.CAM)
    arrays = {"positions": MG._np(pos), "orientations": MG._np(m._global_map_dict["orientations"])}
    for k, v in m._model.all_fields_params.items():
        arrays["init:param:" + k] = MG._np(v.clone())
    steps = []
    it = 0
    for act in ACTIVE:
        if act == "grow":
            m._global_map_dict["num"] = 6
            m._add_fields(2)
            _randomise(m, g, [4, 5])
            for k, v in m._model.all_fields_params.items():
                arrays[f"grow{it}:param:" + k] = MG._np(v[4:6].clone())
            steps.append({"grow": 2, "before_iteration": it})
            continue
        fid = torch.tensor(act)
        F = len(act)
        ijs = torch.stack([torch.randint(0, 480, (F, RAYS), generator=g), torch.randint(0, 640, (F, RAYS), generator=g)], -1)
        c2ws = MG._rand_c2w(g, (F, RAYS))
        near = torch.rand(F, RAYS, generator=g) * 0.5 + 0.5
        far = near + 1.5 + torch.rand(F, RAYS, generator=g)
        gt = near + (far - near) * (torch.rand(F, RAYS, generator=g) * 1.2 - 0.1)
        gt[torch.rand(F, RAYS, generator=g) < 0.15] = 0.0
        jit, jit_g = torch.rand(F, RAYS, SC, generator=g), torch.rand(F, RAYS, SG, generator=g)
        t_rgbd = torch.cat([torch.rand(F, RAYS, 3, generator=g), gt[..., None] * 0.9], -1)
        depth_mask = gt > 0
        term_target = (torch.rand(F, RAYS, generator=g) < 0.8).float()
        term_mask = torch.rand(F, RAYS, generator=g) < 0.9
        target = ref.run_mapping.Target(ijs, c2ws, near, far, gt, fid, t_rgbd, torch.ones_like(depth_mask), depth_mask,
                                        term_target, term_mask)
        with ref_loader.injected_jitter(jit, jit_g):
            pred = m._render_ijs(target.ijs, target.c2ws, cam, near_distances=near.clone(), far_distances=far.clone(),
                                 gt_distances=gt.clone(), field_ids=fid, use_vmap=True)
        losses = m._compute_losses(target, pred)
        n_depth = int((depth_mask * (pred.term_probs > 0.8)).sum())
        assert n_depth > 5, f"iteration {it}: only {n_depth} rays pass the depth mask -- degenerate fixture"
        assert torch.isfinite(losses["combined"])
        m._update_step(losses, fid)
        pre = f"it{it}:"
        arrays.update({pre + "field_ids": MG._np(fid), pre + "ijs": MG._np(ijs), pre + "c2ws": MG._np(c2ws),
                       pre + "near": MG._np(near), pre + "far": MG._np(far), pre + "gt": MG._np(gt),
                       pre + "jitter": MG._np(jit), pre + "jitter_guided": MG._np(jit_g),
                       pre + "target_rgbds": MG._np(t_rgbd), pre + "depth_mask": MG._np(depth_mask),
                       pre + "term_target": MG._np(term_target), pre + "term_mask": MG._np(term_mask),
                       pre + "out_rgbds": MG._np(pred.rgbds), pre + "out_term_probs": MG._np(pred.term_probs),
                       pre + "loss": MG._np(losses["combined"])})
        for k, v in m._model.vmap_fields_params.items():
            if v.grad is not None:
                arrays[pre + "grad:" + k] = MG._np(v.grad)
        for k, v in m._model.all_fields_params.items():
            arrays[pre + "param:" + k] = MG._np(v.clone())
            st = m._optim_state[k]
            arrays[pre + "exp_avg:" + k] = MG._np(st["exp_avg"].clone())
            arrays[pre + "exp_avg_sq:" + k] = MG._np(st["exp_avg_sq"].clone())
            arrays[pre + "step:" + k] = MG._np(torch.as_tensor(st["step"]).clone())
        steps.append({"iteration": it, "active": act, "rays_in_depth_mask": n_depth,
                      "loss": float(losses["combined"])})
        print(steps[-1])
        it += 1
    arrays["training_iterations"] = MG._np(m._global_map_dict["training_iterations"])
    loss_cfg = {k: cfg[k] for k in ("termination_weight", "photometric_weight", "photometric_loss", "depth_weight",
                                    "depth_loss", "freespace_weight", "tsdf_weight", "truncation_distance",
                                    "learning_rate", "adam_eps", "adam_weight_decay")}
    MG._save("train_steps", {"case": "train_steps", "camera": MG.CAM, "field_kwargs": MG.NERF4,
                             "config": {k: cfg[k] for k in MG.RENDER_KEYS}, "loss_config": loss_cfg,
                             "num_samples": SC, "num_samples_depth_guided": SG, "steps": steps}, arrays)


if __name__ == "__main__":
    main()
