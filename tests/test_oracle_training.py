"""oracle/training.py (losses + Adam on the active fields) against the golden trajectory of the unmodified
reference's training iteration (tests/golden/train_steps.npz, made by oracle/make_training_fixture.py)."""
import torch

import golden_util as G
from oracle import training as T


def test_training_trajectory_matches_reference():
    meta, a = G.load("train_steps")
    fs, rs, cam = G.field_spec(meta["field_kwargs"]), G.render_spec(meta), G.camera_spec(meta["camera"])
    lc = meta["loss_config"]
    ls = T.LossSpec(**{k: lc[k] for k in T.LossSpec.__dataclass_fields__})
    state = {"params": {k: v.clone() for k, v in G.params(a, "init:param:").items()}}
    state["optim"] = T.new_optim_state(state["params"])

    def grow(rows):
        n = next(iter(rows.values())).shape[0]
        state["params"] = {k: torch.cat([v, rows[k]]) for k, v in state["params"].items()}
        state["optim"] = T.new_optim_state(state["params"], state["optim"], n)

    def iteration(it, x):
        losses, grads, pred = T.training_iteration(
            state["params"], state["optim"], a["positions"], a["orientations"], x["field_ids"], cam, rs, fs, ls,
            x["ijs"], x["c2ws"], x["near"], x["far"], x["gt"], x["jitter"], x["jitter_guided"], x["target_rgbds"],
            x["depth_mask"], x["term_target"], x["term_mask"], lc["learning_rate"], lc["adam_eps"], lc["adam_weight_decay"])
        assert torch.allclose(pred.rgbds.detach(), x["out_rgbds"], atol=2e-5, rtol=1e-5)
        assert abs(losses["combined"].item() - x["loss"].item()) < 2e-5 * abs(x["loss"].item())
        for k, g in grads.items():
            if f"grad:{k}" not in x:
                assert g is None or not g.any(), k
                continue
            ref = x[f"grad:{k}"]
            assert (g - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-9, (it, k)
        # the update itself is pinned separately below on the reference's own gradients; along the trajectory the
        # parameters stay within the Adam step bound and almost all elements agree closely
        for k, v in state["params"].items():
            ref = x[f"param:{k}"]
            d = (v - ref).abs()
            assert d.max().item() <= 2.5 * lc["learning_rate"] * (it + 1), (it, k, d.max().item())
            assert (d > 1e-5).float().mean().item() < 0.02, (it, k)
        # continue from the reference's state so that sign flips of near-zero gradients do not accumulate
        state["params"] = {k: x[f"param:{k}"].clone() for k in state["params"]}
        for k, s in state["optim"].items():
            s["exp_avg"], s["exp_avg_sq"] = x[f"exp_avg:{k}"].clone(), x[f"exp_avg_sq:{k}"].clone()
            assert int(s["step"]) == int(x[f"step:{k}"].item()), (it, k)

    G.replay_training(meta, a, iteration, grow)


def test_adam_update_on_reference_gradients():
    """The Adam restatement alone: fed the reference's own gradients it reproduces the reference's parameter and
    moment tables (including rows of inactive fields, which must not move) to rounding."""
    meta, a = G.load("train_steps")
    lc = meta["loss_config"]
    state = {"params": {k: v.clone() for k, v in G.params(a, "init:param:").items()}}
    state["optim"] = T.new_optim_state(state["params"])

    def grow(rows):
        n = next(iter(rows.values())).shape[0]
        state["params"] = {k: torch.cat([v, rows[k]]) for k, v in state["params"].items()}
        state["optim"] = T.new_optim_state(state["params"], state["optim"], n)

    def iteration(it, x):
        grads = {k: x.get(f"grad:{k}") for k in state["params"]}
        T.adam_update(state["params"], state["optim"], x["field_ids"], grads, lc["learning_rate"], lc["adam_eps"],
                      lc["adam_weight_decay"])
        for k, v in state["params"].items():
            assert torch.allclose(v, x[f"param:{k}"], atol=1e-7, rtol=1e-6), (it, k)
            for key in ("exp_avg", "exp_avg_sq"):  # lerp cancels: absolute error of an ulp of the larger operand
                ref = x[f"{key}:{k}"]
                assert torch.allclose(state["optim"][k][key], ref, atol=3e-7 * ref.abs().max().item(), rtol=1e-5), (it, k, key)
            assert int(state["optim"][k]["step"]) == int(x[f"step:{k}"].item())

    G.replay_training(meta, a, iteration, grow)
    assert a["training_iterations"].tolist() == [3, 2, 4, 2, 2, 3]  # how often each field was active


def test_adam_update_vs_torch_optim_random_schedules():
    """The Adam restatement against torch.optim.Adam itself driven the way the reference drives it (gather rows and
    moments, step, scatter: ngm/run_mapping.py:679-707, 1191-1221) on random active sets, with tensors that skip
    iterations, weight decay on and off, and gradients spanning ten orders of magnitude."""
    for seed, wd, eps in ((0, 1e-5, 1e-15), (1, 0.0, 1e-8), (2, 1e-2, 1e-15)):
        g = torch.Generator().manual_seed(seed)
        shapes = {"w": (7, 5), "b": (7,), "s": ()}
        n = 9
        ours = {k: torch.randn(n, *s, generator=g) for k, s in shapes.items()}
        ref = {k: v.clone() for k, v in ours.items()}
        st = T.new_optim_state(ours)
        rst = {k: {"step": torch.tensor(0.0), "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
               for k, v in ref.items()}
        for it in range(6):
            ids = torch.randperm(n, generator=g)[: int(torch.randint(1, n + 1, (1,), generator=g))]
            grads = {k: torch.randn(len(ids), *s, generator=g) * 10.0 ** float(torch.randint(-8, 2, (1,), generator=g))
                     for k, s in shapes.items()}
            if it % 3 == 1:
                grads["s"] = None
            vm = {k: v[ids].clone().requires_grad_(True) for k, v in ref.items()}
            opt = torch.optim.Adam(list(vm.values()), lr=1e-3, eps=eps, weight_decay=wd)
            for k, p in vm.items():
                p.grad = grads[k]
                if rst[k]["step"].item() > 0:
                    opt.state[p] = {"step": rst[k]["step"], "exp_avg": rst[k]["exp_avg"][ids],
                                    "exp_avg_sq": rst[k]["exp_avg_sq"][ids]}
            opt.step()
            with torch.no_grad():
                for k, p in vm.items():
                    ref[k][ids] = p
                    if p in opt.state and len(opt.state[p]):
                        rst[k]["step"] = opt.state[p]["step"]
                        rst[k]["exp_avg"][ids] = opt.state[p]["exp_avg"]
                        rst[k]["exp_avg_sq"][ids] = opt.state[p]["exp_avg_sq"]
            T.adam_update(ours, st, ids, grads, 1e-3, eps, wd)
            for k in shapes:
                assert torch.allclose(ours[k], ref[k], atol=1e-7, rtol=2e-6), (seed, it, k)
                for key in ("exp_avg", "exp_avg_sq"):
                    r = rst[k][key]
                    assert torch.allclose(st[k][key], r, atol=3e-7 * r.abs().max().item() + 1e-30, rtol=1e-5), (seed, it, k, key)
                assert int(st[k]["step"]) == int(rst[k]["step"].item()), (seed, it, k)
