"""Training step on the GPU: ngm_adam_step (Adam on the active rows, in place) and the whole mapping iteration
(render under autograd -> losses -> backward -> Adam) against the golden trajectory of the unmodified reference
(tests/golden/train_steps.npz; ngm/run_mapping.py:668-707, 1164-1221)."""
import pytest
import torch

import golden_util as G
from oracle import training as T
from tests_support import make_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close_moments(ours, ref, what):
    assert torch.allclose(ours.cpu(), ref, atol=3e-7 * ref.abs().max().item(), rtol=1e-5), what


def test_adam_step_on_reference_gradients():
    """The kernel alone, fed the reference's own gradients: parameters, both moments and the step counts of every
    iteration (inactive rows untouched, growth in between) to rounding."""
    from neural_graph_mapping_b200 import optim

    meta, a = G.load("train_steps")
    lc = meta["loss_config"]
    state = {"params": {k: v.to(DEV) for k, v in G.params(a, "init:param:").items()}}
    state["optim"] = optim.new_optim_state(state["params"])

    def grow(rows):
        state["params"] = {k: torch.cat([v, rows[k].to(DEV)]) for k, v in state["params"].items()}
        state["optim"] = optim.new_optim_state(state["params"], state["optim"])

    def iteration(it, x):
        ids = x["field_ids"].to(DEV)
        vmap = {k: v[ids].clone().requires_grad_(True) for k, v in state["params"].items()}
        for k, v in vmap.items():
            v.grad = x[f"grad:{k}"].to(DEV) if f"grad:{k}" in x else None
        optim.adam_step(state["params"], vmap, state["optim"], ids, lc["learning_rate"], lc["adam_eps"],
                        lc["adam_weight_decay"])
        for k, v in state["params"].items():
            ref = x[f"param:{k}"]
            assert torch.allclose(v.cpu(), ref, atol=1e-7, rtol=1e-6), (it, k)
            assert torch.equal(vmap[k].detach(), v[ids]), (it, k)  # the active copy holds the updated rows
            _close_moments(state["optim"][k]["exp_avg"], x[f"exp_avg:{k}"], (it, k))
            _close_moments(state["optim"][k]["exp_avg_sq"], x[f"exp_avg_sq:{k}"], (it, k))
            assert int(state["optim"][k]["step"].item()) == int(x[f"step:{k}"].item()), (it, k)

    G.replay_training(meta, a, iteration, grow)


def test_training_iterations_golden():
    """The drop-in flow of the patched driver: _render_ijs (calls _set_vmap_fields, :500) -> _compute_losses
    (the reference's own, restated in oracle/training.py) -> _update_step, iteration by iteration."""
    from neural_graph_mapping_b200 import optim

    meta, a = G.load("train_steps")
    import neural_graph_mapping_b200 as ngm

    lc = meta["loss_config"]
    ls = T.LossSpec(**{k: lc[k] for k in T.LossSpec.__dataclass_fields__})
    a0 = dict(a)
    a0.update({"param:" + k: v for k, v in G.params(a, "init:param:").items()})
    meta0 = dict(meta)
    meta0["config"] = dict(meta["config"], learning_rate=lc["learning_rate"], adam_eps=lc["adam_eps"],
                           adam_weight_decay=lc["adam_weight_decay"])
    st = make_state(meta0, {k: v for k, v in a0.items() if not k.startswith(("it", "grow", "init"))}, DEV)
    st._reference_flow = True
    st._optim_state = optim.new_optim_state(st._model.all_fields_params)
    cam = ngm.Camera(**meta["camera"])

    def grow(rows):
        ap = st._model.all_fields_params
        st._model.all_fields_params = {k: torch.cat([v, rows[k].to(DEV)]) for k, v in ap.items()}
        st._optim_state = optim.new_optim_state(st._model.all_fields_params, st._optim_state)

    def iteration(it, x):
        d = lambda k: x[k].to(DEV)  # noqa: E731
        pred = st._render_ijs(d("ijs"), d("c2ws"), cam, d("field_ids"), True, d("near"), d("far"), d("gt"),
                              jitter=d("jitter"), jitter_guided=d("jitter_guided"))
        assert torch.allclose(pred.rgbds.detach().cpu(), x["out_rgbds"], atol=1e-4, rtol=1e-4)
        losses = T.compute_losses(ls, pred, d("target_rgbds"), d("depth_mask"), d("term_target"), d("term_mask"))
        assert abs(losses["combined"].item() - x["loss"].item()) < 1e-4 * abs(x["loss"].item())
        st._update_step(losses, d("field_ids"))
        for k, p in st._model.vmap_fields_params.items():
            if f"grad:{k}" not in x:
                continue
            ref = x[f"grad:{k}"]
            err = (p.grad.cpu() - ref).abs().max().item()
            assert err <= 2e-4 * ref.abs().max().item() + 1e-8, (it, k, err)
        for k, v in st._model.all_fields_params.items():
            dd = (v.cpu() - x[f"param:{k}"]).abs()
            # Adam with eps = 1e-15 moves every element by ~lr whatever the gradient's size, so an element whose
            # gradient is rounding noise may step the other way; everything else agrees closely
            assert dd.max().item() <= 2.5 * lc["learning_rate"], (it, k, dd.max().item())
            assert (dd > 2e-5).float().mean().item() < 0.02, (it, k, (dd > 2e-5).float().mean().item())
        # continue from the reference's tables so that those sign flips do not accumulate
        st._model.all_fields_params = {k: x[f"param:{k}"].to(DEV) for k in st._model.all_fields_params}
        for k, s in st._optim_state.items():
            s["exp_avg"], s["exp_avg_sq"] = x[f"exp_avg:{k}"].to(DEV), x[f"exp_avg_sq:{k}"].to(DEV)

    G.replay_training(meta, a, iteration, grow)
    assert st._global_map_dict["training_iterations"].tolist() == a["training_iterations"].tolist()


@pytest.mark.parametrize("F_all,F_act,wd", [(40, 32, 1e-5), (5, 5, 0.0), (3, 0, 0.0)])
def test_adam_step_vs_torch_adam_large(F_all, F_act, wd):
    """4x128 / NeRF-8 parameter shapes (rows of 6,144 ... 16,384 ... 4 elements, not multiples of the kernel's
    chunk) for several steps against the reference's own sequence on the GPU: gather -> torch.optim.Adam ->
    scatter (ngm/run_mapping.py:679-707, 1191-1221); a tensor without a gradient keeps its step count."""
    from neural_graph_mapping_b200 import optim

    g = torch.Generator().manual_seed(F_all)
    shapes = {"_linears.0.weight": (128, 48), "_linears.0.bias": (128,), "_linears.1.weight": (128, 128),
              "_linears.1.bias": (128,), "_linears.4.weight": (4, 128), "_linears.4.bias": (4,), "_neus_sd": ()}
    ours = {k: torch.randn(F_all, *s, generator=g).to(DEV) for k, s in shapes.items()}
    ref = {k: v.clone() for k, v in ours.items()}
    state = optim.new_optim_state(ours)
    ref_state = {k: {"step": torch.tensor(0.0), "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
                 for k, v in ref.items()}
    lr, eps = 1e-3, 1e-15
    for it in range(4):
        ids = torch.randperm(F_all, generator=g)[:F_act].to(DEV)
        grads = {k: (torch.randn(F_act, *s, generator=g) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))).to(DEV)
                 for k, s in shapes.items()}
        if it % 2 == 1:
            grads["_neus_sd"] = None  # did not take part this iteration
        # reference sequence
        vm = {k: v[ids].clone().requires_grad_(True) for k, v in ref.items()}
        opt = torch.optim.Adam(list(vm.values()), lr=lr, eps=eps, weight_decay=wd)
        for k, p in vm.items():
            p.grad = grads[k]
            if ref_state[k]["step"].item() > 0:
                opt.state[p] = {"step": ref_state[k]["step"], "exp_avg": ref_state[k]["exp_avg"][ids],
                                "exp_avg_sq": ref_state[k]["exp_avg_sq"][ids]}
        if F_act:
            opt.step()
        with torch.no_grad():
            for k, p in vm.items():
                ref[k][ids] = p
                if p in opt.state and len(opt.state[p]):
                    ref_state[k]["step"] = opt.state[p]["step"]
                    ref_state[k]["exp_avg"][ids] = opt.state[p]["exp_avg"]
                    ref_state[k]["exp_avg_sq"][ids] = opt.state[p]["exp_avg_sq"]
        # ours
        vo = {k: v[ids].clone().requires_grad_(True) for k, v in ours.items()}
        for k, p in vo.items():
            p.grad = grads[k]
        optim.adam_step(ours, vo, state, ids, lr, eps, wd)
        for k in shapes:
            assert torch.allclose(ours[k], ref[k], atol=1e-7, rtol=2e-6), (it, k, (ours[k] - ref[k]).abs().max().item())
            for key in ("exp_avg", "exp_avg_sq"):
                r = ref_state[k][key]
                assert torch.allclose(state[k][key], r, atol=3e-7 * r.abs().max().item() + 1e-30, rtol=1e-5), (it, k, key)
            if F_act:
                assert int(state[k]["step"].item()) == int(ref_state[k]["step"].item()), (it, k)


def test_adam_step_errors():
    from neural_graph_mapping_b200 import optim

    p = {"w": torch.zeros(4, 8, device=DEV)}
    st = optim.new_optim_state(p)
    v = {"w": p["w"][:2].clone().requires_grad_(True)}
    v["w"].grad = torch.ones(2, 8, device=DEV)
    with pytest.raises(ValueError, match="field ids"):
        optim.adam_step(p, v, st, torch.tensor([0, 1, 2], device=DEV), 1e-3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        optim.adam_step({"w": torch.zeros(4, 8)}, v, optim.new_optim_state({"w": torch.zeros(4, 8)}), None, 1e-3)
    with pytest.raises(ValueError, match="betas"):
        optim.adam_step(p, v, st, torch.tensor([0, 1], device=DEV), 1e-3, betas=(1.0, 0.999))
