"""Host logic of the training-step mirror (neural_graph_mapping_b200/optim.py) that needs no GPU: the
optimizer-state layout of the reference (_add_fields, ngm/run_mapping.py:371-389), the gather of _set_vmap_fields
(:668-680) and the loud failure without CUDA tensors."""
import pytest
import torch

from neural_graph_mapping_b200 import optim
from oracle import training as T


def _params(n, g):
    return {"_linears.0.weight": torch.randn(n, 8, 6, generator=g), "_linears.0.bias": torch.randn(n, 8, generator=g),
            "_neus_sd": torch.ones(n)}


def test_new_optim_state_layout_and_growth_match_oracle():
    g = torch.Generator().manual_seed(0)
    p = _params(3, g)
    st = optim.new_optim_state(p)
    ref = T.new_optim_state(p)
    assert st.keys() == ref.keys() == p.keys()
    for k in p:
        assert set(st[k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert st[k]["step"].item() == 0 and st[k]["exp_avg"].shape == p[k].shape and not st[k]["exp_avg"].any()
    # pretend two steps happened, then two fields are added
    for k in p:
        st[k]["step"] += 2
        st[k]["exp_avg"].copy_(torch.randn(p[k].shape, generator=g))
        st[k]["exp_avg_sq"].copy_(torch.rand(p[k].shape, generator=g))
        ref[k] = {"step": 2, "exp_avg": st[k]["exp_avg"].clone(), "exp_avg_sq": st[k]["exp_avg_sq"].clone()}
    grown = {k: torch.cat([v, torch.zeros(2, *v.shape[1:])]) for k, v in p.items()}
    st2, ref2 = optim.new_optim_state(grown, st), T.new_optim_state(grown, ref, 2)
    for k in p:
        assert st2[k]["step"] is st[k]["step"] and int(st2[k]["step"].item()) == ref2[k]["step"] == 2
        assert torch.equal(st2[k]["exp_avg"], ref2[k]["exp_avg"]) and torch.equal(st2[k]["exp_avg_sq"], ref2[k]["exp_avg_sq"])
        assert not st2[k]["exp_avg"][3:].any() and not st2[k]["exp_avg_sq"][3:].any()


class _Model:
    def __init__(self, p):
        self.all_fields_params, self.vmap_fields_params = p, None

    def set_vmap_fields(self, ids):  # ngm/models.py:266-276
        self.vmap_fields_params = {k: v[ids] for k, v in self.all_fields_params.items()}


class _Driver:
    _single_field_id = None

    def __init__(self, p):
        self._model = _Model(p)


def test_set_vmap_fields_gathers_leaves():
    p = _params(5, torch.Generator().manual_seed(1))
    for v in p.values():
        v.requires_grad_(True)  # the gather must not record a graph into the full tables
    d = _Driver(p)
    ids = torch.tensor([3, 0])
    optim.set_vmap_fields(d, ids)
    for k, v in d._model.vmap_fields_params.items():
        assert v.is_leaf and v.requires_grad and torch.equal(v.detach(), p[k].detach()[ids])


def test_adam_step_has_no_cpu_path():
    p = _params(4, torch.Generator().manual_seed(2))
    v = {k: t[:2].clone().requires_grad_(True) for k, t in p.items()}
    for t in v.values():
        t.grad = torch.ones_like(t)
    with pytest.raises(RuntimeError, match="no CPU path"):
        optim.adam_step(p, v, optim.new_optim_state(p), torch.tensor([0, 1]), 1e-3)
    # nothing has a gradient: nothing to do, not an error (torch.optim.Adam.step() with no grads)
    for t in v.values():
        t.grad = None
    st = optim.new_optim_state(p)
    optim.adam_step(p, v, st, torch.tensor([0, 1]), 1e-3)
    assert all(s["step"].item() == 0 for s in st.values())


def test_adam_hyper_follows_the_drivers_optimizer():
    d = _Driver(_params(2, torch.Generator().manual_seed(3)))
    d._learning_rate, d._adam_eps, d._adam_weight_decay = 1e-3, 1e-15, 1e-5
    assert optim.adam_hyper(d) == dict(lr=1e-3, eps=1e-15, weight_decay=1e-5, betas=(0.9, 0.999))
    # the reference's own throw-away optimizer object (run_mapping.py:357-362), re-scheduled by the user
    d._optimizer = torch.optim.Adam([torch.zeros((), requires_grad=True)], lr=1e-3, eps=1e-15, weight_decay=1e-5)
    d._optimizer.param_groups[0]["lr"] = 2.5e-4
    assert optim.adam_hyper(d) == dict(lr=2.5e-4, eps=1e-15, weight_decay=1e-5, betas=(0.9, 0.999))
    d._optimizer = torch.optim.RMSprop([torch.zeros((), requires_grad=True)], lr=1e-3)
    with pytest.raises(NotImplementedError, match="RMSprop"):
        optim.adam_hyper(d)
    d._optimizer = torch.optim.Adam([torch.zeros((), requires_grad=True)], amsgrad=True)
    with pytest.raises(NotImplementedError, match="amsgrad"):
        optim.adam_hyper(d)


def test_update_step_host_sequence(monkeypatch):
    """zero_grad -> backward -> training_iterations -> one adam_step call with the driver's hyper-parameters
    (ngm/run_mapping.py:1183-1193); the kernel call itself is covered by the GPU tests."""
    p = _params(4, torch.Generator().manual_seed(4))
    d = _Driver(p)
    d._learning_rate, d._adam_eps, d._adam_weight_decay, d._optim_state = 1e-3, 1e-15, 1e-5, None
    d._global_map_dict = {"training_iterations": torch.zeros(4, dtype=torch.long)}
    ids = torch.tensor([2, 0])
    optim.set_vmap_fields(d, ids)
    vm = d._model.vmap_fields_params
    vm["_neus_sd"].grad = torch.ones(2)  # stale gradient from an earlier iteration: must be cleared
    calls = []
    monkeypatch.setattr(optim, "adam_step", lambda *a, **k: calls.append((a, k)))
    loss = (vm["_linears.0.weight"] ** 2).sum() + vm["_linears.0.bias"].sum()
    optim.update_step(d, {"combined": loss}, ids)
    assert d._global_map_dict["training_iterations"].tolist() == [1, 0, 1, 0]
    assert vm["_neus_sd"].grad is None and torch.equal(vm["_linears.0.bias"].grad, torch.ones(2, 8))
    assert torch.allclose(vm["_linears.0.weight"].grad, 2 * p["_linears.0.weight"][ids])
    (a, k), = calls
    assert a[0] is p and a[1] is vm and a[2] is d._optim_state and a[3] is ids
    assert k == dict(lr=1e-3, eps=1e-15, weight_decay=1e-5, betas=(0.9, 0.999))
    assert set(d._optim_state) == set(p)  # created on first use in the reference's layout
