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
