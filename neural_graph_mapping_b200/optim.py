"""Training-step mirror: the optimizer half of the reference's mapping iteration.

The reference keeps ONE set of Adam moments for all fields (``NeuralGraphMap._optim_state``,
ngm/run_mapping.py:371-389) and, every iteration, gathers the active fields' parameters and moments into a
throw-away ``torch.optim.Adam`` (``_set_vmap_fields``, :668-707), steps it, and scatters parameters and moments
back (``_update_step``, :1183-1221).  ``ngm_adam_step`` does the same update in place on the full tables in one
launch, so the two methods shrink to what is left below; ``_optim_state`` keeps the reference's layout
(``{name: {"step", "exp_avg", "exp_avg_sq"}}``), so ``_add_fields`` (:365-389) and checkpoints are untouched.
"""
import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


def new_optim_state(all_fields_params: Dict[str, torch.Tensor], old_state: Optional[dict] = None) -> dict:
    """``_optim_state`` for the current ``all_fields_params`` (ngm/run_mapping.py:371-389): zero moments for rows
    that ``old_state`` does not cover (newly added fields), the shared step count carried over."""
    state = {}
    for name, p in all_fields_params.items():
        s = {"step": torch.tensor(0.0), "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        if old_state is not None and name in old_state:
            n_old = old_state[name]["exp_avg"].shape[0]
            s["step"] = old_state[name]["step"]
            s["exp_avg"][:n_old] = old_state[name]["exp_avg"]
            s["exp_avg_sq"][:n_old] = old_state[name]["exp_avg_sq"]
        state[name] = s
    return state


def _table(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"{name}: the CUDA Adam step needs a contiguous fp32 CUDA tensor "
                           f"(got {t.dtype} on {t.device}); there is no CPU path")
    return t


def adam_step(all_fields_params: Dict[str, torch.Tensor], vmap_fields_params: Dict[str, torch.Tensor],
              optim_state: dict, field_ids: Optional[torch.Tensor], lr: float, eps: float = 1e-8,
              weight_decay: float = 0.0, betas: Tuple[float, float] = (0.9, 0.999)) -> None:
    """One ``torch.optim.Adam`` step on rows ``field_ids`` of ``all_fields_params``, from the gradients held by
    ``vmap_fields_params`` (the gathered leaves the render ran on), in place; the updated rows are also written
    into ``vmap_fields_params``.  Parameters without a gradient are skipped, as torch does.  ``field_ids`` must be
    unique (the reference's scatter ``all[p][ids] = ...`` is equally undefined otherwise)."""
    work = []
    dev = None
    n_active = None
    for name, p in vmap_fields_params.items():
        g = p.grad
        if g is None:
            continue
        st = optim_state[name]
        full = _table(all_fields_params[name], f"all_fields_params[{name}]")
        m, v = _table(st["exp_avg"], f"exp_avg[{name}]"), _table(st["exp_avg_sq"], f"exp_avg_sq[{name}]")
        if m.shape != full.shape or v.shape != full.shape:
            raise ValueError(f"{name}: optimizer state {tuple(m.shape)} does not match the parameters {tuple(full.shape)}")
        g = _table(g.contiguous(), f"grad[{name}]")
        act = _table(p.detach(), f"vmap_fields_params[{name}]")
        if g.shape != act.shape or act.shape[1:] != full.shape[1:]:
            raise ValueError(f"{name}: gradient / active / full shapes disagree")
        if n_active is not None and act.shape[0] != n_active:
            raise ValueError(f"{name}: {act.shape[0]} active rows, other tensors have {n_active}")
        dev, n_active = full.device, act.shape[0]
        work.append((st, full, m, v, g, act))
    if not work:
        return
    ids = None
    if field_ids is not None:
        ids = field_ids.to(device=dev, dtype=torch.int64).contiguous()
        if ids.numel() != n_active:
            raise ValueError(f"{ids.numel()} field ids for {n_active} active rows")
    for b in betas:
        if not 0.0 <= b < 1.0:
            raise ValueError(f"betas outside [0, 1): {betas}")  # torch.optim.Adam's own check
    groups: Dict[int, list] = {}
    for st, full, m, v, g, act in work:  # everything validated: now count the step
        d = _lib.NgmAdamParam()
        d.param_all, d.exp_avg_all, d.exp_avg_sq_all = full.data_ptr(), m.data_ptr(), v.data_ptr()
        d.grad, d.param_active = g.data_ptr(), act.data_ptr()
        d.row = full[0].numel() if full.shape[0] else 0
        st["step"] += 1  # the reference's shared per-tensor step (:1213), a float tensor like torch's
        groups.setdefault(int(st["step"].item()), []).append(d)
    with torch.cuda.device(dev):
        for step, descs in groups.items():  # normally one group: every tensor has seen the same number of steps
            for i in range(0, len(descs), _lib.NGM_ADAM_MAX_PARAMS):
                chunk = descs[i:i + _lib.NGM_ADAM_MAX_PARAMS]
                arr = (_lib.NgmAdamParam * len(chunk))(*chunk)
                a = _lib.NgmAdamArgs()
                a.params, a.num_params = arr, len(chunk)
                a.field_ids, a.num_active, a.step = _lib.ptr(ids), n_active, step
                a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = lr, betas[0], betas[1], eps, weight_decay
                _lib.check(_lib.lib.ngm_adam_step(C.byref(a), _lib.stream_ptr(dev)))


def set_vmap_fields(driver, field_ids: torch.Tensor) -> None:
    """Drop-in for ``NeuralGraphMap._set_vmap_fields`` (ngm/run_mapping.py:668-707): gather the active rows and
    make them leaves; the optimizer-state gather is gone (the moments stay in the full tables)."""
    if getattr(driver, "_single_field_id", None) is not None:
        return
    with torch.no_grad():
        driver._model.set_vmap_fields(field_ids)
    for p in driver._model.vmap_fields_params.values():
        if p.is_floating_point():
            p.requires_grad_()


def adam_hyper(driver) -> dict:
    """Hyper-parameters of the driver's optimizer.  The reference builds ``torch.optim.Adam(lr=_learning_rate,
    eps=_adam_eps, weight_decay=_adam_weight_decay)`` with default betas (ngm/run_mapping.py:347-362); if the driver
    holds that object its (possibly re-scheduled) ``param_groups[0]`` wins, and anything but plain Adam is refused
    rather than silently replaced."""
    opt = getattr(driver, "_optimizer", None)
    if opt is not None:
        if type(opt) is not torch.optim.Adam:
            raise NotImplementedError(f"the CUDA update step implements torch.optim.Adam, not {type(opt).__name__}; "
                                      "use install(..., optimizer=False)")
        g = opt.param_groups[0]
        if g.get("amsgrad") or g.get("maximize"):
            raise NotImplementedError("amsgrad / maximize are not implemented by the CUDA update step")
        return dict(lr=float(g["lr"]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]),
                    betas=tuple(float(b) for b in g["betas"]))
    return dict(lr=float(driver._learning_rate), eps=float(driver._adam_eps),
                weight_decay=float(driver._adam_weight_decay), betas=(0.9, 0.999))


def update_step(driver, loss_dict: dict, field_ids: torch.Tensor) -> None:
    """Drop-in for ``NeuralGraphMap._update_step`` (ngm/run_mapping.py:1183-1221)."""
    if getattr(driver, "_single_field_id", None) is not None:
        raise NotImplementedError("single_field_id debugging mode is not implemented by the CUDA path")
    model = driver._model
    for p in model.vmap_fields_params.values():
        p.grad = None  # optimizer.zero_grad() (:1185)
    loss_dict["combined"].backward()  # :1186
    driver._global_map_dict["training_iterations"][field_ids] += 1  # :1188
    if driver._optim_state is None:
        driver._optim_state = new_optim_state(model.all_fields_params)
    adam_step(model.all_fields_params, model.vmap_fields_params, driver._optim_state, field_ids, **adam_hyper(driver))
    if hasattr(model, "repack_rows"):  # keep the persistent fp16 weight images in step with the rows just updated
        model.repack_rows(field_ids)
