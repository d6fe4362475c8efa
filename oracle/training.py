"""Plain-PyTorch (CPU, fp32) restatement of the reference's training iteration around the render path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``ngm/`` abbreviates
``/root/reference/src/neural_graph_mapping/``.  Pinned by ``tests/golden/train_steps.npz``, which
``oracle/make_training_fixture.py`` produced by running the unmodified reference (its own ``torch.optim.Adam``
included) on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import restatement as R


@dataclass
class LossSpec:
    """Loss configuration read by ``_compute_losses`` (ngm/run_mapping.py:160-175; neural_graph_map.yaml:19-33)."""

    termination_weight: float = 0.0
    photometric_weight: float = 1.0
    photometric_loss: str = "l1"
    depth_weight: float = 1.0
    depth_loss: str = "huber"
    freespace_weight: float = 40.0
    tsdf_weight: float = 50.0
    truncation_distance: float = 0.1


def photometric_loss(mode, measured, rendered, rendered_vars=None):
    """ngm/losses.py:10-41.  (The argument order at the call site, run_mapping.py:1820-1825, passes the prediction
    as ``measured`` and the target as ``rendered``; l1 / l2 are symmetric, gaussian_nll is not.)"""
    if mode == "l1":
        return torch.mean(torch.abs(measured - rendered))
    if mode == "l2":
        return torch.mean((measured - rendered) ** 2)
    if mode == "gaussian_nll":
        nlls = 0.5 * (rendered - measured) ** 2 / rendered_vars + torch.log(torch.sqrt(rendered_vars))
        loss = nlls.mean()
        return torch.mean(torch.abs(measured - rendered)) if loss > 2 else loss
    raise ValueError(mode)


def depth_loss(mode, measured, rendered, rendered_vars=None):
    """ngm/losses.py:44-78."""
    if mode == "huber":
        return torch.nn.functional.huber_loss(rendered, measured, delta=0.05)
    if mode == "gaussian_nll":
        v = rendered_vars + 1e-15
        return (0.5 * (rendered - measured) ** 2 / v + torch.log(torch.sqrt(v))).mean()
    if mode == "laplacian_nll":
        return (torch.abs(measured - rendered) / torch.sqrt(0.5 * rendered_vars + 1e-6)
                + 0.5 * torch.log(2 * rendered_vars + 1e-6)).mean()
    raise ValueError(mode)


def compute_losses(spec: LossSpec, prediction, target_rgbds, target_depth_mask, target_term_probs,
                   target_term_mask) -> Dict[str, torch.Tensor]:
    """``NeuralGraphMap._compute_losses`` (ngm/run_mapping.py:1769-1871), multi-field branch."""
    depth_mask = target_depth_mask * (prediction.term_probs > 0.8)  # :1787
    rgb_mask = depth_mask  # :1788
    out = {}
    term = ((prediction.term_probs[target_term_mask] - target_term_probs[target_term_mask]) ** 2).mean()  # :1803-1806
    combined = 0 + spec.termination_weight * term
    out["termination"] = term
    photo = photometric_loss(spec.photometric_loss, prediction.rgbds[rgb_mask][:, :3], target_rgbds[rgb_mask][:, :3],
                             prediction.color_vars[rgb_mask])  # :1820-1825
    combined = combined + spec.photometric_weight * photo
    out["photometric"] = photo
    dl = depth_loss(spec.depth_loss, target_rgbds[depth_mask][:, 3], prediction.rgbds[depth_mask][:, 3],
                    prediction.depth_vars[depth_mask])  # :1830-1835
    combined = combined + spec.depth_weight * dl
    out["depth"] = dl
    if prediction.freespace_geometry is not None:  # :1842-1846
        fs = ((prediction.freespace_geometry - spec.truncation_distance) ** 2).mean()
        combined = combined + spec.freespace_weight * fs
        out["freespace"] = fs
    if prediction.tsdf_residuals is not None:  # :1848-1851
        ts = (prediction.tsdf_residuals ** 2).mean()
        combined = combined + spec.tsdf_weight * ts
        out["tsdf"] = ts
    out["combined"] = combined
    return out


def new_optim_state(all_params: Dict[str, torch.Tensor], old_state: Optional[dict] = None, num_new: int = 0) -> dict:
    """The optimizer-state half of ``_add_fields`` (ngm/run_mapping.py:371-389)."""
    state = {}
    for p, v in all_params.items():
        state[p] = {"step": 0, "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
        if old_state is not None:
            state[p]["step"] = old_state[p]["step"]
            state[p]["exp_avg"][: v.shape[0] - num_new] = old_state[p]["exp_avg"]
            state[p]["exp_avg_sq"][: v.shape[0] - num_new] = old_state[p]["exp_avg_sq"]
    return state


def adam_update(all_params: Dict[str, torch.Tensor], optim_state: dict, field_ids: torch.Tensor,
                grads: Dict[str, Optional[torch.Tensor]], lr: float, eps: float, weight_decay: float,
                betas=(0.9, 0.999)) -> None:
    """``_set_vmap_fields`` + ``optimizer.step()`` + the scatter of ``_update_step`` (ngm/run_mapping.py:679-707,
    1191-1221) as explicit arithmetic: gather rows and moments, torch.optim.Adam's update (plain Adam, weight decay
    added to the gradient, bias corrections in double), scatter back.  Tensors without a gradient are skipped and
    keep their step count, as torch does."""
    b1, b2 = betas
    with torch.no_grad():
        for name, g in grads.items():
            if g is None:
                continue
            st = optim_state[name]
            w = all_params[name][field_ids]
            m, v = st["exp_avg"][field_ids], st["exp_avg_sq"][field_ids]
            st["step"] = int(st["step"]) + 1
            t = st["step"]
            if weight_decay != 0:
                g = g + weight_decay * w
            m = m + (g - m) * (1 - b1)
            v = v * b2 + (1 - b2) * g * g
            bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
            denom = v.sqrt() / (bc2 ** 0.5) + eps
            w = w - (lr / bc1) * (m / denom)
            all_params[name][field_ids] = w
            st["exp_avg"][field_ids] = m
            st["exp_avg_sq"][field_ids] = v


def training_iteration(all_params, optim_state, positions, orientations, field_ids, cam, rspec, fspec, lspec: LossSpec,
                       ijs, c2ws, near, far, gt, jitter, jitter_guided, target_rgbds, depth_mask, term_target, term_mask,
                       lr, eps, weight_decay):
    """One mapping iteration (ngm/run_mapping.py:1164-1221): render the active fields under autograd, losses,
    backward, Adam on the active rows.  Returns (losses, gradients of the gathered parameters, prediction)."""
    vmap = {k: v[field_ids].clone().requires_grad_(v.dtype.is_floating_point) for k, v in all_params.items()}
    ar = torch.arange(len(field_ids))
    pred = R.render_rays(ijs, c2ws, cam, rspec, fspec, vmap, positions[field_ids], orientations[field_ids], field_ids=ar,
                         use_vmap=True, near_distances=near.clone(), far_distances=far.clone(), gt_distances=gt.clone(),
                         jitter=jitter, jitter_guided=jitter_guided)
    losses = compute_losses(lspec, pred, target_rgbds, depth_mask, term_target, term_mask)
    losses["combined"].backward()
    grads = {k: v.grad for k, v in vmap.items()}
    adam_update(all_params, optim_state, field_ids, grads, lr, eps, weight_decay)
    return losses, grads, pred
