"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference is single-GPU (no collective anywhere).  Rays never interact and fields interact
only read-only, so the path shards with NO data-path collective; the one exchange step is a
single all-gather of the rendered tiles (9 fp32 = 36 B per ray: rgbd 4, colour var 3, depth var
1, term prob 1) so every rank ends with the full Prediction.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .renderer import Prediction, render_rays

FLOATS_PER_RAY = 9


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def world_info(group=None) -> Tuple[int, int]:
    d = _dist()
    return (d.get_world_size(group), d.get_rank(group)) if d else (1, 0)


def shard_range(num_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``num_items`` fields (or pixel rows) over ranks."""
    base, rem = divmod(num_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def owned_fields(field_ids: torch.Tensor, num_fields: int, world: int, rank: int) -> torch.Tensor:
    """Training at N > 1: rank r owns the contiguous rows ``shard_range(num_fields, world, r)`` of
    the per-field tables (parameters AND Adam moments live only there), so of a global set of target fields it
    renders, back-propagates and updates exactly the ones it owns -- the gradient of a field is local to its
    rays, hence no all-reduce.  Returns the boolean mask of ``field_ids`` this rank owns."""
    f0, f1 = shard_range(num_fields, world, rank)
    return (field_ids >= f0) & (field_ids < f1)


def packed_views(flat: torch.Tensor, n_rays: int):
    """Views of one rank's packed tile [rgbd 4n | colour var 3n | depth var n | term n]."""
    assert flat.numel() == FLOATS_PER_RAY * n_rays
    o = 0
    rgbd = flat[o:o + 4 * n_rays].view(n_rays, 4); o += 4 * n_rays
    cvar = flat[o:o + 3 * n_rays].view(n_rays, 3); o += 3 * n_rays
    dvar = flat[o:o + n_rays]; o += n_rays
    term = flat[o:o + n_rays]
    return rgbd, cvar, dvar, term


def gather_tiles(local_flat: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally sized packed tiles -> (world, len(local_flat))."""
    d = _dist()
    world, _ = world_info(group)
    if d is None or world == 1:
        return local_flat.view(1, -1)
    out = torch.empty(world, local_flat.numel(), device=local_flat.device, dtype=local_flat.dtype)
    d.all_gather_into_tensor(out.view(-1), local_flat.contiguous(), group=group)
    return out


def render_rays_gathered(driver, ijs, c2ws, camera, field_ids, near=None, far=None, gt=None,
                         return_packed: bool = False, group=None, **kw):
    """Each rank renders ITS OWN (F, R) ray batch (e.g. its keyframe / its pixel tiles); the packed
    tiles of all ranks are all-gathered.  Returns the packed (world, 9*F*R) buffer or a list of
    per-rank ``Prediction`` views into it."""
    F, R = ijs.shape[0], ijs.shape[1]
    n = F * R
    local = torch.empty(FLOATS_PER_RAY * n, device=ijs.device, dtype=torch.float32)
    rgbd, cvar, dvar, term = packed_views(local, n)
    render_rays(driver, ijs, c2ws, camera, field_ids, True, near, far, gt,
                out=(rgbd.view(F, R, 4), cvar.view(F, R, 3), dvar.view(F, R), term.view(F, R)), **kw)
    buf = gather_tiles(local, group)
    if return_packed:
        return buf
    preds = []
    for r in range(buf.shape[0]):
        a, b, c, d = packed_views(buf[r], n)
        preds.append(Prediction(a.view(F, R, 4), b.view(F, R, 3), c.view(F, R), d.view(F, R), None, None))
    return preds


def render_rays_sharded(driver, ijs, c2ws, camera, field_ids, near=None, far=None, gt=None, group=None, **kw):
    """Training-batch shape: the GLOBAL (F, R) batch is known on every rank;
    rank r renders fields [f0, f1) -- so it only ever touches its own fields' parameters -- and
    the tiles are all-gathered into the full (F, R) Prediction on every rank."""
    world, rank = world_info(group)
    F, R = ijs.shape[0], ijs.shape[1]
    if F % world != 0:
        raise ValueError(f"num_fields={F} must be divisible by the world size {world}")
    f0, f1 = shard_range(F, world, rank)

    def sl(x, per_ray=True):
        if x is None:
            return None
        return x[f0:f1] if (per_ray and torch.is_tensor(x) and x.dim() >= 2 and x.shape[0] == F) else x

    c2 = c2ws if c2ws.dim() == 2 else c2ws[f0:f1]
    buf = render_rays_gathered(driver, ijs[f0:f1], c2, camera, field_ids[f0:f1], sl(near), sl(far), sl(gt),
                               return_packed=True, group=group, **kw)
    n = (f1 - f0) * R
    parts = [packed_views(buf[r], n) for r in range(buf.shape[0])]
    cat = lambda i, shape: torch.cat([p[i] for p in parts]).view(*shape)  # noqa: E731
    return Prediction(cat(0, (F, R, 4)), cat(1, (F, R, 3)), cat(2, (F, R)), cat(3, (F, R)), None, None)
