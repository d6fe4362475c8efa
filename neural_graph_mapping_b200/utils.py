"""Host-side helpers of the render path.

``str_to_object`` keeps the contract of the reference's plugin mechanism (ngm/utils.py:114-138): a
YAML type string names either something visible in the caller's scope or a fully qualified
``package.module.Attr``; unknown names give ``None`` (the callers raise their own ``ValueError``).
``batched_evaluation`` keeps the contract of ngm/utils.py:220-251 (evaluate along dim 0 in blocks, return
what one call over all inputs would have returned) but writes every block into output buffers allocated once,
so a frame rendered in blocks never holds two copies of its outputs.
"""
from __future__ import annotations

import importlib
import sys
from typing import Any, Callable, Optional

import torch


def str_to_object(name: str) -> Any:
    """Resolve a type string: caller scope first, then ``importlib`` on the dotted path."""
    scope = sys._getframe(1)
    for table in (scope.f_locals, scope.f_globals):
        if name in table:
            return table[name]
    parts = name.split(".")
    for split in range(len(parts), 0, -1):  # longest importable module prefix, then attribute walk
        try:
            obj = importlib.import_module(".".join(parts[:split]))
        except ImportError:
            continue
        try:
            for attr in parts[split:]:
                obj = getattr(obj, attr)
        except AttributeError:
            return None
        return obj
    return None


def _alloc_like(block: torch.Tensor, rows: int) -> torch.Tensor:
    return torch.empty((rows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)


def batched_evaluation(model: Callable, inputs: torch.Tensor, block_size: int, progressbar: bool = False):
    """Evaluate ``model`` over ``inputs`` in blocks of ``block_size`` rows.

    Tensor members whose leading dimension equals the block's row count are written straight into
    preallocated full-size outputs; members of data-dependent length (e.g. the masked free-space /
    TSDF tensors of a ``Prediction``) are collected and concatenated; non-tensor members are returned
    as the tuple of per-block values.
    """
    total = int(inputs.shape[0])
    starts = range(0, total, int(block_size))
    if progressbar:
        from tqdm import tqdm

        starts = tqdm(starts)
    dense: Optional[list] = None   # per member: full-size buffer, or None
    ragged: Optional[list] = None  # per member: list of per-block values
    single = False
    for begin in starts:
        stop = min(begin + int(block_size), total)
        result = model(inputs[begin:stop])
        single = not isinstance(result, tuple)
        members = (result,) if single else tuple(result)
        if dense is None:
            dense = [(_alloc_like(m, total) if torch.is_tensor(m) and m.dim() > 0 and m.shape[0] == stop - begin
                      else None) for m in members]
            ragged = [[] for _ in members]
        for i, m in enumerate(members):
            if dense[i] is not None and torch.is_tensor(m) and m.shape[0] == stop - begin:
                dense[i][begin:stop] = m
            else:
                if dense[i] is not None:  # a later block broke the pattern: fall back to collecting
                    ragged[i] = [dense[i][:begin]] if begin else []
                    dense[i] = None
                ragged[i].append(m)
    if dense is None:
        return []
    merged = []
    for buf, parts in zip(dense, ragged):
        if buf is not None:
            merged.append(buf)
        elif parts and torch.is_tensor(parts[0]):
            merged.append(torch.cat(parts))
        else:
            merged.append(tuple(parts))
    return merged[0] if single else tuple(merged)
