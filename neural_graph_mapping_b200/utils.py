"""Small host-side helpers restated from ngm/utils.py (the parts on the render path)."""
from __future__ import annotations

import inspect
from pydoc import locate
from typing import Any, Callable

import torch


def str_to_object(name: str) -> Any:
    """Resolve a (fully qualified) name to an object -- the reference's plugin mechanism
    (ngm/utils.py:114-138): caller locals, caller globals, then import."""
    caller = inspect.currentframe().f_back
    if name in caller.f_locals:
        return caller.f_locals[name]
    if name in caller.f_globals:
        return caller.f_globals[name]
    return locate(name)


def batched_evaluation(model: Callable, inputs: torch.Tensor, block_size: int, progressbar: bool = False):
    """Evaluate in blocks along dim 0 and concatenate tuple-wise (ngm/utils.py:220-251)."""
    outs = []
    iterator = range(0, inputs.shape[0], block_size)
    if progressbar:
        from tqdm import tqdm

        iterator = tqdm(iterator)
    for start in iterator:
        end = min(start + block_size, inputs.shape[0])
        outs.append(model(inputs[start:end]))
    if isinstance(outs[0], tuple):
        outs = tuple(torch.cat(x) if isinstance(x[0], torch.Tensor) else x for x in zip(*outs))
    elif isinstance(outs[0], torch.Tensor):
        outs = torch.cat(outs)
    return outs
