"""CPU restatement of the permutohedral-lattice multi-resolution hash encoding.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

**PARITY UNPINNED.**  The reference delegates this encoding to the third-party
package ``permutohedral_encoding`` (fork ``roym899/permutohedral_encoding`` @
``bf445adb4b3aa77eeff42d3443bf04b6aad8338b``, ``pyproject.toml:20``); its source is
NOT under ``/root/reference`` and it cannot be installed offline.  The only call
sites are ``ngm/positional_encodings.py:6,19,52-62,66`` (wrapper kwargs, width
``nr_levels*nr_feat_per_level (+pos_dim)``) and ``ngm/models.py:145``.  There are no
tests, golden vectors or fixtures for it in the reference.  What follows restates the
*published* algorithm (Adams et al. 2010 permutohedral lattice; as used by PermutoSDF,
Rosu & Behnke 2023, the upstream of the pinned fork):

per point, per level ``l``:
  1. ``cf_i = (x_i + shift[l][i]) * scale[l][i]``,  ``scale[l][i] = 1/(sqrt((i+1)(i+2)) * sigma_l)``
  2. elevate to the hyperplane ``sum = 0`` in ``d+1`` dims:
     ``E_d = -d*cf_{d-1}``, ``E_i = sum_{j>=i} cf_j - i*cf_{i-1}``, ``E_0 = sum_j cf_j``
  3. nearest remainder-0 point ``rem0_i = round(E_i/(d+1))*(d+1)`` (ties down),
     ``s = sum(rem0)/(d+1)``; ``rank_i`` = number of coordinates with a larger residual
     ``E - rem0`` (ties: earlier index ranks higher); fix-up ``rank += s`` with wrap.
  4. barycentric weights ``b[d-rank_i] += delta_i``, ``b[d+1-rank_i] -= delta_i``,
     ``delta_i = (E_i - rem0_i)/(d+1)``; ``b[0] += 1 + b[d+1]``.
  5. for ``r = 0..d``: key_i = rem0_i + r - (d+1)*[rank_i > d-r]  (first d coords),
     hash ``h = ((..((0+key_0)*P + key_1)*P ..)+key_{d-1})*P  mod 2^32``, ``P = 2531011``,
     slot ``h mod capacity``; out[l] += b[r] * table[l][slot].
Output layout: level-major, feature-minor, then (optionally) the raw points * scaling.

The CUDA kernel is checked bit-for-bit-class (fp32 tolerance) against THIS file, and
this file is cross-checked against an independent scalar transcription in
``tests/test_oracle_permuto.py``.  Compatibility with checkpoints trained through the
upstream extension is NOT claimed.
"""
from __future__ import annotations

import math

import numpy as np
import torch

HASH_PRIME = 2531011


def scale_factors(kwargs: dict) -> torch.Tensor:
    """(nr_levels, pos_dim) fp32.  ``sigma_l = geomspace(coarsest, finest, nr_levels)``
    (ngm/positional_encodings.py:50)."""
    d = kwargs.get("pos_dim", 3)
    sig = np.geomspace(kwargs["coarsest_scale"], kwargs["finest_scale"], num=kwargs["nr_levels"])
    sf = np.zeros((kwargs["nr_levels"], d), dtype=np.float64)
    for l, s in enumerate(sig):
        for i in range(d):
            sf[l, i] = 1.0 / (math.sqrt((i + 1) * (i + 2)) * s)
    return torch.from_numpy(sf.astype(np.float32))


def init_params(kwargs: dict, generator: torch.Generator) -> dict:
    d = kwargs.get("pos_dim", 3)
    cap = 2 ** kwargs["log2_hashmap_size"]
    init_scale = kwargs.get("init_scale", 1e-5)
    table = (torch.rand(kwargs["nr_levels"], cap, kwargs["nr_feat_per_level"], generator=generator)
             * 2 - 1) * init_scale
    if kwargs.get("appply_random_shift_per_level", True):  # (sic) wrapper spelling, :30
        shift = torch.randn(kwargs["nr_levels"], d, generator=generator) * 10.0
    else:
        shift = torch.zeros(kwargs["nr_levels"], d)
    return {"_encoding.lattice_values": table, "_encoding.random_shift_per_level": shift}


def encode(points: torch.Tensor, table: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor,
           concat_points: bool = False, concat_points_scaling: float = 1.0) -> torch.Tensor:
    """points (...,d) fp32; table (L,cap,feat); shift (L,d); scale (L,d) -> (..., L*feat [+d])."""
    leading = points.shape[:-1]
    x = points.reshape(-1, points.shape[-1]).to(torch.float32)
    n, d = x.shape
    L, cap, feat = table.shape
    d1 = d + 1
    outs = []
    for l in range(L):
        cf = (x + shift[l]) * scale[l]  # (n,d) fp32
        # elevate (sequential fp32 sums in the published order: i = d .. 1)
        E = torch.empty(n, d1, dtype=torch.float32)
        sm = torch.zeros(n, dtype=torch.float32)
        for i in range(d, 0, -1):
            E[:, i] = sm - i * cf[:, i - 1]
            sm = sm + cf[:, i - 1]
        E[:, 0] = sm
        v = E * (1.0 / d1)
        up = torch.ceil(v) * d1
        down = torch.floor(v) * d1
        rem0 = torch.where(up - E < E - down, up, down).to(torch.int64)
        s = torch.div(rem0.sum(-1), d1, rounding_mode="trunc")
        resid = E - rem0.to(torch.float32)
        rank = torch.zeros(n, d1, dtype=torch.int64)
        for i in range(d):
            for j in range(i + 1, d1):
                lt = resid[:, i] < resid[:, j]
                rank[:, i] += lt
                rank[:, j] += ~lt
        rank = rank + s[:, None]
        neg, big = rank < 0, rank > d
        rank = rank + neg * d1 - big * d1
        rem0 = rem0 + neg * d1 - big * d1
        delta = (E - rem0.to(torch.float32)) * (1.0 / d1)
        bary = torch.zeros(n, d + 2, dtype=torch.float32)
        for i in range(d1):  # sequential accumulation order i = 0..d
            bary.scatter_add_(1, (d - rank[:, i])[:, None], delta[:, i:i + 1])
            bary.scatter_add_(1, (d + 1 - rank[:, i])[:, None], -delta[:, i:i + 1])
        bary[:, 0] = bary[:, 0] + (1.0 + bary[:, d + 1])
        acc = torch.zeros(n, feat, dtype=torch.float32)
        for r in range(d1):
            key = rem0[:, :d] + r - d1 * (rank[:, :d] > d - r)
            h = torch.zeros(n, dtype=torch.int64)
            for i in range(d):
                h = ((h + key[:, i]) * HASH_PRIME) & 0xFFFFFFFF
            slot = h % cap
            acc = acc + bary[:, r:r + 1] * table[l][slot]
        outs.append(acc)
    out = torch.cat(outs, dim=-1)
    if concat_points:
        out = torch.cat((out, x * concat_points_scaling), dim=-1)
    return out.reshape(*leading, -1)


def encode_scalar(p, table, shift, scale):
    """Independent pure-Python scalar transcription for ONE point (cross-check only)."""
    L, cap, feat = table.shape
    d = len(p)
    d1 = d + 1
    f32 = np.float32
    out = []
    for l in range(L):
        cf = [f32(f32(p[i]) + f32(shift[l][i])) * f32(scale[l][i]) for i in range(d)]
        E = [f32(0)] * d1
        sm = f32(0)
        for i in range(d, 0, -1):
            E[i] = f32(sm - f32(f32(i) * cf[i - 1]))
            sm = f32(sm + cf[i - 1])
        E[0] = sm
        rem0, ssum = [0] * d1, 0
        for i in range(d1):
            v = f32(E[i] * f32(1.0 / d1))
            up, down = f32(np.ceil(v) * d1), f32(np.floor(v) * d1)
            rem0[i] = int(up) if f32(up - E[i]) < f32(E[i] - down) else int(down)
            ssum += rem0[i]
        ssum = int(ssum / d1)
        rank = [0] * d1
        for i in range(d):
            di = f32(E[i] - f32(rem0[i]))
            for j in range(i + 1, d1):
                if di < f32(E[j] - f32(rem0[j])):
                    rank[i] += 1
                else:
                    rank[j] += 1
        for i in range(d1):
            rank[i] += ssum
            if rank[i] < 0:
                rank[i] += d1
                rem0[i] += d1
            elif rank[i] > d:
                rank[i] -= d1
                rem0[i] -= d1
        bary = [f32(0)] * (d + 2)
        for i in range(d1):
            delta = f32(f32(E[i] - f32(rem0[i])) * f32(1.0 / d1))
            bary[d - rank[i]] = f32(bary[d - rank[i]] + delta)
            bary[d + 1 - rank[i]] = f32(bary[d + 1 - rank[i]] - delta)
        bary[0] = f32(bary[0] + f32(f32(1.0) + bary[d + 1]))
        acc = [f32(0)] * feat
        for r in range(d1):
            h = 0
            for i in range(d):
                key = rem0[i] + r
                if rank[i] > d - r:
                    key -= d1
                h = ((h + key) * HASH_PRIME) & 0xFFFFFFFF
            slot = h % cap
            for c in range(feat):
                acc[c] = f32(acc[c] + f32(bary[r] * f32(table[l][slot][c])))
        out.extend(acc)
    return np.array(out, dtype=np.float32)
