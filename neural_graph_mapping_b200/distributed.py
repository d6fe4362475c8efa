"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference is single-GPU (no collective anywhere).  Rays never interact and fields interact
only read-only, so the path shards with NO data-path collective; the one exchange step is a
single all-gather of the rendered tiles (9 fp32 = 36 B per ray: rgbd 4, colour var 3, depth var
1, term prob 1) so every rank ends with the full Prediction.

Three partitions of a render batch:

* ``render_rays_gathered``: every rank renders ITS OWN (F, R) batch (its keyframe) -- weak scaling;
* ``render_rays_sharded``:  the GLOBAL (F, R) batch split by FIELD, rank r gets fields [f0, f1) and therefore only
  ever touches its own fields' parameters (the training / BASELINE config 4 shape);
* ``render_rays_split``:    the GLOBAL (F, R) batch split by RAY inside every field (one keyframe over N GPUs).

The all-gather can be left in flight (``async_gather=True`` returns a :class:`PendingTiles`): NCCL runs on its own
stream, so step i's collective overlaps step i+1's render and the caller waits only when it needs the tiles.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .renderer import Prediction, _next_seed, render_rays

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


class PendingTiles:
    """An all-gather of rendered tiles that may still be in flight.  ``wait()`` orders the current CUDA stream
    after the collective (no host block) and returns the (world, 9 n) buffer; ``local`` is this rank's own tile,
    valid in stream order right after the render (e.g. for a device-to-host copy that need not wait)."""

    def __init__(self, buffer: torch.Tensor, local: torch.Tensor, work=None) -> None:
        self.buffer, self.local, self._work = buffer, local, work

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self.buffer


def gather_tiles(local_flat: torch.Tensor, group=None, out: Optional[torch.Tensor] = None, async_op: bool = False):
    """All-gather equally sized packed tiles -> (world, len(local_flat)); with ``async_op`` a PendingTiles."""
    d = _dist()
    world, _ = world_info(group)
    if d is None or world == 1:
        buf = local_flat.view(1, -1)
        return PendingTiles(buf, local_flat) if async_op else buf
    if out is None:
        out = torch.empty(world, local_flat.numel(), device=local_flat.device, dtype=local_flat.dtype)
    work = d.all_gather_into_tensor(out.view(-1), local_flat.contiguous(), group=group, async_op=async_op)
    return PendingTiles(out, local_flat, work) if async_op else out


class _EventWork:
    """``wait()`` orders the current stream after an event recorded on another stream (no host block)."""

    def __init__(self, event) -> None:
        self._event = event

    def wait(self) -> None:
        torch.cuda.current_stream().wait_event(self._event)


class TileExchange:
    """The tile exchange FUSED into the kernel that writes the Prediction (replaces the NCCL all-gather).

    A symmetric buffer of ``slots x world`` tiles (``torch.distributed._symmetric_memory``: the same allocation on
    every rank, each mapped into every process over NVLink, plus ONE NVSwitch multicast mapping where the fabric has
    it).  Rank r renders straight into tile r of a slot and the compositor kernel (or the opt-in single fused kernel) repeats its Prediction stores at
    the byte offsets ``mirrors`` -- the multicast mapping (the switch replicates the store to all ranks) or, without
    multicast, the world-1 peer mappings -- so the transfer rides along with the math, ray by ray, and the exchange
    ends with a barrier over the buffer's signal pads instead of a collective.

    Slots form a ring over the steps of a loop (``acquire`` hands out slot ``step % slots``).  A rank reads a step's
    tiles after that step's barrier and BEFORE it issues its next render (stream order); ``acquire`` orders the
    render of step i after the barrier of step ``i - slots + 1``: every rank has then finished the render that FOLLOWS the
    step being overwritten, hence its reads of it.  With 3 slots a rank may run one step
    ahead of the slowest peer (the slack the in-flight NCCL all-gather had); with 2 every step waits for all ranks.
    """

    def __init__(self, n_rays: int, device, group=None, slots: int = 3, multicast: bool = True) -> None:
        import torch.distributed._symmetric_memory as symm

        d = _dist()
        if d is None:
            raise RuntimeError("TileExchange needs an initialised process group")
        self.world, self.rank = world_info(group)
        if self.world - 1 > 8 and not multicast:
            raise ValueError("at most 8 peer mirrors")
        self.n_rays, self.slots = int(n_rays), int(slots)
        self.tile_floats = FLOATS_PER_RAY * self.n_rays
        self.stride = (self.tile_floats + 3) // 4 * 4  # tiles start 16-byte aligned (the rgbd vector store)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.buffer = symm.empty(self.slots * self.world * self.stride, dtype=torch.float32, device=self.device)
            self.handle = symm.rendezvous(self.buffer, group if group is not None else d.group.WORLD)
            self.buffer.zero_()
        ptrs = [int(x) for x in self.handle.buffer_ptrs]
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0) if multicast else 0
        self.multicast = mc != 0
        if self.multicast:
            self.mirrors = [mc - ptrs[self.rank]]
        else:
            self.mirrors = [ptrs[r] - ptrs[self.rank] for r in range(self.world) if r != self.rank]
        self._side = torch.cuda.Stream(device=self.device)
        self._step = 0
        self._done = {}  # step -> event recorded after that step's barrier
        torch.cuda.synchronize(self.device)
        self.handle.barrier(channel=0)  # every rank's zero fill has landed before anyone's first mirrored store

    def acquire(self) -> int:
        """Slot of the next step; orders the current stream after the barrier that frees it (no host block)."""
        ev = self._done.pop(self._step - self.slots + 1, None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return self._step % self.slots

    def tile(self, slot: int, rank: Optional[int] = None) -> torch.Tensor:
        r = self.rank if rank is None else rank
        o = (slot * self.world + r) * self.stride
        return self.buffer[o:o + self.tile_floats]

    def gathered(self, slot: int) -> torch.Tensor:
        """(world, 9 n) view of a slot: every rank's tile, valid after ``finish``."""
        return self.buffer[slot * self.world * self.stride:(slot + 1) * self.world * self.stride].view(
            self.world, self.stride)[:, :self.tile_floats]

    def finish(self, slot: int, async_op: bool = False):
        """Barrier over the signal pads: afterwards every rank's tile of ``slot`` is complete on every rank.  With
        ``async_op`` the barrier runs on a side stream (the next render does not wait for the slowest peer) and a
        PendingTiles is returned."""
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._side):  # all barriers of an exchange run on ONE stream, in step order
            self._side.wait_event(ev)
            self.handle.barrier(channel=0)
            done = torch.cuda.Event()
            done.record()
        self._done[self._step] = done
        self._step += 1
        if not async_op:
            torch.cuda.current_stream().wait_event(done)
            return self.gathered(slot)
        return PendingTiles(self.gathered(slot), self.tile(slot), _EventWork(done))


def render_rays_gathered(driver, ijs, c2ws, camera, field_ids, near=None, far=None, gt=None,
                         return_packed: bool = False, group=None, async_gather: bool = False,
                         buffers: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                         exchange: Optional["TileExchange"] = None, **kw):
    """Each rank renders ITS OWN (F, R) ray batch (e.g. its keyframe / its pixel tiles); the packed
    tiles of all ranks are all-gathered.  Returns the packed (world, 9*F*R) buffer, a PendingTiles
    (``async_gather``), or a list of per-rank ``Prediction`` views.  ``buffers`` = (local 9n, gathered world x 9n)
    preallocated tensors to reuse (a step loop double-buffers them).  ``exchange`` = a TileExchange: the
    tiles travel as mirrored stores of the compositor kernel itself instead of an NCCL all-gather."""
    F, R = ijs.shape[0], ijs.shape[1]
    n = F * R
    if exchange is not None:
        ex, slot = exchange, exchange.acquire()
        if n != ex.n_rays:
            raise ValueError(f"the exchange was sized for {ex.n_rays} rays per rank, got {n}")
        rgbd, cvar, dvar, term = packed_views(ex.tile(slot), n)
        render_rays(driver, ijs, c2ws, camera, field_ids, True, near, far, gt,
                    out=(rgbd.view(F, R, 4), cvar.view(F, R, 3), dvar.view(F, R), term.view(F, R)),
                    mirrors=ex.mirrors, **kw)
        res = ex.finish(slot, async_op=async_gather)
        if async_gather or return_packed:
            return res
        preds = []
        for r in range(ex.world):
            a, b, c, d = packed_views(res[r], n)
            preds.append(Prediction(a.view(F, R, 4), b.view(F, R, 3), c.view(F, R), d.view(F, R), None, None))
        return preds
    if buffers is not None:
        local, out = buffers
    else:
        local, out = torch.empty(FLOATS_PER_RAY * n, device=ijs.device, dtype=torch.float32), None
    rgbd, cvar, dvar, term = packed_views(local, n)
    render_rays(driver, ijs, c2ws, camera, field_ids, True, near, far, gt,
                out=(rgbd.view(F, R, 4), cvar.view(F, R, 3), dvar.view(F, R), term.view(F, R)), **kw)
    if async_gather:
        return gather_tiles(local, group, out, async_op=True)
    buf = gather_tiles(local, group, out)
    if return_packed:
        return buf
    preds = []
    for r in range(buf.shape[0]):
        a, b, c, d = packed_views(buf[r], n)
        preds.append(Prediction(a.view(F, R, 4), b.view(F, R, 3), c.view(F, R), d.view(F, R), None, None))
    return preds


def shared_seed(device, group=None) -> int:
    """One jitter seed for all ranks: rank 0 draws it (torch's CPU generator, like every render), an 8-byte
    broadcast hands it out."""
    d = _dist()
    world, rank = world_info(group)
    seed = _next_seed()
    if d is None or world == 1:
        return seed
    t = torch.tensor([seed], dtype=torch.int64, device=device)
    src = d.get_global_rank(group, 0) if group is not None else 0
    d.broadcast(t, src=src, group=group)
    return int(t.item())


def _leading_slice(x, lead_shape, sl):
    """Slice a per-ray input by ``sl`` along its leading dims if it carries them; scalars / shared values pass."""
    if x is None or not torch.is_tensor(x) or x.dim() < len(lead_shape) or tuple(x.shape[:len(lead_shape)]) != lead_shape:
        return x
    return x[sl]


def render_rays_sharded(driver, ijs, c2ws, camera, field_ids, near=None, far=None, gt=None, group=None,
                        jitter=None, jitter_guided=None, seed: Optional[int] = None, **kw):
    """Training-batch shape: the GLOBAL (F, R) batch is known on every rank;
    rank r renders fields [f0, f1) -- so it only ever touches its own fields' parameters -- and
    the tiles are all-gathered into the full (F, R) Prediction on every rank.

    Per-ray inputs given for the whole batch (``near``/``far``/``gt``, per-ray ``c2ws``, ``jitter`` /
    ``jitter_guided``) are sliced to the shard; a shared (4, 4) ``c2ws`` is passed through.  Without injected
    jitter all ranks use ONE seed (``seed`` or a broadcast from rank 0) and every shard's samples keep their global
    indices, so the sharded render is bit-identical to the same batch rendered on one GPU with that seed."""
    world, rank = world_info(group)
    F, R = ijs.shape[0], ijs.shape[1]
    if F % world != 0:
        raise ValueError(f"num_fields={F} must be divisible by the world size {world}")
    f0, f1 = shard_range(F, world, rank)
    sl = slice(f0, f1)
    lead = (F, R)
    c2 = c2ws if c2ws.numel() == 16 else _leading_slice(c2ws, lead, sl)
    gt_s = _leading_slice(gt, lead, sl)
    S = int(driver._num_samples)
    St = S + (int(driver._num_samples_depth_guided) if gt is not None else 0)
    if jitter is None and seed is None:
        seed = shared_seed(ijs.device, group)
    buf = render_rays_gathered(
        driver, ijs[sl], c2, camera, field_ids[sl], _leading_slice(near, lead, sl), _leading_slice(far, lead, sl), gt_s,
        return_packed=True, group=group, jitter=_leading_slice(jitter, lead, sl),
        jitter_guided=_leading_slice(jitter_guided, lead, sl), seed=seed, sample_offset=f0 * R * St, **kw)
    n = (f1 - f0) * R
    parts = [packed_views(buf[r], n) for r in range(buf.shape[0])]
    cat = lambda i, shape: torch.cat([p[i] for p in parts]).view(*shape)  # noqa: E731
    return Prediction(cat(0, (F, R, 4)), cat(1, (F, R, 3)), cat(2, (F, R)), cat(3, (F, R)), None, None)


def render_rays_split(driver, ijs, c2ws, camera, field_ids, near=None, far=None, gt=None, group=None,
                      jitter=None, jitter_guided=None, return_packed: bool = False, async_gather: bool = False,
                      buffers=None, **kw):
    """ONE batch over N GPUs by RAY (strong scaling of a keyframe render): rank r renders rays
    [r0, r1) of EVERY field, so all ranks need all fields' parameters (replicated, ~9 MB fp16 for 75 fields) and the
    per-rank kernel time drops by N.  R must be divisible by the world size.  The result is the full (F, R)
    Prediction on every rank (or the packed (world, 9 F R/N) buffer / a PendingTiles).  In-kernel jitter of a split
    render is a valid stratified draw but not the single-GPU stream (ray indices restart per shard); inject
    ``jitter`` for bit-identity."""
    world, rank = world_info(group)
    F, R = ijs.shape[0], ijs.shape[1]
    if R % world != 0:
        raise ValueError(f"rays_per_field={R} must be divisible by the world size {world}")
    r0, r1 = shard_range(R, world, rank)
    sl = (slice(None), slice(r0, r1))
    lead = (F, R)
    c2 = c2ws if c2ws.numel() == 16 else _leading_slice(c2ws, lead, sl)
    res = render_rays_gathered(
        driver, ijs[sl], c2, camera, field_ids, _leading_slice(near, lead, sl), _leading_slice(far, lead, sl),
        _leading_slice(gt, lead, sl), return_packed=True, group=group, async_gather=async_gather, buffers=buffers,
        jitter=_leading_slice(jitter, lead, sl), jitter_guided=_leading_slice(jitter_guided, lead, sl), **kw)
    if async_gather or return_packed:
        return res
    Rs = r1 - r0
    parts = [packed_views(res[r], F * Rs) for r in range(res.shape[0])]
    cat = lambda i, tail: torch.cat([p[i].view(F, Rs, *tail) for p in parts], 1)  # noqa: E731
    return Prediction(cat(0, (4,)), cat(1, (3,)), cat(2, ()), cat(3, ()), None, None)


def render_image_sharded(driver, c2w: torch.Tensor, camera, group=None, seed: Optional[int] = None):
    """``render_image`` (ngm/run_mapping.py:402-437: full frame through the kNN path over ALL fields) with the pixel
    rows split over the ranks and one all-gather of the (rgbd, depth variance) tiles -- the re-render of BASELINE
    config 5 (after a pose-graph update only two small tensors, the field positions / orientations, have changed; every
    rank holds all field parameters).  All ranks use one jitter seed and global sample indices, so the frame equals
    the single-GPU frame rendered with that seed.  Returns (rgbds (H, W, 4), depth_vars (H, W)) on every rank."""
    world, rank = world_info(group)
    h, w = camera.height, camera.width
    if h % world != 0:
        raise ValueError(f"image height {h} must be divisible by the world size {world}")
    dev = driver._device
    r0, r1 = shard_range(h, world, rank)
    if seed is None:
        seed = shared_seed(torch.device(dev), group)
    ijs = torch.cartesian_prod(torch.arange(r0, r1, device=dev), torch.arange(w, device=dev))
    with torch.no_grad():
        p = render_rays(driver, ijs, c2w, camera, seed=seed, sample_offset=r0 * w * int(driver._num_samples))
    n = (r1 - r0) * w
    local = torch.empty(5 * n, device=ijs.device, dtype=torch.float32)
    local[:4 * n].view(n, 4).copy_(p.rgbds)
    local[4 * n:].copy_(p.depth_vars)
    buf = gather_tiles(local, group)
    rgbd = torch.cat([buf[r, :4 * n].view(r1 - r0, w, 4) for r in range(buf.shape[0])])
    dvar = torch.cat([buf[r, 4 * n:].view(r1 - r0, w) for r in range(buf.shape[0])])
    return rgbd, dvar
