"""Multi-GPU partitioning of the denoising loop.  One process per GPU (torch.distributed).

Two independent axes:
  * images (trajectories) are independent -> shard the batch across ranks, NO collective on the data path;
    the final latents are gathered once at the end (`gather_images`).
  * when the batch is smaller than the number of GPUs, the (1 + N) conditioning rows of one image are
    independent UNet evaluations that meet only in the CFG combine -> shard rows across the ranks of an image
    group and all-gather the eps rows (32-128 KB per rank per step) into every rank of the group; each rank
    then runs the tiny fused step redundantly so x stays replicated (no broadcast).
The partitioning arithmetic is plain Python (tested on CPU with the gloo backend, world size 2); the
all-gather runs on NCCL over NVLink when the tensors are CUDA tensors.
"""
from dataclasses import dataclass
from typing import List

import torch


@dataclass
class Partition:
    world: int
    rank: int
    images: List[int]        # global image indices this rank works on
    rows: List[int]          # conditioning rows (0 = unconditional) this rank evaluates for each of its images
    group_ranks: List[int]   # ranks sharing the same images (row-sharding group); [rank] when images are sharded
    rows_total: int

    @property
    def needs_allgather(self):
        return len(self.group_ranks) > 1


def partition(batch: int, rows_total: int, world: int, rank: int) -> Partition:
    """batch >= world: contiguous image shards (sizes differ by at most 1), every rank evaluates all rows.
    batch < world: ranks are split into `batch` groups (sizes differ by at most 1); the ranks of a group split the
    rows of that image between them (ranks beyond rows_total in a group stay idle but still step x)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    if batch <= 0 or rows_total <= 0:
        return Partition(world, rank, [], [], [rank], rows_total)
    if batch >= world:
        base, extra = divmod(batch, world)
        start = rank * base + min(rank, extra)
        n = base + (1 if rank < extra else 0)
        return Partition(world, rank, list(range(start, start + n)), list(range(rows_total)), [rank], rows_total)
    base, extra = divmod(world, batch)
    # group g owns ranks [g_start, g_start + size)
    g, g_start = 0, 0
    while True:
        size = base + (1 if g < extra else 0)
        if rank < g_start + size:
            break
        g_start += size
        g += 1
    members = list(range(g_start, g_start + size))
    local = rank - g_start
    active = min(size, rows_total)
    if local < active:
        rb, re_ = divmod(rows_total, active)
        r0 = local * rb + min(local, re_)
        rows = list(range(r0, r0 + rb + (1 if local < re_ else 0)))
    else:
        rows = []
    return Partition(world, rank, [g], rows, members, rows_total)


class RowGatherPlan:
    """Preallocated buffers of the per-step eps all-gather (one `all_gather_into_tensor` on the compute stream, no per-step
    allocation, list or cat): `send` [max_rows, L] is what this rank contributes, `recv` [size * max_rows, L] what it
    receives.  When every active rank owns the same number of rows and no rank idles, `recv` already IS the row-ordered
    [rows_total, L] result; otherwise one index_select with a fixed index picks the valid rows."""

    def __init__(self, part: Partition, L: int, dtype, device):
        size = len(part.group_ranks)
        self.active = min(size, part.rows_total)
        self.max_rows = -(-part.rows_total // self.active)
        self.send = torch.zeros(self.max_rows, L, dtype=dtype, device=device)
        self.recv = torch.empty(size * self.max_rows, L, dtype=dtype, device=device)
        rb, re_ = divmod(part.rows_total, self.active)
        idx = []
        for local in range(self.active):
            n = rb + (1 if local < re_ else 0)
            idx += [local * self.max_rows + k for k in range(n)]
        self.direct = idx == list(range(part.rows_total)) and size * self.max_rows == part.rows_total
        self.index = None if self.direct else torch.tensor(idx, dtype=torch.long, device=device)
        self.key = (tuple(part.group_ranks), part.rows_total, L, dtype, torch.device(device))


_ROW_PLANS = {}
_GROUPS = {}  # world size -> {member ranks: process group}


def allgather_eps_rows(local_rows: torch.Tensor, part: Partition, group=None) -> torch.Tensor:
    """local_rows: [len(part.rows), L] eps rows this rank computed for its image.  Returns [rows_total, L] on every
    rank of the group (row order 0..rows_total-1): a preallocated `all_gather_into_tensor` (row counts differ by <= 1, so
    every rank sends max_rows rows; ranks with fewer leave the tail of their block unused)."""
    import torch.distributed as dist
    if not part.needs_allgather:
        return local_rows
    if local_rows.ndim != 2:
        raise ValueError("allgather_eps_rows expects [rows, L]; idle ranks pass an empty [0, L] tensor")
    L = local_rows.shape[1]
    key = (tuple(part.group_ranks), part.rows_total, L, local_rows.dtype, local_rows.device)
    plan = _ROW_PLANS.get(key)
    if plan is None:
        plan = _ROW_PLANS[key] = RowGatherPlan(part, L, local_rows.dtype, local_rows.device)
    if local_rows.shape[0]:
        plan.send[: local_rows.shape[0]].copy_(local_rows)
    dist.all_gather_into_tensor(plan.recv, plan.send, group=group)
    return plan.recv if plan.direct else plan.recv.index_select(0, plan.index)


def shared_noise_sampler(base, group, src):
    """Row-sharded groups step a REPLICATED x: every stochastic draw (ancestral noise, churn, the img2img start noise) must
    be the same tensor on every rank of the group.  The group leader (global rank `src`) draws - through the caller's
    `noise_sampler` if there is one - and broadcasts; the other ranks' own draws are discarded (their RNG streams still
    advance in step)."""
    import torch.distributed as dist

    def draw(x):
        n = base(x) if base is not None else torch.randn_like(x)
        n = n.to(x.device, torch.float32).contiguous()
        dist.broadcast(n, src=src, group=group)
        return n
    return draw


def gather_images(local_x: torch.Tensor, batch: int, world: int, group=None) -> torch.Tensor:
    """All ranks receive the full [batch, ...] tensor of final latents (image-sharded case)."""
    import torch.distributed as dist
    if world == 1:
        return local_x
    counts = [len(partition(batch, 1, world, r).images) for r in range(world)]
    mx = max(counts)
    buf = local_x.new_zeros((mx,) + tuple(local_x.shape[1:]))
    buf[: local_x.shape[0]] = local_x
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def sample_sharded(wrapper, *, steps, batch, shape, x_T, conditioning, unconditional_conditioning, world=None, rank=None, **kw):
    """Run `wrapper.sampler.sample` for a global batch of `batch` images on all ranks of the default process group.
    batch >= world: images are sharded (no data-path collective); batch < world: the ranks of an image's group split
    its (1 + N) conditioning rows and all-gather eps every step.  x_T is the GLOBAL [batch, C, h, w] start noise (same
    on every rank).  Returns the full [batch, C, h, w] result on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    rows_total = 1 + len(conditioning["and"]) + len(conditioning.get("not", []))
    part = partition(batch, rows_total, world, rank)
    sampler = wrapper.sampler
    groups = None
    if batch < world:
        # one NCCL sub-group per image (every rank must create every group, in the same order)
        # (communicators are cached: creating one costs far more than a whole generation)
        groups, seen = _GROUPS.setdefault(world, {}), set()
        for r in range(world):
            members = tuple(partition(batch, rows_total, world, r).group_ranks)
            if members not in seen:
                seen.add(members)
                if members not in groups:
                    groups[members] = dist.new_group(list(members)) if len(members) > 1 else None
        grp = groups[tuple(part.group_ranks)]
        sampler.denoiser.set_row_partition(part, grp)
        if len(part.group_ranks) > 1:  # one draw per group, not per rank (ancestral / churn / img2img noise)
            shared = shared_noise_sampler(kw.get("noise_sampler"), grp, part.group_ranks[0])
            kw["noise_sampler"] = shared
            sampler.noise_sync = shared
    else:
        sampler.denoiser.set_row_partition(None)
        sampler.noise_sync = None
    mine = x_T[part.images]
    try:
        out = sampler.sample(steps=steps, batch_size=len(part.images), shape=shape, x_T=mine, conditioning=conditioning,
                             unconditional_conditioning=unconditional_conditioning, **kw)
    finally:
        sampler.noise_sync = None
    if batch >= world:
        return gather_images(out, batch, world)
    # row-sharded: every rank of a group holds the same image; gather one copy per image
    outs = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(outs, out.contiguous())
    firsts = {}
    for r in range(world):
        img = partition(batch, rows_total, world, r).images[0]
        firsts.setdefault(img, outs[r])
    return torch.cat([firsts[i] for i in range(batch)])
