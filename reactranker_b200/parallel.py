"""Data-parallel training of the hot path over the GPUs of one box (new functionality: the reference is single-device,
SURVEY.md §2a / §8e).

The path shards naturally: reactant groups are independent (every loss is a sum over groups, the encoder is per
molecule), so whole groups are assigned to ranks as contiguous runs balanced by atom count and the only exchange step is
one SUM all-reduce of the flat gradient (3.16 MB at h=300, 12.1 MB at h=600) over NCCL / NVLink.  Exactness w.r.t. the
single-device step is kept by two rules:

* every rank packs its shard with the GLOBAL batch's ``max_num_bonds`` (padding-row multiplicities, SURVEY.md §0 trap 1);
* every rank divides its loss by the GLOBAL normaliser (groups / items / ordered pairs) and gradients are summed.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_groups(group_atoms: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split groups 0..G-1 into ``world`` contiguous runs ``[lo, hi)`` with balanced total atom count; never splits a
    group.  Every rank gets at least one group when G >= world; trailing ranks may be empty otherwise."""
    w = np.asarray(group_atoms, dtype=np.float64)
    G = len(w)
    if world <= 1:
        return [(0, G)]
    cum = np.concatenate(([0.0], np.cumsum(w)))
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(np.searchsorted(cum, target, side="left"))
        if cut > 0 and abs(cum[cut - 1] - target) <= abs(cum[min(cut, G)] - target):
            cut -= 1
        cut = max(cut, bounds[-1] + (1 if G - bounds[-1] > world - r else 0))   # leave groups for the remaining ranks
        cut = min(cut, G - (world - r)) if G >= world else min(cut, G)
        cut = max(cut, bounds[-1])
        bounds.append(cut)
    bounds.append(G)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_rows(scope: Sequence[int], lo: int, hi: int) -> Tuple[int, int]:
    """Row range of groups ``[lo, hi)`` inside a batch whose group sizes are ``scope``."""
    off = np.concatenate(([0], np.cumsum(np.asarray(scope, dtype=np.int64))))
    return int(off[lo]), int(off[hi])


class GradSync:
    """SUM all-reduce of all gradients through one flat bucket (one collective per step)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        if self.flat is None or self.flat.device != grads[0].device:
            self.flat = torch.empty(self.numel, dtype=grads[0].dtype, device=grads[0].device)
        views = list(torch.split(self.flat, [g.numel() for g in grads]))
        torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_([g.reshape(-1) for g in grads], views)
        for p, g in zip(self.params, grads):
            if p.grad is None:
                p.grad = g


def broadcast_parameters(model, src: int = 0, group=None):
    """Make every rank start from rank ``src``'s weights (and frozen buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
