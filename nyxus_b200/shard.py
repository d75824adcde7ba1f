"""Multi-GPU sharding of the path: ROIs/slides are independent units, so slides are partitioned across the ranks of one
node balanced by pixel count (LPT greedy), every rank runs the whole per-slide pipeline on its own B200, and the feature
rows are gathered to rank 0 over host memory -- no data-path collective, no inter-GPU reduction (BASELINE.json north_star;
the reference has no multi-device support at all: SURVEY.md section 2.1).

One process per GPU (torchrun); torch.distributed is plumbing only (rendezvous + an object gather on a gloo group)."""
from typing import Callable, List, Sequence, Tuple

import numpy as np


def lpt_partition(weights: Sequence[float], nparts: int) -> List[List[int]]:
    """Longest-processing-time-first greedy: heaviest unit to the currently lightest rank.  Deterministic."""
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    loads = [0.0] * nparts
    parts: List[List[int]] = [[] for _ in range(nparts)]
    for i in order:
        r = min(range(nparts), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += float(weights[i])
    for p in parts:
        p.sort()
    return parts


def gather_rows(local: List[Tuple[int, np.ndarray]], rank: int, world: int, group=None):
    """local: [(slide index, table[nrows, ncols])] of this rank.  Returns on rank 0 the list for ALL slides ordered by
    slide index (None elsewhere).  Host-side gather."""
    if world == 1:
        return sorted(local, key=lambda t: t[0])
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    if rank != 0:
        return None
    merged = [t for part in out for t in part]
    return sorted(merged, key=lambda t: t[0])


def featurize_sharded(slides: Sequence[Tuple[np.ndarray, np.ndarray]], featurize_fn: Callable, rank: int, world: int,
                      weights: Sequence[float] = None, group=None):
    """slides: [(intensity, label)] known to every rank (or loadable by index); featurize_fn(inten, label) -> table.
    Returns on rank 0: (table of all rows, slide index per row); rows ordered by slide then ascending label."""
    if weights is None:
        weights = [int(np.prod(s[1].shape)) for s in slides]
    mine = lpt_partition(weights, world)[rank]
    local = [(i, np.asarray(featurize_fn(slides[i][0], slides[i][1]))) for i in mine]
    allrows = gather_rows(local, rank, world, group)
    if allrows is None:
        return None
    if not allrows:
        return np.zeros((0, 0)), np.zeros((0,), np.int32)
    tab = np.concatenate([t for _, t in allrows], axis=0)
    idx = np.concatenate([np.full(len(t), i, np.int32) for i, t in allrows])
    return tab, idx
