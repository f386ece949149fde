"""Sharding of the hot path across the GPUs of one box: one process per GPU, independent units, no
exchange in the data path.

The reference parallelises the same way — one future per contract symbol over a process pool
(reference src/interpolation/batch_processor.py:234-239) or per size-balanced batch of symbols
(reference src/interpolation/optimized_batch_processor.py:123-164: sort by rows x interpolation factor,
cut greedily).  Here the unit is a contract (kernels A/B) or a whole chain = underlying x trading day
(surface pipeline); units are dealt to ranks longest-processing-time-first, each rank packs its units
into one CSR batch and runs the kernels on its own GPU.  The only collective is the optional final
gather of the (small) candle tables, over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import heapq
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def lpt_partition(costs: Sequence[float], n_shards: int) -> List[List[int]]:
    """Longest-processing-time-first: heaviest unit to the currently lightest shard.  Deterministic
    (ties by index), every unit in exactly one shard, units inside a shard in ascending index order."""
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    heap = [(0.0, s) for s in range(n_shards)]
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for i in order:
        load, s = heapq.heappop(heap)
        shards[s].append(int(i))
        heapq.heappush(heap, (load + float(costs[i]), s))
    for sh in shards:
        sh.sort()
    return shards


def imbalance(costs: Sequence[float], shards: List[List[int]]) -> float:
    """max / mean shard load (1.0 = perfectly balanced)."""
    costs = np.asarray(costs, dtype=np.float64)
    loads = np.array([costs[s].sum() if len(s) else 0.0 for s in shards])
    return float(loads.max() / loads.mean()) if loads.mean() > 0 else 1.0


def contract_costs(src_off: np.ndarray, src_ts: np.ndarray, step: int) -> np.ndarray:
    """Output rows each contract will produce (span / step + 1): the cost model of kernels A and B, the
    analogue of the reference's ``row_count * interpolation_factor`` (optimized_batch_processor.py:129-133)."""
    src_off = np.asarray(src_off, dtype=np.int64)
    src_ts = np.asarray(src_ts, dtype=np.int64)
    n = len(src_off) - 1
    out = np.zeros(n, dtype=np.float64)
    a, b = src_off[:-1], src_off[1:]
    ok = b > a
    if ok.any():
        out[ok] = (src_ts[b[ok] - 1] - src_ts[a[ok]]) // step + 1
    return out


def my_units(costs: Sequence[float], rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return lpt_partition(costs, world)[rank]


def gather_tables(local: Dict[str, torch.Tensor], dst: int = 0, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Final gather: every rank contributes a table (dict of equally long 1-D tensors, same keys and dtypes on
    all ranks); ``dst`` receives their concatenation in rank order (+ a ``_rank`` column), the others ``None``.
    Row counts are exchanged first (one small all_gather); the columns then travel point to point, each rank's rows
    landing directly in their slice of ``dst``'s output — no padding, nothing sent to ranks that do not need it
    (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    keys = sorted(local)
    n_local = int(local[keys[0]].numel()) if keys else 0
    for k in keys:
        if local[k].dim() != 1 or local[k].numel() != n_local:
            raise ValueError(f"gather_tables: column {k!r} is not a 1-D tensor of {n_local} rows")
    dev = local[keys[0]].device if keys else torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
    counts = [int(c) for c in torch.cat(counts).tolist()]
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops, out = [], {}
    if rank == dst:
        for k in keys:
            out[k] = torch.empty(int(starts[-1]), dtype=local[k].dtype, device=dev)
            out[k][int(starts[rank]):int(starts[rank + 1])] = local[k]
        for r in range(world):
            if r == dst or counts[r] == 0:
                continue
            for k in keys:
                ops.append(dist.P2POp(dist.irecv, out[k][int(starts[r]):int(starts[r + 1])], peer(r), group))
    elif n_local:
        for k in keys:
            ops.append(dist.P2POp(dist.isend, local[k].contiguous(), peer(dst), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    out["_rank"] = torch.repeat_interleave(torch.arange(world, dtype=torch.int32, device=dev),
                                           torch.tensor(counts, dtype=torch.int64, device=dev), output_size=int(sum(counts)))
    return out
