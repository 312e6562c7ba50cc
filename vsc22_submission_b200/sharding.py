"""Multi-GPU plumbing (one process per GPU, torch.distributed): where the hot path shards.

* Encoding: frames/videos are independent -> contiguous ranges per rank, weights
  replicated, NO data-path collective (the reference: DistributedSampler + per-rank npz +
  dist.barrier, D/infer/extract_ref_feats.py:33-36).
* Similarity: the reference bank is sharded by rows (the reference replicates it instead,
  vsc/exhaustive_search.py:232-234 ``co.shard = False``); every rank searches its shard with GLOBAL
  row ids (``DeviceIndex.set_id_offset``) and ONE all-gather of the ``[nq, k]`` partial results
  followed by a k-way merge gives the global top-k.
* Global candidate search (``DeviceIndex.global_search``): every rank finds the global_k best (query row, bank row)
  pairs against ITS bank shard (global bank ids through ``set_id_offset``); one all-gather of the padded
  ``[global_k]`` partial lists and a merge by (score, query row, bank row) gives the exact global list
  (``merge_partial_global_topk``) -- the union of per-shard top-K always contains the global top-K.
* Score-norm prep: column moments are all-reduced once per bank (``global_low_var_dim``).

Works with NCCL (CUDA tensors) and gloo (CPU tensors: the world_size-2 tests in
tests/test_sharding_cpu.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank's share of n units (rows / frames); sizes differ by <= 1."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_partial_topk(D: torch.Tensor, I: torch.Tensor, k: int, keep_max: bool = True,
                       group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's partial (D [nq, k_r], I [nq, k_r] with global ids, -1 = padding) and
    merge to the global best-k per row: best-first, ties to the lower id (faiss semantics)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        Ds = [torch.empty_like(D) for _ in range(world)]
        Is = [torch.empty_like(I) for _ in range(world)]
        dist.all_gather(Ds, D.contiguous(), group=group)
        dist.all_gather(Is, I.contiguous(), group=group)
        D, I = torch.cat(Ds, dim=1), torch.cat(Is, dim=1)
    # order by (score best-first, id ascending); padding entries (id -1) go last
    bad = I < 0
    key = torch.where(bad, torch.full_like(D, float("-inf") if keep_max else float("inf")), D)
    idkey = torch.where(bad, torch.full_like(I, torch.iinfo(torch.int64).max), I)
    order = torch.argsort(idkey, dim=1, stable=True)
    key, D, I = torch.gather(key, 1, order), torch.gather(D, 1, order), torch.gather(I, 1, order)
    order = torch.argsort(key, dim=1, descending=keep_max, stable=True)[:, :k]
    return torch.gather(D, 1, order), torch.gather(I, 1, order)


def merge_partial_global_topk(scores: torch.Tensor, qrows: torch.Tensor, brows: torch.Tensor, global_k: int,
                              keep_max: bool = True, group: Optional[dist.ProcessGroup] = None
                              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather every rank's partial global-search result (variable length <= global_k; bank rows are GLOBAL ids) and
    keep the ``global_k`` best pairs overall (all of them when ``global_k <= 0``, threshold mode): best first, ties by
    (query row, bank row) -- the order ``DeviceIndex.global_search`` produces on one GPU."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        n = torch.tensor([scores.numel()], dtype=torch.int64, device=scores.device)
        ns = [torch.empty_like(n) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        cap = max(int(max(x.item() for x in ns)), 1)
        pad = lambda t, fill: torch.cat([t, torch.full((cap - t.numel(),), fill, dtype=t.dtype, device=t.device)])
        parts = []
        for t, fill in ((scores, 0.0), (qrows, -1), (brows, -1)):
            outs = [torch.empty(cap, dtype=t.dtype, device=t.device) for _ in range(world)]
            dist.all_gather(outs, pad(t.contiguous(), fill), group=group)
            parts.append(torch.cat([o[: int(m.item())] for o, m in zip(outs, ns)]))
        scores, qrows, brows = parts
    order = torch.argsort(brows, stable=True)
    order = order[torch.argsort(qrows[order], stable=True)]
    order = order[torch.argsort(scores[order], descending=keep_max, stable=True)]
    if global_k > 0:
        order = order[:global_k]
    return scores[order], qrows[order], brows[order]


def global_low_var_dim(z_shard: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> int:
    """argmin of the column variance of the row-sharded noise bank (score_normalization.py:72)."""
    z = z_shard.double()
    mom = torch.stack([z.sum(0), (z * z).sum(0), torch.full((z.shape[1],), float(z.shape[0]), dtype=torch.float64,
                                                             device=z.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mom, group=group)
    mean = mom[0] / mom[2]
    return int((mom[1] / mom[2] - mean * mean).argmin().item())


def gather_descriptors(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Optional single all-gather of the per-rank descriptors [n_r, D] (replaces the reference's
    per-rank npz files + rank-0 merge, extract_ref_feats.py:35-57). Ranks hold shard_range() slices."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = max(shard_range(n_total, world, r)[1] - shard_range(n_total, world, r)[0] for r in range(world))
    pad = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, world, r)
        parts.append(outs[r][: b - a])
    return torch.cat(parts, dim=0)
