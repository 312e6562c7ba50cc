"""Multi-GPU plumbing (one process per GPU, torch.distributed): where the hot path shards.

* Encoding: frames/videos are independent -> contiguous ranges per rank, weights
  replicated, NO data-path collective (the reference: DistributedSampler + per-rank npz +
  dist.barrier, D/infer/extract_ref_feats.py:33-36).
* Similarity: the reference bank is sharded by rows (the reference replicates it instead,
  vsc/exhaustive_search.py:232-234 ``co.shard = False``); every rank searches its shard with GLOBAL
  row ids (``DeviceIndex.set_id_offset``); the ``[nq, k]`` partial results travel as ONE 64-bit key per entry
  (score bits | ~id) in ONE all-gather and a device kernel merges them k-way (csrc/merge.cu).
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


def _pack_cpu(D: torch.Tensor, I: torch.Tensor, keep_max: bool) -> torch.Tensor:
    """The key format of csrc/merge.cu on CPU tensors (host logic of the gloo tests): (score bits << 32) | ~id."""
    u = D.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    okey = torch.where(u >= 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
    if not keep_max:
        okey = (~okey) & 0xFFFFFFFF
    keys = (okey << 32) | ((~I) & 0xFFFFFFFF)
    return torch.where(I < 0, torch.zeros_like(keys), keys)


def _merge_cpu(keys: torch.Tensor, k: int, keep_max: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    parts, nq, kin = keys.shape
    flat = keys.permute(1, 0, 2).reshape(nq, parts * kin)
    # unsigned 64-bit order on signed storage: flip the sign bit
    order = torch.argsort(flat ^ (-0x8000000000000000), dim=1, descending=True, stable=True)[:, :k]
    best = torch.gather(flat, 1, order)
    okey = (best >> 32) & 0xFFFFFFFF
    if not keep_max:
        okey = (~okey) & 0xFFFFFFFF
    u = torch.where(okey >= 0x80000000, okey ^ 0x80000000, (~okey) & 0xFFFFFFFF)
    D = (u.to(torch.int64) - ((u >= 0x80000000).to(torch.int64) << 32)).to(torch.int32).view(torch.float32)
    I = (~best) & 0xFFFFFFFF
    pad = best == 0
    fmax = torch.finfo(torch.float32).max
    D = torch.where(pad, torch.full_like(D, -fmax if keep_max else fmax), D)
    return D, torch.where(pad, torch.full_like(I, -1), I)


def merge_partial_topk(D: torch.Tensor, I: torch.Tensor, k: int, keep_max: bool = True,
                       group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank's partial (D [nq, k_r], I [nq, k_r] with GLOBAL ids < 2^32, -1 = padding) -> the global best-k per row,
    best first, ties to the lower id (faiss semantics).  Scores and ids are packed into one 64-bit key per entry, so the
    exchange is a SINGLE all-gather of [nq, k_r] words; the k-way merge runs in a device kernel (csrc/merge.cu)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if D.is_cuda:
        from . import search
        keys = search.pack_topk(D, I, keep_max)
        if world > 1:
            gathered = torch.empty((world,) + tuple(keys.shape), dtype=torch.int64, device=keys.device)
            dist.all_gather_into_tensor(gathered, keys, group=group)
        else:
            gathered = keys.unsqueeze(0)
        return search.merge_packed_topk(gathered, k, keep_max)
    keys = _pack_cpu(D, I, keep_max)            # CPU tensors: gloo, world_size-2 tests of the host logic
    if world > 1:
        parts = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(parts, keys, group=group)
        keys = torch.stack(parts)
    else:
        keys = keys.unsqueeze(0)
    return _merge_cpu(keys, k, keep_max)


def gather_partial_topk_multi(parts, keep_max: bool = True, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Several partial top-k results of the same query rows in ONE all-gather: ``parts`` = [(D, I), ...] with D / I
    [nq, k_i]; the packed keys are laid side by side ([nq, sum k_i]) and gathered once -> int64 [world, nq, sum k_i].
    Merge result i with ``search.merge_packed_topk_cols(keys, col0_i, k_i, k)`` (config 3's step: the noise search's
    top-nk and the reference search's top-k travel together)."""
    from . import search
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nq = parts[0][0].shape[0]
    ld = sum(D.shape[1] for D, _ in parts)
    keys = torch.empty((nq, ld), dtype=torch.int64, device=parts[0][0].device)
    c0 = 0
    for D, I in parts:
        search.pack_topk_into(keys, c0, D, I, keep_max)      # straight into its columns: no strided copy afterwards
        c0 += D.shape[1]
    if world == 1:
        return keys.unsqueeze(0)
    gathered = torch.empty((world, nq, ld), dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(gathered, keys, group=group)
    return gathered


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


def global_low_var_dim(z_shard: torch.Tensor, n_total: Optional[int] = None, group: Optional[dist.ProcessGroup] = None):
    """argmin of the column variance of the row-sharded noise bank (score_normalization.py:72).  CUDA shards: every shard
    computes (sum, sum of squared deviations from ITS mean, sum^2 / n) per column with fixed-order device kernels; ONE
    all-reduce of the float64 [3d] vector; the shards' moments combine exactly (M2 = sum M2_r + sum S_r^2 / n_r - S^2 / N);
    the result stays on the device (int32 [1] tensor, no host sync) -- feed it to ``search.sn_transform``.  ``n_total``: rows over all shards (default: all-reduced too).  CPU tensors (gloo
    tests): the same two passes in torch, returns an int."""
    sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    if n_total is None:
        n = torch.tensor([z_shard.shape[0]], dtype=torch.float64, device=z_shard.device)
        if sharded:
            dist.all_reduce(n, group=group)
        n_total = int(n.item())
    if z_shard.is_cuda:
        # one collective: per-shard (sum, centred sum of squares, sum^2 / n) combine exactly over the shards
        from . import search
        m3 = search.col_moments_local(z_shard)
        if sharded:
            dist.all_reduce(m3, group=group)
        return search.var_argmin_moments(m3, n_total)
    z = z_shard.double()
    sums = z.sum(0)
    if sharded:
        dist.all_reduce(sums, group=group)
    ss = ((z - sums / n_total) ** 2).sum(0)
    if sharded:
        dist.all_reduce(ss, group=group)
    return int(ss.argmin().item())


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
