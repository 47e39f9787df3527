"""Multi-GPU plumbing of the hot path: variants shard, the null model is replicated.

One process per GPU.  Every variant is independent given the null model (only diag(Cov) / k x k blocks are used),
so the path needs NO data-path collective: contiguous variant shards per rank, one broadcast of the null-model
operands at run start (NCCL over NVLink on the GPU box, gloo in the CPU tests), results stay per rank.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def shard_range(rank: int, world: int, total: int) -> tuple[int, int]:
    """Contiguous, balanced variant shard [first, last) of `rank`; the shards tile [0, total) exactly."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def pack_nullmodel(Sigma_i, Sigma_iX, cov, residuals):
    """null-model operands -> flat arrays in a fixed order (what travels in the broadcast)"""
    S = sp.csc_matrix(Sigma_i); S.sort_indices()
    n, q = np.asarray(Sigma_iX).shape
    return [S.data.astype(np.float64), S.indices.astype(np.int64), S.indptr.astype(np.int64),
            np.ascontiguousarray(np.asarray(Sigma_iX, dtype=np.float64).T).ravel(),
            np.asarray(cov, dtype=np.float64).ravel(order="F").copy(), np.asarray(residuals, dtype=np.float64).copy(),
            np.array([n, q], dtype=np.int64)]


def unpack_nullmodel(parts):
    n, q = (int(x) for x in parts[6])
    Sigma_i = sp.csc_matrix((parts[0], parts[1], parts[2]), shape=(n, n))
    Sigma_iX = np.asfortranarray(parts[3].reshape(q, n).T)
    cov = np.asfortranarray(parts[4].reshape(q, q, order="F"))
    return Sigma_i, Sigma_iX, cov, parts[5]


def broadcast_nullmodel(parts, rank: int, world: int, device=None):
    """Replicate rank 0's packed null model on every rank with torch.distributed (already initialised).
    `device`: torch device the collective runs on (cuda:<local_rank> for NCCL, None/cpu for gloo)."""
    if world == 1:
        return parts
    import torch
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    dtypes = [torch.float64, torch.int64, torch.int64, torch.float64, torch.float64, torch.float64, torch.int64]
    sizes = torch.tensor([x.size for x in parts] if rank == 0 else [0] * 7, dtype=torch.int64, device=dev)
    dist.broadcast(sizes, 0)
    out = []
    for i, sz in enumerate(sizes.tolist()):
        t = torch.from_numpy(parts[i]).to(dev) if rank == 0 else torch.empty(sz, dtype=dtypes[i], device=dev)
        dist.broadcast(t, 0)
        out.append(t.cpu().numpy())
    return out


def broadcast_dense_operands(arrays, rank: int, world: int, device=None):
    """Replicate rank 0's dense FP64 operands — the dense-GRM P and residuals, the SPA operands XW / XXWX_inv / muhat,
    a conditional-analysis X_adj — on every rank (same collective as broadcast_nullmodel: one broadcast per operand at
    run start).  `arrays`: list of ndarrays on rank 0 (ignored elsewhere).  Returns F-ordered 2-D arrays (a vector
    comes back as n x 1)."""
    if rank == 0:
        arrays = [np.asarray(a, dtype=np.float64) for a in arrays]
        arrays = [np.asfortranarray(a.reshape(-1, 1) if a.ndim < 2 else a) for a in arrays]
        if len(arrays) > 16:
            raise ValueError("at most 16 operands per call")
    if world == 1:
        return arrays
    import torch
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    flat = [0] * 33
    if rank == 0:
        flat[0] = len(arrays)
        for i, a in enumerate(arrays):
            flat[1 + 2 * i], flat[2 + 2 * i] = a.shape
    meta = torch.tensor(flat, dtype=torch.int64, device=dev)
    dist.broadcast(meta, 0)
    m = meta.tolist()
    out = []
    for i in range(m[0]):
        r, c = m[1 + 2 * i], m[2 + 2 * i]
        t = (torch.from_numpy(arrays[i].ravel(order="F").copy()).to(dev) if rank == 0
             else torch.empty(r * c, dtype=torch.float64, device=dev))
        dist.broadcast(t, 0)
        out.append(t.cpu().numpy().reshape((r, c), order="F"))
    return out
