"""N > 1 host logic on CPU (gloo, world_size 2): null-model replication and variant sharding."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from staarpipeline_b200 import shard, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = None
    if rank == 0:
        nm = synth.nullmodel_sparse_grm(300, seed=5, q=4, shuffle=True)
        parts = shard.pack_nullmodel(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    parts = shard.broadcast_nullmodel(parts, rank, world)
    Si, SX, cov, res = shard.unpack_nullmodel(parts)
    first, last = shard.shard_range(rank, world, 1001)
    # every rank derives its genotype shard from the global variant index: shards are disjoint and reproducible
    codes = synth.dosage_codes(300, first, min(8, last - first))
    q.put((rank, first, last, float(Si.sum()), float(SX.sum()), float(cov.sum()), float(res.sum()), int(codes.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    from staarpipeline_b200 import shard, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nm = synth.nullmodel_sparse_grm(300, seed=5, q=4, shuffle=True)
    for rank, first, last, s1, s2, s3, s4, _ in got:
        assert (first, last) == shard.shard_range(rank, 2, 1001)
        assert s1 == float(nm.Sigma_i.sum()) and s2 == float(nm.Sigma_iX.sum())
        assert s3 == float(nm.cov.sum()) and s4 == float(nm.residuals.sum())
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == 1001
    # the shard of rank 1 equals the same variants generated in a single-process run
    assert got[1][7] == int(synth.dosage_codes(300, got[1][1], 8).sum())


@pytest.mark.parametrize("world,total", [(1, 10), (2, 11), (4, 10), (8, 10_000_000), (8, 3)])
def test_shard_ranges_tile_exactly(world, total):
    from staarpipeline_b200 import shard
    edges = [shard.shard_range(r, world, total) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    from staarpipeline_b200 import shard, synth
    nm = synth.nullmodel_sparse_grm(120, seed=2, q=3)
    Si, SX, cov, res = shard.unpack_nullmodel(shard.pack_nullmodel(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals))
    assert (Si != nm.Sigma_i).nnz == 0 and np.array_equal(SX, nm.Sigma_iX)
    assert np.array_equal(cov, nm.cov) and np.array_equal(res, nm.residuals)


def _worker_dense(rank, world, port, q):
    import torch.distributed as dist
    from staarpipeline_b200 import shard, synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops = None
    if rank == 0:
        nb = synth.nullmodel_binary(257, seed=4, q=3, intercept=-2.0)
        ops = [nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat]
    XW, XXi, res, mu = shard.broadcast_dense_operands(ops, rank, world)
    q.put((rank, XW.shape, XXi.shape, res.shape, XW.flags.f_contiguous, float(XW[1, 5]), float(XXi[7, 2]), float(res.sum()), float(mu[11, 0])))
    dist.barrier()
    dist.destroy_process_group()


def test_dense_operand_broadcast_world2():
    """SPA / dense-GRM / conditional operands replicate like the sparse null model: same values, shapes and layout"""
    from staarpipeline_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dense, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nb = synth.nullmodel_binary(257, seed=4, q=3, intercept=-2.0)
    for rank, s1, s2, s3, fcont, a, b, c, d in got:
        assert s1 == (3, 257) and s2 == (257, 3) and s3 == (257, 1) and fcont
        assert a == float(nb.XW[1, 5]) and b == float(nb.XXWX_inv[7, 2]) and c == float(nb.residuals.sum()) and d == float(nb.muhat[11])
