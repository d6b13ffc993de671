"""CPU, world_size 2 over gloo: the host side of the multi-GPU path — contiguous video
shards, local top-K with global ids, exchange (all-to-all by query block + all-gather of the merged blocks, or one
all-gather of every list), merge — gives the same top-K as one unsharded corpus.
The merge kernel itself (dkd_merge_topk) is GPU-only and is checked in tests/test_gpu_kernels.py; here
`merge_fn` is a numpy stand-in with the same (score desc, id asc) order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_topk(scores, K, id_base):
    """(score desc, id asc) top-K with -inf / -1 padding, like dkd_topk."""
    M, N = scores.shape
    order = np.argsort(-scores, axis=1, kind="stable")[:, :K]
    s = np.full((M, K), -np.inf, np.float32)
    i = np.full((M, K), -1, np.int32)
    k = min(K, N)
    s[:, :k] = np.take_along_axis(scores, order, 1)[:, :k]
    i[:, :k] = order[:, :k] + id_base
    return torch.from_numpy(s), torch.from_numpy(i)


def _np_merge(gs, gi):
    """(G, M, K) -> (M, K): stand-in for dkd_merge_topk (padding entries have id < 0)."""
    G, M, K = gs.shape
    s = gs.permute(1, 0, 2).reshape(M, G * K).numpy()
    i = gi.permute(1, 0, 2).reshape(M, G * K).numpy()
    out_s = np.full((M, K), -np.inf, np.float32)
    out_i = np.full((M, K), -1, np.int32)
    for m in range(M):
        keep = i[m] >= 0
        ss, ii = s[m][keep], i[m][keep]
        order = np.lexsort((ii, -ss))[:K]
        out_s[m, : len(order)] = ss[order]
        out_i[m, : len(order)] = ii[order]
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _worker(rank, world, port, Nv, M, K, q, exchange="query_block"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as ge
        ge.load_package()
        from dkd_b200 import engine
        g = torch.Generator().manual_seed(123)
        scores = torch.randn(M, Nv, generator=g)
        scores[:, 3::7] = 0.5                       # exact ties across shards: lower global id must win
        lo, hi = engine.shard_range(Nv, rank, world)
        ls, li = _np_topk(scores[:, lo:hi].numpy(), K, lo)
        ms, mi = engine.merge_shards(ls, li, merge_fn=_np_merge, exchange=exchange)
        ref = O.topk_ids(scores.numpy(), K)
        kk = min(K, Nv)
        ok = np.array_equal(mi.numpy()[:, :kk], ref[:, :kk]) and np.array_equal(
            ms.numpy()[:, :kk], np.take_along_axis(scores.numpy(), ref[:, :kk], 1))
        if exchange == "query_block":       # the un-gathered form: this rank's query block only
            bs, bi, (q_lo, q_hi) = engine.merge_shards(ls, li, merge_fn=_np_merge, gather=False)
            ok = ok and np.array_equal(bi.numpy(), mi.numpy()[q_lo:q_hi]) and np.array_equal(
                bs.numpy(), ms.numpy()[q_lo:q_hi]) and (q_lo, q_hi) == ((M + world - 1) // world * rank,
                                                                      min((M + world - 1) // world * (rank + 1), M))
        q.put((rank, bool(ok), (lo, hi)))
    finally:
        dist.destroy_process_group()


def _run(world, Nv, M, K, exchange="query_block"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Nv, M, K, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_sharded_merge_world2_equals_global_topk():
    res = _run(2, Nv=501, M=37, K=100)
    assert all(ok for _, ok, _ in res)
    assert [r[2] for r in res] == [(0, 251), (251, 501)]


def test_sharded_merge_world2_all_gather_exchange():
    """The all-gather form (every rank gathers every list) gives the same result as the query-block exchange."""
    res = _run(2, Nv=501, M=37, K=100, exchange="all_gather")
    assert all(ok for _, ok, _ in res)


def test_sharded_merge_world3_ragged_query_blocks():
    """3 ranks, 37 queries: blocks of 13 / 13 / 11 queries (the last one padded on the wire)."""
    res = _run(3, Nv=200, M=37, K=100)
    assert all(ok for _, ok, _ in res)


def test_sharded_merge_world2_small_shards_pad():
    """Shards smaller than K: padded (-inf, -1) entries must not survive the merge."""
    res = _run(2, Nv=60, M=5, K=100)
    assert all(ok for _, ok, _ in res)


def test_shard_range_partitions_the_corpus(dkd):
    from dkd_b200 import engine
    for Nv in (1, 7, 2179, 1_000_000):
        for world in (1, 2, 4, 8):
            edges = [engine.shard_range(Nv, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == Nv
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert all(lo <= hi for lo, hi in edges)
