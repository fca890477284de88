"""The chunk-sharded search plumbing (rag_foundation_b200/sharded.py) under world_size = 2 on the
gloo backend: contiguous shards, one all-gather of packed keys, merge == single-index answer.
The CUDA entry points are replaced by the oracle here (CPU container); the GPU suite runs the
same class with the real kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rag_foundation_b200.sharded import ShardedSearcher, shard_range, unpack_keys_torch


def test_shard_range_partitions_exactly():
    for n, w in [(0, 2), (1, 2), (7, 3), (100_000_000, 8), (10, 16)]:
        prev = 0
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            assert lo == prev and hi >= lo
            prev = hi
        assert prev == n
        sizes = [shard_range(n, r, w)[1] - shard_range(n, r, w)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_unpack_keys_torch():
    keys = torch.tensor([[(26 << 32) | (0xFFFFFFFF - 755), 0]], dtype=torch.int64)
    ids, sc, valid = unpack_keys_torch(keys)
    assert ids.tolist() == [[755, -1]] and sc.tolist() == [[26, 0]] and valid.tolist() == [[True, False]]


def _worker(rank, world, port, n_total, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle as co, rf1
        zb = rf1.zipf_bucket_table()
        lo, hi = shard_range(n_total, rank, world)
        F = co.synth_rows(seed, lo, hi - lo, zb)
        seg = np.zeros(hi - lo, np.uint32)

        def local_search(q, scope, k):
            keys = np.stack([co.score_topk_keys(F, seg, q[i].numpy(), scope, k=k, id_base=lo) for i in range(q.shape[0])])
            return torch.from_numpy(keys.view(np.int64))

        def merge(gathered, k):
            g = gathered.numpy().view(np.uint64)
            out = np.stack([co.merge_topk(g[:, i, :], k) for i in range(g.shape[1])])
            return torch.from_numpy(out.view(np.int64))

        def local_df(scope):
            df, n = co.bucket_df(F, seg, scope)
            return torch.from_numpy(np.concatenate([df, [n]]).astype(np.int64))

        s = ShardedSearcher(local_search, merge, local_df=local_df, weights_fn=co.idf_weights)
        assert s.world == world and s.rank == rank
        np.save(os.path.join(out_dir, f"w_{rank}.npy"), s.scope_weights([0]))
        q = torch.from_numpy(np.stack([co.synth_query(seed, i, zb) for i in range(5)]))
        keys = s.search_keys(q, [0], 10)
        ids, sc, valid = unpack_keys_torch(keys)
        np.save(os.path.join(out_dir, f"ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"sc_{rank}.npy"), sc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_search_equals_single_index(tmp_path):
    from oracle import c_oracle as co, rf1
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    n_total, seed, world = 30_001, 2, 2
    mp.spawn(_worker, args=(world, port, n_total, seed, str(tmp_path)), nprocs=world, join=True)
    zb = rf1.zipf_bucket_table()
    F = co.synth_rows(seed, 0, n_total, zb)
    seg = np.zeros(n_total, np.uint32)
    w_all = co.idf_weights(*co.bucket_df(F, seg, [0]))   # RF-1w: all-reduced statistic == whole-corpus statistic
    for r in range(world):
        assert (np.load(tmp_path / f"w_{r}.npy") == w_all).all()
        ids = np.load(tmp_path / f"ids_{r}.npy")
        sc = np.load(tmp_path / f"sc_{r}.npy")
        for i in range(5):
            w_ids, w_sc, _ = co.score_topk(F, seg, co.synth_query(seed, i, zb), [0])
            assert ids[i].tolist() == w_ids.tolist() and sc[i].tolist() == w_sc.tolist()


# ---- whole stores per rank (configs[4]) -----------------------------------------------------------
def _store_worker(rank, world, port, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle as co, rf1
        from rag_foundation_b200.sharded import StoreShardedSearcher
        zb = rf1.zipf_bucket_table()
        id_base = StoreShardedSearcher.id_base_for(rank, world)
        rows, segs, names = [], [], []

        def open_local(name):
            names.append(name)
            return len(names) - 1

        def local_search(q, scopes, k):
            F = np.concatenate(rows) if rows else np.zeros((0, 256), np.int8)
            sg = np.concatenate(segs) if segs else np.zeros(0, np.uint32)
            keys = np.stack([co.score_topk_keys(F, sg, q[i].numpy(), scopes[i], k=k, id_base=id_base) if len(F) and scopes[i]
                             else np.zeros(k, np.uint64) for i in range(q.shape[0])])
            return torch.from_numpy(keys.view(np.int64))

        def merge(gathered, k):
            g = gathered.numpy().view(np.uint64)
            return torch.from_numpy(np.stack([co.merge_topk(g[:, i, :], k) for i in range(g.shape[1])]).view(np.int64))

        s = StoreShardedSearcher(local_search, merge, open_local)
        for g in range(5):                      # 5 stores x 700 rows: rank 0 owns 0, 2, 4; rank 1 owns 1, 3
            assert s.open_store(f"fileSearchStores/s{g}") == g
            if s.owner(g) == rank:
                rows.append(co.synth_rows(seed, g * 700, 700, zb))
                segs.append(np.full(700, s.local_seg[g], np.uint32))
        assert s.open_store("fileSearchStores/s3") == 3 and sorted(s.local_seg) == [g for g in range(5) if g % world == rank]
        scopes = [[0], [1], [3], [4], [0, 1], [1, 2, 3, 4], [2, 2], [], [1, 3]]
        q = torch.from_numpy(np.stack([co.synth_query(seed, i, zb) for i in range(len(scopes))]))
        ids, sc, valid = s.search(q, scopes, 10)
        np.save(os.path.join(out_dir, f"st_ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"st_sc_{rank}.npy"), sc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_store_sharded_search_equals_single_index(tmp_path):
    from oracle import c_oracle as co, rf1
    from rag_foundation_b200.sharded import StoreShardedSearcher
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    seed, world = 6, 2
    mp.spawn(_store_worker, args=(world, port, seed, str(tmp_path)), nprocs=world, join=True)
    zb = rf1.zipf_bucket_table()
    scopes = [[0], [1], [3], [4], [0, 1], [1, 2, 3, 4], [2, 2], [], [1, 3]]
    # independent restatement: every row with its global id, brute-force rank by (score desc, id asc)
    F = co.synth_rows(seed, 0, 5 * 700, zb)
    store = np.repeat(np.arange(5), 700)
    gid = np.zeros(5 * 700, np.int64)
    for g in range(5):
        r = g % world
        gid[g * 700:(g + 1) * 700] = StoreShardedSearcher.id_base_for(r, world) + (g // world) * 700 + np.arange(700)
    for r in range(world):
        ids = np.load(tmp_path / f"st_ids_{r}.npy")
        sc = np.load(tmp_path / f"st_sc_{r}.npy")
        for i, scope in enumerate(scopes):
            s = F.astype(np.int64) @ co.synth_query(seed, i, zb).astype(np.int64)
            rows = np.nonzero(np.isin(store, scope))[0]
            order = rows[np.lexsort((gid[rows], -s[rows]))][:10]
            m = len(order)
            assert ids[i][:m].tolist() == gid[order].tolist() and sc[i][:m].tolist() == s[order].tolist(), (r, scope)
            assert (ids[i][m:] == -1).all()
