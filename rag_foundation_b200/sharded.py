"""Chunk-sharded search across the GPUs of one box (BASELINE.json configs[3]).

Each rank (one process per GPU, torch.distributed) owns a contiguous range of global chunk ids in
its own Engine (id_base = first id of the shard).  A search is: local fused score+top-k kernel ->
ONE all-gather of the nq x k packed 64-bit keys (80 B per query per rank at k = 10) -> k-way merge
kernel on every rank.  Because the packed key is a total order on (score desc, id asc), the merged
result is bit-identical to a single-GPU scan of the whole corpus.

The local search and the merge are injected callables so the same plumbing runs under gloo on CPU
tensors in the test-suite (there the callables are test doubles); `for_engine` wires the real
CUDA entry points.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of global chunk ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def unpack_keys_torch(keys: torch.Tensor):
    """int64 tensor of packed RF-1 keys -> (ids int64, scores int32, valid bool)."""
    valid = keys != 0
    low = keys & 0xFFFFFFFF
    ids = torch.where(valid, 0xFFFFFFFF - low, torch.full_like(keys, -1))
    scores = (keys >> 32).to(torch.int32)   # scores < 2^31, so the arithmetic shift is exact
    return ids, scores, valid


def allreduce_scope_weights(local_df: Callable[[Sequence[int]], torch.Tensor], weights_fn, scope: Sequence[int],
                            group: Optional[dist.ProcessGroup] = None):
    """RF-1w weights of a sharded corpus (oracle/SPEC.md): every rank counts its own shard
    (`local_df(scope)` -> int64 [dim + 1]: df per bucket, then the live row count), ONE integer
    all-reduce (sum) makes the statistic global, and `weights_fn(df, n)` turns it into the uint8 [dim]
    weights -- identical on every rank, so the merged ranking equals the single-engine one."""
    stat = local_df(scope)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stat, op=dist.ReduceOp.SUM, group=group)
    host = stat.cpu().numpy()
    return weights_fn(host[:-1].astype("uint64"), int(host[-1]))


def engine_local_df(engine) -> Callable[[Sequence[int]], torch.Tensor]:
    def local_df(scope: Sequence[int]) -> torch.Tensor:
        dev = torch.device("cuda", torch.cuda.current_device())
        stat = torch.zeros(int(getattr(engine, "dim", 256)) + 1, dtype=torch.int64, device=dev)
        engine.scope_df_device(scope, stat.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        return stat
    return local_df


class ShardedSearcher:
    def __init__(self, local_search: Callable[[torch.Tensor, Sequence[int], int], torch.Tensor],
                 merge: Callable[[torch.Tensor, int], torch.Tensor], group: Optional[dist.ProcessGroup] = None,
                 local_df: Optional[Callable[[Sequence[int]], torch.Tensor]] = None, weights_fn=None):
        self.local_search = local_search
        self.merge = merge
        self.local_df = local_df
        self.weights_fn = weights_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather_buf: Optional[torch.Tensor] = None

    @classmethod
    def for_engine(cls, engine, group: Optional[dist.ProcessGroup] = None) -> "ShardedSearcher":
        """Real path: CUDA kernels through the C-ABI on torch's current stream."""

        def local_search(q: torch.Tensor, scope: Sequence[int], k: int) -> torch.Tensor:
            assert q.is_cuda and q.dtype == torch.int8 and q.is_contiguous()
            out = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
            engine.search_keys_device(q.data_ptr(), q.shape[0], scope, k, out.data_ptr(),
                                      torch.cuda.current_stream(q.device).cuda_stream)
            return out

        def merge(gathered: torch.Tensor, k: int) -> torch.Tensor:
            n_lists, nq, _ = gathered.shape
            out = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
            engine.merge_topk_device(gathered.data_ptr(), n_lists, nq, k, out.data_ptr(),
                                     torch.cuda.current_stream(gathered.device).cuda_stream)
            return out

        return cls(local_search, merge, group, local_df=engine_local_df(engine), weights_fn=engine.idf_weights)

    def scope_weights(self, scope: Sequence[int]):
        """uint8 [256] RF-1w weights from the corpus-wide document frequencies (see allreduce_scope_weights)."""
        if self.local_df is None or self.weights_fn is None:
            raise RuntimeError("this searcher was built without the document-frequency callables")
        return allreduce_scope_weights(self.local_df, self.weights_fn, scope, self.group)

    def search_keys(self, q: torch.Tensor, scope: Sequence[int], k: int = 10, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """q int8 [nq, 256] (replicated on every rank) -> merged packed keys int64 [nq, k] on every rank."""
        local = self.local_search(q, scope, k)
        if self.world == 1:
            if out is not None:
                out.copy_(local)
                return out
            return local
        shape = (self.world, local.shape[0], k)
        if self._gather_buf is None or tuple(self._gather_buf.shape) != shape or self._gather_buf.device != local.device:
            self._gather_buf = torch.empty(shape, dtype=torch.int64, device=local.device)
        # rank-major concatenation along dim 0 (the form both NCCL and gloo accept)
        dist.all_gather_into_tensor(self._gather_buf.view(shape[0] * shape[1], k), local.contiguous(), group=self.group)
        merged = self.merge(self._gather_buf, k)
        if out is not None:
            out.copy_(merged)
            return out
        return merged

    def search(self, q: torch.Tensor, scope: Sequence[int], k: int = 10):
        return unpack_keys_torch(self.search_keys(q, scope, k))


class StoreShardedSearcher:
    """Whole stores per rank (multi-tenant corpora, BASELINE.json configs[4]): global
    store number g lives on rank g % world, so a query scoped to one store is scored by exactly one
    GPU and the corpus needs no data-path collective.  Every rank receives the query batch, scores
    the part of each scope it owns in ONE launch (`rf_search_keys_device_scoped`; a query with
    nothing local costs an empty block), and the nq x k packed keys are all-gathered and merged so
    every rank holds the answer -- the same exchange as the chunk-sharded path, which also makes a
    scope that spans ranks (a user's stores on several GPUs) exact.

    Chunk ids must be unique across ranks: build each rank's Engine with a distinct `id_base`
    (`id_base_for`).  Stores are numbered in creation order; every rank calls `open_store` with the
    same names in the same order (SPMD), only the owner creates the segment."""

    def __init__(self, local_search: Callable[[torch.Tensor, Sequence[Sequence[int]], int], torch.Tensor],
                 merge: Callable[[torch.Tensor, int], torch.Tensor], open_local: Callable[[str], int],
                 group: Optional[dist.ProcessGroup] = None):
        self.local_search = local_search
        self.merge = merge
        self.open_local = open_local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.by_name: dict = {}          # store name -> global store number
        self.local_seg: dict = {}        # global store number -> this rank's engine segment (owned stores only)
        self._gather_buf: Optional[torch.Tensor] = None

    @staticmethod
    def id_base_for(rank: int, world: int) -> int:
        """Disjoint chunk-id ranges per rank (ids are 32-bit, 0xFFFFFFFF is reserved)."""
        return rank * ((1 << 32) // max(1, world))

    @classmethod
    def for_engine(cls, engine, group: Optional[dist.ProcessGroup] = None) -> "StoreShardedSearcher":
        def local_search(q: torch.Tensor, scopes, k: int) -> torch.Tensor:
            assert q.is_cuda and q.dtype == torch.int8 and q.is_contiguous()
            out = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
            engine.search_keys_device_scoped(q.data_ptr(), q.shape[0], scopes, k, out.data_ptr(),
                                             torch.cuda.current_stream(q.device).cuda_stream)
            return out

        def merge(gathered: torch.Tensor, k: int) -> torch.Tensor:
            n_lists, nq, _ = gathered.shape
            out = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
            engine.merge_topk_device(gathered.data_ptr(), n_lists, nq, k, out.data_ptr(),
                                     torch.cuda.current_stream(gathered.device).cuda_stream)
            return out

        return cls(local_search, merge, engine.open_store, group)

    def owner(self, store: int) -> int:
        return store % self.world

    def open_store(self, name: str) -> int:
        """-> global store number (idempotent).  Collective in the SPMD sense: same calls on every rank."""
        g = self.by_name.get(name)
        if g is None:
            g = self.by_name[name] = len(self.by_name)
            if self.owner(g) == self.rank:
                self.local_seg[g] = self.open_local(name)
        return g

    def local_scopes(self, scopes: Sequence[Sequence[int]]):
        """Per query: the engine segments of the scope's stores this rank owns (possibly none)."""
        return [[self.local_seg[g] for g in dict.fromkeys(sc) if g in self.local_seg] for sc in scopes]

    def prepare(self, scopes: Sequence[Sequence[int]]):
        """Resolve a batch's scopes to this rank's CSR once (a serving loop reuses it across calls)."""
        from .engine import scopes_to_csr
        return scopes_to_csr(self.local_scopes(scopes))

    def search_keys(self, q: torch.Tensor, scopes, k: int = 10) -> torch.Tensor:
        """q int8 [nq, 256] (replicated); scopes: global store numbers per query, or the result of
        `prepare` -> packed keys int64 [nq, k] on every rank."""
        local = self.local_search(q, scopes if isinstance(scopes, tuple) else self.local_scopes(scopes), k)
        if self.world == 1:
            return local
        shape = (self.world, local.shape[0], k)
        if self._gather_buf is None or tuple(self._gather_buf.shape) != shape or self._gather_buf.device != local.device:
            self._gather_buf = torch.empty(shape, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(self._gather_buf.view(shape[0] * shape[1], k), local.contiguous(), group=self.group)
        return self.merge(self._gather_buf, k)

    def search(self, q: torch.Tensor, scopes: Sequence[Sequence[int]], k: int = 10):
        return unpack_keys_torch(self.search_keys(q, scopes, k))


class FusedStoreShardedSearcher(StoreShardedSearcher):
    """StoreShardedSearcher with the exchange inside the kernels (rf_search_keys_device_scoped_fused): each rank's
    scan stores every query's k keys into every rank's gather buffer over NVLink (symmetric memory) and releases
    a flag; a one-warp-per-query kernel behind it acquires the flags and merges.  Two launches per batch on each
    rank, no NCCL collective on the data path; the per-query plans come from the engine's device-resident store
    table, so the host sends only the scope lists.  Same results as the all-gather path."""

    def __init__(self, engine, nq_cap: int = 1024, k: int = 10, group: Optional[dist.ProcessGroup] = None):
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        base = StoreShardedSearcher.for_engine(engine, group)
        super().__init__(base.local_search, base.merge, base.open_local, group)
        self.engine = engine
        self.k, self.nq_cap = k, nq_cap
        pg = group or dist.group.WORLD
        if self.world > 8:
            raise ValueError("the fused exchange serves the GPUs of one box (world <= 8)")
        dev = torch.device("cuda", torch.cuda.current_device())
        self._keys = symm_mem.empty((4 * self.world * nq_cap * k,), dtype=torch.int64, device=dev)
        self._flags = symm_mem.empty((4 * self.world * nq_cap,), dtype=torch.int32, device=dev)
        self._keys.zero_()
        self._flags.zero_()
        self._hk = symm_mem.rendezvous(self._keys, pg)
        self._hf = symm_mem.rendezvous(self._flags, pg)
        self._keys_ptrs = np.asarray([int(p) for p in self._hk.buffer_ptrs], dtype=np.uint64)
        self._flag_ptrs = np.asarray([int(p) for p in self._hf.buffer_ptrs], dtype=np.uint64)
        self._timeout = torch.zeros(1, dtype=torch.int32, device=dev)
        self._seq = 0
        torch.cuda.synchronize(dev)
        dist.barrier(pg)          # every rank has zeroed its flags before anyone publishes

    def prepare_fused(self, scopes: Sequence[Sequence[int]]):
        """Resolve a batch's scopes (global store numbers) once: the queries THIS rank has stores for and their local
        segments (CSR), and for every query of the batch the ranks that own part of its scope (what its merge waits
        for; identical on every rank)."""
        import numpy as np
        from .engine import scopes_to_csr
        q_index, local, masks = [], [], np.zeros(len(scopes), np.uint8)
        for i, sc in enumerate(scopes):
            mine = []
            for g in dict.fromkeys(sc):
                if g < 0 or g >= len(self.by_name):
                    continue
                masks[i] |= 1 << self.owner(g)
                if g in self.local_seg:
                    mine.append(self.local_seg[g])
            if mine:
                q_index.append(i)
                local.append(mine)
        segs, off = scopes_to_csr(local) if local else (np.zeros(1, np.uint32), np.zeros(1, np.uint32))
        return ("fused", segs, off, np.asarray(q_index, np.uint32), masks)

    def search_keys(self, q: torch.Tensor, scopes, k: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """q int8 [nq, 256] (the whole batch, replicated on every rank); scopes: global store numbers per query, or the
        result of `prepare_fused` -> packed keys int64 [nq, k] on every rank."""
        k = k or self.k
        prepared = scopes if (isinstance(scopes, tuple) and len(scopes) == 5 and scopes[0] == "fused") else None
        if q.shape[0] > self.nq_cap or k != self.k:
            return super().search_keys(q, scopes, k)
        if prepared is None:
            prepared = self.prepare_fused(scopes)
        assert q.is_cuda and q.dtype == torch.int8 and q.is_contiguous()
        if out is None:
            out = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
        self._seq += 1
        _, segs, off, q_index, masks = prepared
        self.engine.search_keys_device_scoped_fused(q.data_ptr(), int(q_index.size), (segs, off), k, out.data_ptr(),
                                                    torch.cuda.current_stream(q.device).cuda_stream, self.rank, self.world,
                                                    self.nq_cap, self._seq, self._keys_ptrs, self._flag_ptrs, self._timeout.data_ptr(),
                                                    nq_total=q.shape[0], q_index=q_index, owner_masks=masks)
        return out

    def timed_out(self) -> bool:
        return bool(self._timeout.item())


class FusedShardedSearcher:
    """Sharded search with the exchange fused into the scan kernel: its finishing block stores the
    rank's top-k straight into every rank's gather buffer over NVLink (symmetric memory from
    torch.distributed._symmetric_memory), releases a flag, acquires the peers' flags and merges --
    compute and collective in ONE kernel, no NCCL launch on the data path.  Same results as
    ShardedSearcher (the packed key is a total order)."""

    # Batches of this many queries or more go through the collective path instead: its local search may
    # take the tensor-core kernels (one pass over the shard for the whole batch), which the fused scan
    # kernel -- one pass per query -- cannot match (8 queries over a 12.5 M-chunk shard: ~1 ms vs ~3.8 ms).
    BATCH_THRESHOLD = 4

    def __init__(self, engine, nq_cap: int = 64, k: int = 10, group: Optional[dist.ProcessGroup] = None):
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        self.engine = engine
        self.k = k
        self.nq_cap = nq_cap
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("the fused exchange serves the GPUs of one box (world <= 8)")
        dev = torch.device("cuda", torch.cuda.current_device())
        self._keys = symm_mem.empty((4 * self.world * nq_cap * k,), dtype=torch.int64, device=dev)
        self._flags = symm_mem.empty((4 * self.world * nq_cap,), dtype=torch.int32, device=dev)
        self._keys.zero_()
        self._flags.zero_()
        self._hk = symm_mem.rendezvous(self._keys, self.group)
        self._hf = symm_mem.rendezvous(self._flags, self.group)
        self._keys_ptrs = np.asarray([int(p) for p in self._hk.buffer_ptrs], dtype=np.uint64)
        self._flag_ptrs = np.asarray([int(p) for p in self._hf.buffer_ptrs], dtype=np.uint64)
        self._timeout = torch.zeros(1, dtype=torch.int32, device=dev)
        self._seq = 0
        self._batched = ShardedSearcher.for_engine(engine, self.group)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)          # every rank has zeroed its flags before anyone publishes

    def search_keys(self, q: torch.Tensor, scope: Sequence[int], k: Optional[int] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        k = k or self.k
        if q.shape[0] >= self.BATCH_THRESHOLD or q.shape[0] > self.nq_cap or k != self.k:
            return self._batched.search_keys(q, scope, k, out=out)
        assert q.is_cuda and q.dtype == torch.int8 and q.is_contiguous()
        if out is None:
            out = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
        self._seq += 1
        self.engine.search_keys_device_fused(q.data_ptr(), q.shape[0], scope, k, out.data_ptr(),
                                             torch.cuda.current_stream(q.device).cuda_stream, self.rank, self.world,
                                             self.nq_cap, self._seq, self._keys_ptrs, self._flag_ptrs, self._timeout.data_ptr())
        return out

    def scope_weights(self, scope: Sequence[int]):
        return allreduce_scope_weights(engine_local_df(self.engine), self.engine.idf_weights, scope, self.group)

    def timed_out(self) -> bool:
        return bool(self._timeout.item())

    def search(self, q: torch.Tensor, scope: Sequence[int], k: Optional[int] = None):
        return unpack_keys_torch(self.search_keys(q, scope, k))
