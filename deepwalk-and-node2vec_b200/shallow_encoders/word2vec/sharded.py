"""
Embedding tables striped over the HBM of several B200s (one process per GPU) and addressed as ONE flat array.

The reference trains on a single device (`devices: '1'`, configs/sge_sg_cora.yaml:30) with both tables in one
`nn.Embedding` each (word2vec/model.py:22-23).  Here `W2VBase`'s two tables can instead be `ShardedTable`s: a flat
fp32 [vocab x emb] virtual range in every process, whose stripes (2 MiB by default) live round-robin on the G GPUs
and are mapped into every peer over NVLink / NVSwitch (CUDA virtual memory management, csrc/shard.cu).  The fused
SGNS kernel is handed the flat pointer and runs unchanged: its row gathers and `red.global.add.v4.f32` scatters reach
the owner GPU's L2 directly, so there is no separate all-to-all of rows and gradients, no staging buffer and no
barrier between GPUs -- all GPUs do Hogwild SGD on one model.

Host plumbing only (pointers, file descriptors, sockets); every kernel is behind the C ABI (`_native`).
"""
import math
import os
import socket
import struct
from typing import List, Optional

import torch

from shallow_encoders import _native as nat

_FD_BATCH = 64


class _Spec:
    __slots__ = ('world', 'rank', 'stripe_rows')

    def __init__(self, world, rank, stripe_rows):
        self.world, self.rank, self.stripe_rows = world, rank, stripe_rows


def stripe_owner(stripe: int, world: int) -> int:
    return stripe % world


def local_rows(vocab: int, stripe_rows: int, world: int, rank: int) -> int:
    """Rows of [0, vocab) whose stripe is owned by `rank` (python restatement of se_shard_local_rows)."""
    n_stripes = -(-vocab // stripe_rows)
    total = 0
    for s in range(rank, n_stripes, world):
        total += min(stripe_rows, vocab - s * stripe_rows)
    return total


def local_to_global(j, stripe_rows: int, world: int, rank: int):
    """Local row id (0 .. local_rows) of `rank` -> table row: what the kernel does for local negatives."""
    return ((j // stripe_rows) * world + rank) * stripe_rows + j % stripe_rows


class FdExchange:
    """All-pairs exchange of file descriptors between the ranks of one node over abstract unix sockets (SCM_RIGHTS).
    `barrier` is any callable that returns once every rank has reached it (torch.distributed.barrier)."""

    def __init__(self, rank: int, world: int, token: str, barrier):
        self.rank, self.world = rank, world
        self.peers = {}
        if world == 1:
            return
        name = lambda r: f'\0se_b200_{token}_{r}'   # noqa: E731  (abstract namespace: nothing to unlink)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(name(rank))
        srv.listen(world)
        barrier()
        for peer in range(rank + 1, world):          # the lower rank connects, the higher accepts
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.connect(name(peer))
            s.sendall(struct.pack('i', rank))
            self.peers[peer] = s
        for _ in range(rank):
            s, _addr = srv.accept()
            (peer,) = struct.unpack('i', self._recv_exact(s, 4))
            self.peers[peer] = s
        srv.close()
        barrier()

    @staticmethod
    def _recv_exact(s, n):
        buf = b''
        while len(buf) < n:
            chunk = s.recv(n - len(buf))
            if not chunk:
                raise RuntimeError('peer closed the fd-exchange socket')
            buf += chunk
        return buf

    def send(self, peer: int, fds: List[int]):
        socket.send_fds(self.peers[peer], [b'F'], list(fds))

    def recv(self, peer: int, n: int) -> List[int]:
        msg, fds, _flags, _addr = socket.recv_fds(self.peers[peer], 1, n)
        if msg != b'F' or len(fds) != n:
            raise RuntimeError(f'fd exchange with rank {peer}: expected {n} descriptors, got {len(fds)}')
        return list(fds)

    def close(self):
        for s in self.peers.values():
            s.close()
        self.peers = {}


def make_exchange(rank: int, world: int, group=None) -> FdExchange:
    """FdExchange whose socket names are unique to this job (token broadcast from rank 0)."""
    import torch.distributed as dist
    if world == 1:
        return FdExchange(0, 1, '', lambda: None)
    tok = [f'{os.getpid()}_{int.from_bytes(os.urandom(4), "little")}' if rank == 0 else None]
    dist.broadcast_object_list(tok, src=0, group=group)
    return FdExchange(rank, world, tok[0], lambda: dist.barrier(group=group))


class ShardedTable:
    """Flat fp32 [vocab x emb] table striped over `world` GPUs; stripe s lives on rank s % world.

    world == 1, or simulate=True (every stripe is created on the calling GPU but the table still reports
    (world, rank): single-GPU tests of the sharding arithmetic), needs no exchange."""

    def __init__(self, vocab: int, emb: int, device, rank: int = 0, world: int = 1, exchange: Optional[FdExchange] = None,
                 stripe_bytes: Optional[int] = None, simulate: bool = False, max_stripes: int = 4096):
        import ctypes
        self.vocab, self.emb = int(vocab), int(emb)
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world)
        self.shape = (self.vocab, self.emb)
        self._handles, self._mapped, self.ptr, self._total = [], [], 0, 0
        lib = nat.load()
        with torch.cuda.device(self.device):
            gran = ctypes.c_int64()
            nat._check(lib.se_shard_granularity(ctypes.byref(gran)))
            row_bytes = 4 * self.emb
            nbytes = self.vocab * row_bytes
            if stripe_bytes is None:
                stripe_bytes = gran.value * row_bytes // math.gcd(gran.value, row_bytes)     # whole rows per stripe
                while -(-nbytes // stripe_bytes) > max_stripes:
                    stripe_bytes *= 2
            if stripe_bytes % gran.value or stripe_bytes % row_bytes:
                raise ValueError(f'stripe_bytes {stripe_bytes} must be a multiple of the allocation granularity '
                                 f'{gran.value} and of the row size {row_bytes}')
            self.stripe_bytes = int(stripe_bytes)
            self.stripe_rows = self.stripe_bytes // row_bytes
            self.n_stripes = -(-nbytes // self.stripe_bytes)
            self._total = self.n_stripes * self.stripe_bytes
            va = ctypes.c_uint64()
            nat._check(lib.se_shard_reserve(self._total, ctypes.byref(va)))
            self.ptr = va.value

            def create_and_map(s):
                h = ctypes.c_uint64()
                nat._check(lib.se_shard_create(self.stripe_bytes, ctypes.byref(h)))
                self._handles.append(h.value)
                nat._check(lib.se_shard_map(self.ptr + s * self.stripe_bytes, self.stripe_bytes, h.value))
                self._mapped.append(s)
                return h.value

            mine = [s for s in range(self.n_stripes) if simulate or world == 1 or stripe_owner(s, world) == rank]
            own_handles = [create_and_map(s) for s in mine]
            if world > 1 and not simulate:
                if exchange is None:
                    raise ValueError('a multi-GPU ShardedTable needs an FdExchange (make_exchange)')
                self._exchange_stripes(lib, exchange, own_handles)
            torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------------------------
    def _exchange_stripes(self, lib, ex: FdExchange, own_handles: List[int]):
        import ctypes
        world, rank = self.world, self.rank
        per_rank = [list(range(r, self.n_stripes, world)) for r in range(world)]
        rounds = -(-max(len(v) for v in per_rank) // _FD_BATCH)
        for k in range(rounds):
            lo, hi = k * _FD_BATCH, (k + 1) * _FD_BATCH
            fds = []
            for h in own_handles[lo:hi]:
                fd = ctypes.c_int()
                nat._check(lib.se_shard_export_fd(h, ctypes.byref(fd)))
                fds.append(fd.value)
            for peer in range(world):
                if peer != rank and fds:
                    ex.send(peer, fds)
            for peer in range(world):
                if peer == rank:
                    continue
                theirs = per_rank[peer][lo:hi]
                if not theirs:
                    continue
                got = ex.recv(peer, len(theirs))
                for s, fd in zip(theirs, got):
                    h = ctypes.c_uint64()
                    nat._check(lib.se_shard_import_fd(fd, ctypes.byref(h)))
                    os.close(fd)
                    self._handles.append(h.value)
                    nat._check(lib.se_shard_map(self.ptr + s * self.stripe_bytes, self.stripe_bytes, h.value))
                    self._mapped.append(s)
            for fd in fds:
                os.close(fd)

    # ------------------------------------------------------------------------------------------------------------
    def spec(self) -> _Spec:
        return _Spec(self.world, self.rank, self.stripe_rows)

    @property
    def is_cuda(self) -> bool:
        return True

    def local_rows(self) -> int:
        return local_rows(self.vocab, self.stripe_rows, self.world, self.rank)

    def owned_rows(self) -> torch.Tensor:
        """int64 ids of the rows whose stripe this rank owns, ascending (== local id order of the negative sampler)."""
        j = torch.arange(self.local_rows(), dtype=torch.int64, device=self.device)
        return local_to_global(j, self.stripe_rows, self.world, self.rank)

    def fill_uniform(self, bound: float, seed: int) -> None:
        nat.table_fill_uniform(self, bound, seed)

    def gather(self, rows: torch.Tensor) -> torch.Tensor:
        return nat.table_gather_rows(self, rows.to(self.device, torch.int64))

    def scatter(self, rows: torch.Tensor, src: torch.Tensor) -> None:
        nat.table_scatter_rows(self, rows.to(self.device, torch.int64), src.to(self.device, torch.float32).contiguous())

    def load_owned(self, full: torch.Tensor) -> None:
        """Copy this rank's rows out of a full [vocab x emb] tensor (host or device) into the table."""
        rows = self.owned_rows()
        self.scatter(rows, full[rows.to(full.device)].to(self.device))

    def to_tensor(self, chunk_rows: int = 1 << 20) -> torch.Tensor:
        """Dense copy of the WHOLE table on this GPU (reads peer stripes over NVLink)."""
        out = torch.empty((self.vocab, self.emb), dtype=torch.float32, device=self.device)
        for lo in range(0, self.vocab, chunk_rows):
            hi = min(self.vocab, lo + chunk_rows)
            nat.table_gather_rows(self, torch.arange(lo, hi, dtype=torch.int64, device=self.device), out=out[lo:hi])
        return out

    def as_rank(self, rank: int) -> '_RankView':
        """The same memory seen as another rank of a SIMULATED sharding (single-GPU tests of the per-rank kernels)."""
        return _RankView(self, rank)

    def close(self) -> None:
        if not self.ptr:
            return
        lib = nat.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for s in self._mapped:
                lib.se_shard_unmap(self.ptr + s * self.stripe_bytes, self.stripe_bytes)
            for h in self._handles:
                lib.se_shard_release(h)
            lib.se_shard_unreserve(self.ptr, self._total)
        self._mapped, self._handles, self.ptr = [], [], 0

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001  (interpreter shutdown)
            pass


class _RankView:
    """Borrowed view of a ShardedTable that reports a different rank; owns nothing."""

    def __init__(self, table: ShardedTable, rank: int):
        assert 0 <= rank < table.world
        self._table, self.rank = table, int(rank)
        self.ptr, self.vocab, self.emb, self.device, self.world, self.stripe_rows = (table.ptr, table.vocab, table.emb, table.device,
                                                                                   table.world, table.stripe_rows)

    def spec(self) -> _Spec:
        return _Spec(self.world, self.rank, self.stripe_rows)


class _CudaArray:
    """__cuda_array_interface__ carrier: lets torch view memory this module mapped itself (no ownership)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {'shape': tuple(int(x) for x in shape), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class ReplicatedTable:
    """One [vocab x emb] table trained by `world` GPUs with the reference's GLOBAL negative draw at NVLink bulk rate.

    Every GPU holds a full WORKING COPY in its own HBM (the unchanged single-GPU kernels run on it: pass this object
    wherever a table is expected), rank r holds the MASTER of the contiguous row chunk r.  `sync()` -- called by every
    rank after its step, between two barriers -- adds the summed updates of all copies to the master and writes the
    result back into every copy, one kernel over peer memory (csrc/replica.cu): synchronous data-parallel SGD with summed
    updates on ONE model.  All copies live in one virtual range (segment g on GPU g, mapped everywhere) built with the
    striped-table plumbing above: a ShardedTable with one stripe per rank.

    simulate=True places every segment on the calling GPU (single-GPU tests); `as_rank(r)` then gives rank r's view."""

    def __init__(self, vocab: int, emb: int, device, rank: int = 0, world: int = 1, exchange: Optional[FdExchange] = None,
                 simulate: bool = False, group=None):
        import ctypes
        self.vocab, self.emb = int(vocab), int(emb)
        self.device = torch.device(device)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.shape = (self.vocab, self.emb)
        lib = nat.load()
        with torch.cuda.device(self.device):
            gran = ctypes.c_int64()
            nat._check(lib.se_shard_granularity(ctypes.byref(gran)))
        row_bytes = 4 * self.emb
        unit = gran.value * row_bytes // math.gcd(gran.value, row_bytes)
        self.seg_bytes = -(-(self.vocab * row_bytes) // unit) * unit
        self.seg_rows = self.seg_bytes // row_bytes
        self.seg_elems = self.seg_bytes // 4
        self._set = ShardedTable(self.world * self.seg_rows, self.emb, self.device, self.rank, self.world, exchange,
                                 stripe_bytes=self.seg_bytes, simulate=simulate)
        self.base = self._set.ptr
        self.ptr = self.base + self.rank * self.seg_bytes
        self.n_elems = self.vocab * self.emb
        self._masters = {}
        self.master = self._master_for(self.rank)

    def _master_for(self, rank: int) -> torch.Tensor:
        if rank not in self._masters:
            lo, hi = nat.replica_chunk(self.n_elems, self.world, rank)
            self._masters[rank] = torch.zeros(max(hi - lo, 4), dtype=torch.float32, device=self.device)
        return self._masters[rank]

    # the kernels see a plain local table (device-scope reductions, negatives over the whole table)
    def spec(self) -> _Spec:
        return _Spec(1, 0, max(self.vocab, 1))

    @property
    def is_cuda(self) -> bool:
        return True

    def owned_row_range(self, rank: Optional[int] = None):
        lo, hi = nat.replica_chunk(self.n_elems, self.world, self.rank if rank is None else rank)
        return lo // self.emb, -(-min(hi, self.n_elems) // self.emb)

    def view(self) -> torch.Tensor:
        """The working copy as a torch tensor that ALIASES the mapped memory (valid until close())."""
        return torch.as_tensor(_CudaArray(self.ptr, self.shape), device=self.device)

    def to_tensor(self) -> torch.Tensor:
        rows = torch.arange(self.vocab, dtype=torch.int64, device=self.device)
        return nat.table_gather_rows(self, rows)

    def gather(self, rows: torch.Tensor) -> torch.Tensor:
        return nat.table_gather_rows(self, rows.to(self.device, torch.int64))

    def scatter(self, rows: torch.Tensor, src: torch.Tensor) -> None:
        nat.table_scatter_rows(self, rows.to(self.device, torch.int64), src.to(self.device, torch.float32).contiguous())

    def fill_uniform(self, bound: float, seed: int) -> None:
        """Every rank fills its own copy with the same Philox content, then takes its master chunk from it."""
        nat.table_fill_uniform(self, bound, seed)
        self.adopt()

    def load(self, full: torch.Tensor) -> None:
        """Every rank loads the same full [vocab x emb] tensor into its copy (checkpoint restore)."""
        self.scatter(torch.arange(self.vocab, dtype=torch.int64, device=self.device), full.to(self.device))
        self.adopt()

    def adopt(self, rank: Optional[int] = None) -> None:
        """master chunk <- working copy (mode 1)."""
        r = self.rank if rank is None else rank
        nat.replica_sync(self.base, self.seg_elems, self.world, r, self.n_elems, self._master_for(r), mode=1)

    def sync_local(self, rank: Optional[int] = None, mode: int = 0, beta: float = 1.0) -> None:
        """This rank's share of the fused reduce-scatter + all-gather; the caller provides the barriers (see `sync_replicated`)."""
        r = self.rank if rank is None else rank
        nat.replica_sync(self.base, self.seg_elems, self.world, r, self.n_elems, self._master_for(r), mode=mode, beta=beta)

    def as_rank(self, rank: int) -> '_ReplicaView':
        return _ReplicaView(self, rank)

    def close(self) -> None:
        self._set.close()
        self.ptr = self.base = 0


class _ReplicaView:
    """Working copy of another rank of a SIMULATED ReplicatedTable (single-GPU tests); owns nothing."""

    def __init__(self, table: ReplicatedTable, rank: int):
        assert 0 <= rank < table.world
        self.ptr = table.base + rank * table.seg_bytes
        self.vocab, self.emb, self.device, self.shape = table.vocab, table.emb, table.device, table.shape

    def spec(self) -> _Spec:
        return _Spec(1, 0, max(self.vocab, 1))


_BARRIER_TOKEN = {}


def device_barrier(device, group=None) -> None:
    """Stream-ordered barrier over the ranks of `group`: a one-element NCCL all-reduce on the current stream (no host sync)."""
    import torch.distributed as dist
    key = (torch.device(device).index, id(group))
    if key not in _BARRIER_TOKEN:
        _BARRIER_TOKEN[key] = torch.zeros(1, dtype=torch.int32, device=device)
    dist.all_reduce(_BARRIER_TOKEN[key], group=group)


def merge_weight(merge, world: int) -> float:
    """'sum' -> 1, 'mean' -> 1 / world, 'stable' -> min(1, 2 / world), or a number in (0, 1].
    'stable' is the largest weight for which merging G parallel chains cannot overshoot on a quadratic: a chain that closes a
    fraction rho of the distance to its optimum moves the merged model by G beta rho, which stays below 2 for every rho <= 1 iff
    beta <= 2 / G.  'sum' (beta = 1) is exact data-parallel SGD but diverges once G rho > 2 (measured: 8 GPUs, profiles/r02_multi_gpu_accuracy.json)."""
    if merge == 'sum':
        return 1.0
    if merge == 'stable':
        return min(1.0, 2.0 / world)
    if merge == 'mean':
        return 1.0 / world
    return float(merge)


def sync_replicated(tables, group=None, merge='stable') -> None:
    """End-of-step synchronisation of ReplicatedTables on every rank: barrier (all steps done) -> each rank's fused
    reduce-scatter + all-gather kernels over peer memory -> barrier (all copies written).  world == 1: nothing to do.
    merge: how the GPUs' updates since the last sync combine (see merge_weight) -- 'stable' (default: weight min(1, 2/G)), 'sum'
    (synchronous SGD with summed updates: every GPU's progress counts in full; fine on 2 GPUs, diverges on 8 when a step updates
    every row many times), 'mean' (local SGD with model averaging: loses 1 - 1/G of the progress per step), or a weight in (0, 1]."""
    tables = [t for t in tables if t.world > 1]
    if not tables:
        return
    device_barrier(tables[0].device, group)
    for t in tables:
        t.sync_local(beta=merge_weight(merge, t.world))
    device_barrier(tables[0].device, group)


def sgns_update_walks_owner_computes(w_in, w_out, my_walks: torch.Tensor, radius: int, n_neg: int, row_offset: int, lr: float, seed: int,
                                     centre_id_base: int, rank: int, world: int, group=None, stats: Optional[torch.Tensor] = None,
                                     alias=None, gather_buf: Optional[torch.Tensor] = None, micro_walks: Optional[int] = None,
                                     grouped: bool = True) -> None:
    """One multi-GPU SGNS step with the reference's GLOBAL negative distribution and little NVLink traffic.  The walks of all ranks
    are all-gathered (4 bytes per token, the only collective); `centre_id_base` is the Philox id of centre 0 of RANK 0's walks, rank
    r's centres follow at r * n_walks * n_centres, so the negatives are the ones a single GPU would draw for the concatenated batch.
    Every rank must call this with the same number of walks.

    grouped (default): every rank buckets the centres of the gathered batch by table row and computes EVERY pair -- positive or
    negative -- whose W_out row it owns (`se_sgns_update_pairs_owned`): W_out never crosses NVLink, a W_in centre row crosses once per
    run of equal rows.  grouped=False (round 1): the positive pairs of this rank's walks run in the window-resident kernel (n_neg = 0,
    context rows fetched from their owners) and every rank processes the owned negatives of every centre in walk order
    (`se_sgns_update_negatives_owned`); a step then applies positives, then negatives.

    `micro_walks` processes the step in slices of that many walks per rank (after ONE all-gather), for steps that update the same
    rows many times (small tables, large batches)."""
    import torch.distributed as dist
    n_walks, length = my_walks.shape
    n_cen = length - 2 * radius
    if world > 1:
        if gather_buf is None:
            gather_buf = torch.empty((world * n_walks, length), dtype=torch.int32, device=my_walks.device)
        dist.all_gather_into_tensor(gather_buf, my_walks.contiguous(), group=group)
        all_walks = gather_buf
    else:
        all_walks = my_walks
    micro = n_walks if not micro_walks else max(1, min(int(micro_walks), n_walks))
    for lo in range(0, n_walks, micro):
        hi = min(n_walks, lo + micro)
        if grouped:
            if micro == n_walks:
                nat.sgns_update_pairs_owned(w_in, w_out, all_walks, radius, n_neg, row_offset, lr, seed, centre_id_base=centre_id_base,
                                            alias=alias, stats=stats, positives=True)
            else:
                for r in range(world):   # the same slice of every rank's walks (contiguous inside rank r's block of the gathered buffer)
                    nat.sgns_update_pairs_owned(w_in, w_out, all_walks[r * n_walks + lo:r * n_walks + hi], radius, n_neg, row_offset, lr, seed,
                                                centre_id_base=centre_id_base + (r * n_walks + lo) * n_cen, alias=alias, stats=stats,
                                                positives=True)
            continue
        nat.sgns_update_walks(w_in, w_out, my_walks[lo:hi], radius, 0, row_offset, lr, seed,
                              centre_id_base=centre_id_base + (rank * n_walks + lo) * n_cen, stats=stats)
        if micro == n_walks:
            nat.sgns_update_negatives_owned(w_in, w_out, all_walks, radius, n_neg, row_offset, lr, seed, centre_id_base=centre_id_base,
                                            alias=alias, stats=stats, grouped=False)
        else:
            for r in range(world):
                nat.sgns_update_negatives_owned(w_in, w_out, all_walks[r * n_walks + lo:r * n_walks + hi], radius, n_neg, row_offset, lr, seed,
                                                centre_id_base=centre_id_base + (r * n_walks + lo) * n_cen, alias=alias, stats=stats,
                                                grouped=False)
