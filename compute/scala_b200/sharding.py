"""Leading-axis sharding across the GPUs of one box, one process per GPU.

Row-major tensors shard into contiguous row blocks, so elementwise graphs and views that do not mix the leading axis
need no exchange at all.  Only two things cross NVLink: partial reductions (a scalar, or one vector of column sums)
and, on request, the row blocks of a sharded result.  The collectives are the library's own (cc_comm_* over NCCL,
include/compute_cuda.h); `torch.distributed` — any backend — is used for nothing but handing rank 0's NCCL unique id to
the other ranks.

Since round 2 the sharding itself lives behind the Tensor API (`Tensor.shard()` / `.gather()`, include/compute_cuda.h: ct_shard,
ct_gather, cc_shard_*): a row block carries its distribution through the lazy graph, `shard.sum()` is the global sum, the fold of
`shard.split(0)` is all-reduced when evaluated, and the split / broadcast / sum matmul over a row block of A gathers from the
contraction's own epilogue. `Communicator` sets the communicator up and keeps the explicit, buffer-level routes for A/B timing.

Host-side logic only: nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Sequence


def shard_rows(rows: int, world: int, rank: int) -> tuple[int, int]:
    """(first row, row count) of `rank`'s block: the first `rows % world` ranks get one extra row"""
    if not (0 <= rank < world) or rows < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(rows, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_shape(shape: Sequence[int], world: int, rank: int) -> list[int]:
    _, n = shard_rows(int(shape[0]), world, rank)
    return [n] + [int(s) for s in shape[1:]]


def shard_offsets(shape: Sequence[int], world: int) -> list[int]:
    """flat element offset of every rank's block (+ the total), for gathering"""
    inner = 1
    for s in shape[1:]:
        inner *= int(s)
    return [shard_rows(int(shape[0]), world, r)[0] * inner for r in range(world)] + [int(shape[0]) * inner]


def exchange_unique_id(dist, make_id) -> bytes:
    """rank 0 creates the NCCL unique id (make_id()), everybody receives it through torch.distributed"""
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


class Communicator:
    """the process' NCCL communicator inside libcompute_cuda.so"""

    def __init__(self, cuda, dist=None, peer: bool = True):
        self.cuda = cuda
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.peer = False
        if self.world > 1:
            uid = exchange_unique_id(dist, cuda.comm_unique_id)
            cuda.comm_init(uid, self.world, self.rank)
            if peer:
                # NVLink peer mailboxes: our own one-shot / fused collectives for the small combines (NCCL stays for the rest)
                cuda.comm_enable_peer()
                self.peer = cuda.comm_peer_enabled()

    def _require_equal_blocks(self, n_floats: int) -> None:
        """gathers need the same block size on every rank; checked collectively once per size, so that uneven shards are an
        IllegalArgument on every rank instead of a hang inside the exchange"""
        seen = self.__dict__.setdefault("_agreed", set())
        if self.world > 1 and n_floats not in seen:
            if not self.cuda.shard_agree(n_floats):
                raise ValueError(f"gather needs equal blocks on every rank (this rank holds {n_floats} floats): pad the leading axis to a "
                                 "multiple of the number of ranks, or leave the result sharded")
            seen.add(n_floats)

    # ---- the same three exchanges through the Tensor API (what user code looks like) ----------------------------------------

    @staticmethod
    def fold(parts):
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    def sharded(self, local_tensor):
        """this rank's row block as a sharded tensor (identity semantics at world size 1)"""
        return local_tensor.shard()

    def matmul_pattern(self, a_block, b):
        """benchmarks.scala:188-191 verbatim on a row block of A and a replicated B: one tcgen05 contraction per rank, C stays a row block"""
        i, j = a_block.shape
        j2, k = b.shape
        assert j == j2
        product = a_block.broadcast([i, j, k]) * b.reshape([1, j, k]).broadcast([i, j, k])
        return self.fold(product.split(1))

    def route_peer(self, on: bool) -> None:
        """A/B switch: small combines over the NVLink peer mailboxes (True) or over NCCL (False)"""
        if self.world > 1:
            self.cuda.comm_route_peer(on)
            self.peer = on

    def close(self) -> None:
        if self.world > 1:
            for arena in self.__dict__.pop("_arenas", {}).values():
                arena.release()
            self.cuda.comm_destroy()

    # ---- the three exchange patterns of the path -------------------------------------------------------------------------

    def full_sum(self, shard_tensor):
        """Tensor.sum over a row-sharded tensor: local deterministic reduction, then all-reduce of ONE float"""
        cuda = self.cuda
        if self.world > 1 and self.peer:
            # fused: the reduction kernel's last block completes the all-reduce over peer memory itself (one launch)
            src = shard_tensor.doBuffer()
            n = 1
            for s in shard_tensor.shape:
                n *= s
            out = cuda.Buffer.alloc(1)
            cuda.reduce_sum_allreduce(src, n, out)
            src.release()
            return out
        part = shard_tensor.sum().doBuffer()
        if self.world > 1:
            cuda.allreduce_sum(part, 1)
        return part

    def axis0_sum(self, local_column_sums):
        """column sums when the sharded axis is the reduced one. `local_column_sums` is the lazy tensor
        `shard.split(0).reduce(_ + _)` built once by the caller: local partial sums, then all-reduce of one row"""
        cuda = self.cuda
        part = local_column_sums.doBuffer()
        if self.world > 1:
            n = 1
            for s in local_column_sums.shape:
                n *= s
            cuda.allreduce_sum(part, n)
        return part

    def axis1_sum(self, local_row_sums, gather: bool = True):
        """row sums (`shard.split(1).reduce(_ + _)`): purely local; all-gather only if every rank wants the whole vector"""
        cuda = self.cuda
        part = local_row_sums.doBuffer()
        if self.world == 1 or not gather:
            return part
        n = 1
        for s in local_row_sums.shape:
            n *= s
        self._require_equal_blocks(n)
        whole = cuda.Buffer.alloc(n * self.world)
        cuda.allgather(part, whole, n)
        part.release()
        return whole

    def gather_arena(self, n_floats: int):
        """a symmetric (peer-mapped) buffer of at least n_floats, kept for reuse; collective on first use of a size"""
        arenas = self.__dict__.setdefault("_arenas", {})
        if n_floats not in arenas:
            arenas[n_floats] = self.cuda.comm_symmetric_alloc(n_floats)
        return arenas[n_floats]

    def matmul_rows(self, a_shard, b_full, m_shard: int, n: int, k: int, gather: bool = False, fused: bool | None = None):
        """C[rows of this rank, :] = A[rows of this rank, :] @ B — A and C row-sharded, B replicated; no exchange unless gathered.
        With gather=True and the peer mailboxes mapped (equal shards, N % 4 == 0) the all-gather is fused into the contraction's
        epilogue (TMA stores into every rank's copy over NVLink); the result is then a view of the communicator's arena, valid
        until the next fused call of the same size. fused=False forces contraction + ncclAllGather."""
        cuda = self.cuda
        if fused is None:
            fused = self.peer
        if gather:
            self._require_equal_blocks(m_shard * n)
        if gather and self.world > 1 and fused and self.peer and n % 4 == 0:
            whole = self.gather_arena(m_shard * n * self.world)
            cuda.matmul_3xtf32_allgather(a_shard, b_full, whole, m_shard, n, k)
            return whole.share()
        c = cuda.Buffer.alloc(m_shard * n)
        cuda.matmul_3xtf32(a_shard, b_full, c, m_shard, n, k)
        if self.world == 1 or not gather:
            return c
        whole = cuda.Buffer.alloc(m_shard * n * self.world)
        cuda.allgather(c, whole, m_shard * n)
        c.release()
        return whole
